# round 2, run Z: the modulated K2 window body and the C++ layer's per-sample signatures (new parity tests), and the delay leg again
# (the block-constant windows must keep their time: 2.33e11 / 0.855 before the change)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "modulated or refused or delay or trigger or config2" > gpurun_out/z_pytest_bank.log 2>&1; tail -4 gpurun_out/z_pytest_bank.log
timeout 300 python -m pytest tests/test_cpp_dropin.py -m gpu -q -x > gpurun_out/z_pytest_cpp.log 2>&1; tail -4 gpurun_out/z_pytest_cpp.log
for i in 1 2; do
timeout 300 python bench.py --workload delay --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('delay', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])"
done
