# round 2, run Z2: K2 with the modulated windows as their own kernel instantiations (the build in lib/) against the one-kernel build of
# run Z (lib_exp/libmaxib200_modw.so), same box, alternating; and the modulated parity tests on the new build
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "modulated or refused or delay or trigger or config2 or fm or cutoff" > gpurun_out/z2_pytest_bank.log 2>&1; tail -3 gpurun_out/z2_pytest_bank.log
for i in 1 2; do
for v in "" modw; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  timeout 300 python bench.py --workload delay --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('delay ${v:-split}', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])"
done
done
