# round 2, run D2: the oscillators' wrap test on the integer pipe -- bank parity, then the kernels the fp64 pipe bounds (out+mix = the e2e loop, mix-only) and the patch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bank.py tests/test_gpu_patch.py -m gpu -q -x > gpurun_out/d2_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/d2_pytest.log | head; tail -3 gpurun_out/d2_pytest.log
for i in 1 2; do
timeout 300 python bench.py --workload svf --steps 100 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('svf', d['value'], round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], d['e2e']['frac_of_resident'], 'mixdown', d['mixdown']['value'], d['mixdown']['e2e']['value'])"
done
timeout 300 python bench.py --workload svf --mix 1 --steps 100 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('svf out+mix', d['value'], round(d['roofline']['frac'],4), d['ms_per_step'])"
timeout 300 python bench.py --workload patch --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('patch', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
