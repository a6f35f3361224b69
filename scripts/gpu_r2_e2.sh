# round 2, run E2: K1 with the store flavour decided outside the sample loop -- bank parity, headline kernels (twice), one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bank.py -m gpu -q -x > gpurun_out/e2_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/e2_pytest.log | head; tail -2 gpurun_out/e2_pytest.log
for i in 1 2; do
timeout 300 python bench.py --workload svf --steps 100 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('svf', d['value'], round(d['roofline']['frac'],4), d['roofline']['launch_ms_median'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_resident'], 'mixdown', d['mixdown']['value'], d['mixdown']['e2e']['value'], d['clocks'])"
done
timeout 300 python bench.py --workload biquad --steps 100 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('biquad', d['value'], round(d['roofline']['frac'],4), d['roofline']['launch_ms_median'], 'e2e', d['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bank_kernel -s 4 -c 1 -o gpurun_out/e2_bank_svf python bench.py --workload svf --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/e2_ncu.log 2>&1; tail -1 gpurun_out/e2_ncu.log
