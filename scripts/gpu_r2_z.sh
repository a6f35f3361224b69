# round 2, run Z: trigger streams as packed bits (parity across the three element types, the patch leg with 32 MiB of triggers per block)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_patch.py -m gpu -q > gpurun_out/z_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/z_pytest.log | head; tail -3 gpurun_out/z_pytest.log
timeout 300 python bench.py --workload patch --steps 20 --warmup 3 > gpurun_out/z_bench_patch.json 2>gpurun_out/z_bench_patch.err; tail -c 400 gpurun_out/z_bench_patch.err
python -c "
import json
d=json.loads(open('gpurun_out/z_bench_patch.json').read().strip().splitlines()[-1]); print('patch', d['value'], round(d['roofline']['frac'],4), d['ms_per_step'], 'interp', d['interpreter']['value'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_resident'], 'cpu', d['cpu_baseline']['value'])"
