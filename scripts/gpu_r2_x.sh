# round 2, run X (final, 1 GPU): the whole GPU suite, smoke, the default bench line (all configurations) and the reference arm exactly as the
# driver runs them, the launch list of the default command, one full capture of the generated patch kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/x_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/x_pytest.log | head -30; tail -3 gpurun_out/x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/x_bench_default.json 2> gpurun_out/x_bench_default.err; echo "bench default rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -c 300 gpurun_out/x_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/x_bench_default.json').read().strip().splitlines()[-1])
print('svf', d['value'], round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
print('mixdown', d['mixdown']['value'], 'e2e', d['mixdown']['e2e']['value'])
for k,w in d['workloads'].items():
    if 'value' in w: print(k, w['value'], round(w['roofline']['frac'],4), 'e2e', w['e2e']['value'], 'cpu', w['cpu_baseline']['value'])
    else:
        for kk, vv in w.items():
            if isinstance(vv, dict) and 'value' in vv: print(k, kk, vv['value'], round(vv['roofline']['frac'],4), vv['ms_per_step'])"
timeout 600 python bench.py --impl reference > gpurun_out/x_bench_reference.json 2> gpurun_out/x_bench_reference.err; echo "bench reference rc=$?"; tail -c 300 gpurun_out/x_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/x_launches_default.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1; echo launch-list rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mxb_fused_patch -c 1 -o gpurun_out/x_fused_patch python bench.py --workload patch --steps 3 --warmup 3 --no-cpu > gpurun_out/x_ncu.log 2>&1; tail -1 gpurun_out/x_ncu.log
