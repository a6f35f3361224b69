# round 2, one-GPU run after the straight-line oscillator increment (div_rn_unchecked / rcp_rn_unchecked): the whole GPU suite (with the
# operand-level comparison against the operators, tests/test_gpu_ieee_div.py), the modulated banks and the polysynth patch timed again
# (before: fm_svf 4.35 ms, fm_delay 3.00 ms, polysynth 4.295 ms per block), and the driver's bench command for the record
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final3_pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/final3_pytest_gpu.log | head; tail -3 gpurun_out/final3_pytest_gpu.log
timeout 300 python bench.py --workload modulated --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
for k in ('fm_svf','fm_delay'): print(k, d[k]['value'], d[k]['ms_per_step'], round(d[k]['roofline']['frac'],4))"
timeout 300 python bench.py --workload patch --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('patch', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'interp', d.get('interpreter'))"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final3_bench_n1.json 2> gpurun_out/final3_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/final3_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
    print('svf', d['value'], round(d['roofline']['frac'], 4), 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
    for k, v in d['workloads'].items():
        if 'value' in v: print(k, v['value'], round(v['roofline']['frac'], 4), 'e2e', v.get('e2e', {}).get('value'))
        else: print(k, {kk: (vv.get('value'), round(vv['roofline']['frac'], 4)) for kk, vv in v.items() if isinstance(vv, dict) and 'roofline' in vv})
    print('mixdown', d['mixdown']['value'])
except Exception as e:
    print('bench line unreadable:', e)
PY
