# round 2, last one-GPU run of the finished tree: the whole GPU suite, smoke(), the driver's bench command; then the K2 mix-tile A/B
# (8 rows = 16 warps/SM against 16 rows = 12 warps/SM, out + mix kernel device-timed with --mix 1), the variant's parity tests and a
# memcheck pass over the modulated K2 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/final_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
    print('svf', d['value'], round(d['roofline']['frac'], 4), 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
    for k, v in d['workloads'].items():
        if 'value' in v: print(k, v['value'], round(v['roofline']['frac'], 4), 'e2e', v.get('e2e', {}).get('value'))
        else: print(k, {kk: (vv.get('value'), round(vv['roofline']['frac'], 4)) for kk, vv in v.items() if isinstance(vv, dict) and 'roofline' in vv})
    print('mixdown', d['mixdown']['value'])
except Exception as e:
    print('bench line unreadable:', e)
PY
for i in 1 2; do
for v in "" mix8; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  timeout 300 python bench.py --workload delay --mix 1 --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('delay+mix ${v:-mix16}', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])"
done
done
export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_mix8.so
timeout 400 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "modulated or delay or trigger or config2 or per_sample" > gpurun_out/final_pytest_mix8.log 2>&1; tail -2 gpurun_out/final_pytest_mix8.log
unset MXB_LIB_PATH
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "modulated_frequency_and_cutoff_with_a_delay_line" > gpurun_out/final_memcheck_k2_mod.log 2>&1; tail -3 gpurun_out/final_memcheck_k2_mod.log
