# round 2, run PF: look-ahead (prefetch.global.L2) on the per-sample frequency stream of K1's modulated instantiations: 0 / 8 (lib/) / 16 steps
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "fm or cutoff or modulated" 2>&1 | tail -2
for i in 1 2; do
for v in pf0 "" pf16; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  timeout 300 python bench.py --workload modulated --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('${v:-pf8}', 'fm_svf', d['fm_svf']['value'], d['fm_svf']['ms_per_step'], round(d['fm_svf']['roofline']['frac'],4))"
done
done
