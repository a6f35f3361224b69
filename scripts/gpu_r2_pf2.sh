# round 2, run PF2: look-ahead on the per-sample frequency stream in K1 and K2 (modulated instantiations): 0 / 4 / 8 (lib/) steps (a fourth library does not fit the 512 MiB snapshot)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "fm or cutoff or modulated" 2>&1 | tail -2
for i in 1 2; do
for v in pf0 pf4 ""; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  timeout 300 python bench.py --workload modulated --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('${v:-pf8}', 'fm_svf', d['fm_svf']['ms_per_step'], round(d['fm_svf']['roofline']['frac'],4), 'fm_delay', d['fm_delay']['ms_per_step'], round(d['fm_delay']['roofline']['frac'],4))"
done
done
