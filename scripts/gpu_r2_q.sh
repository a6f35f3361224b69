# round 2, run Q: K2 with a staggered start of the warps (variants stag1 / stag3: up to 7 / 21 us), at one wave and at the configured size
mkdir -p gpurun_out
for v in "" stag1 stag3; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  for V in 262144 113664; do
  MXB_BENCH_VOICES=$V timeout 300 python bench.py --workload delay --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('delay ${v:-base} V=$V', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'))"
  done
done
