# round 2, run Y: K2 with 8-slot chunks (half the staging per warp: 32 warps/SM by registers) at 2 / 3 / 4 stages against the 16-slot build
mkdir -p gpurun_out
for v in "" c8s2 c8s3 c8s4; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  if [ -n "$v" ]; then timeout 600 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "delay or ring or config2" 2>&1 | tail -2; fi
  for i in 1 2; do
  timeout 300 python bench.py --workload delay --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('delay ${v:-base}', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])"
  done
done
