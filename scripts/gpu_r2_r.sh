# round 2, run R: K4s with 20 (and 16) warps in one CTA per SM instead of 2 x 8 (variants sw20 / sw16), spectral tests on each
mkdir -p gpurun_out
for v in "" sw20 sw16; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  timeout 300 python bench.py --workload mfcc --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('mfcc ${v:-base}', d['value'], round(d['roofline']['frac'],4))"
  timeout 600 python -m pytest tests/test_gpu_spectral.py -m gpu -q 2>&1 | tail -1
done
