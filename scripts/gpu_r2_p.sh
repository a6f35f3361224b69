# round 2, run P: wave-quantisation probe of K2 -- voices chosen so that the 128-voice CTAs fill whole waves of 6 CTAs x 148 SMs
mkdir -p gpurun_out
for v in 262144 227328 113664 340992; do
  MXB_BENCH_VOICES=$v timeout 300 python bench.py --workload delay --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('delay V=$v', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'))"
done
