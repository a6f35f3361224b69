# round 2, run V: K7s against the two-kernel path (timing), one ncu capture of it
mkdir -p gpurun_out
for tp in 0 1; do
  MXB_ISTFT_TWO_PASS=$tp timeout 300 python bench.py --workload spectral --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
for k in ('analysis_mags_phases','resynthesis','analysis_octave_bark'): print('two_pass=$tp', k, d[k]['value'], round(d[k]['roofline']['frac'],4), d[k]['ms_per_step'])"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:istft1024 -c 1 -o gpurun_out/v_istft1024 python bench.py --workload spectral --steps 3 --warmup 3 > gpurun_out/v_ncu.log 2>&1; tail -2 gpurun_out/v_ncu.log
