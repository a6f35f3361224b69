# round 2, run F2: the store pattern's own ceiling (a bank of bare phasors) beside fill_ / copy_ and the saw->SVF bank, same box, same run
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --workload svf --steps 100 --warmup 5 --no-cpu --no-extras 2>gpurun_out/f2_err.log | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; print('svf', d['value'], round(r['frac'],4), r['launch_ms_median'], 'GB/s', round(r['achieved']), 'onbox', {k:(round(v) if isinstance(v,float) else v) for k,v in r['onbox_peaks'].items() if k!='how'}, 'of pattern', r.get('frac_of_onbox_pattern_write'), 'of fill', r.get('frac_of_onbox_fill'))"
done
tail -3 gpurun_out/f2_err.log
