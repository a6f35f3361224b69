# round 2, run J: K4s with L1 steering (frames bypass L1, B fragments stay), whole groups of four k-steps; on-box 1R:2W bandwidth beside K2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/j_pytest.log | head -30; tail -3 gpurun_out/j_pytest.log
one() {  # workload tag steps
  timeout 300 python bench.py --workload $1 --steps $3 --warmup 5 --no-cpu --no-extras > gpurun_out/j_bench_$1_$2.json 2> gpurun_out/j_bench_$1_$2.err; rc=$?
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/j_bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1 $2', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'], d['roofline'].get('onbox_peaks'))
except Exception as e: print('$1 $2 rc=$rc', e, open('gpurun_out/j_bench_$1_$2.err').read()[-300:])"
}
one mfcc base 30
one delay base 40
one svf base 100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_stream -s 3 -c 1 -f -o gpurun_out/prof_r02_stft_stream_v6 python bench.py --workload mfcc --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-stft rc=$?
