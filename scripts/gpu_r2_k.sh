# round 2, run K: K4s stores the next assembly buffer from the last frame's registers; K1 mix tile of 8 rows at 7 CTAs per SM; on-box 1R:2W bandwidth
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/k_pytest.log | head -30; tail -3 gpurun_out/k_pytest.log
one() {  # workload tag steps
  timeout 300 python bench.py --workload $1 --steps $3 --warmup 5 --no-cpu --no-extras > gpurun_out/k_bench_$1_$2.json 2> gpurun_out/k_bench_$1_$2.err; rc=$?
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/k_bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1 $2', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'], d['roofline'].get('onbox_peaks'))
except Exception as e: print('$1 $2 rc=$rc', e, open('gpurun_out/k_bench_$1_$2.err').read()[-300:])"
}
one mfcc base 30
one delay base 40
one svf base 100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_stream -s 3 -c 1 -f -o gpurun_out/prof_r02_stft_stream_v7 python bench.py --workload mfcc --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-stft rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/k_bench_svf_base.json').read().strip().splitlines()[-1]); print('mixdown', d['mixdown']['value'], d['mixdown'].get('fp64_pipe_frac_per_gpu'), 'e2e', d['mixdown']['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_bank_outmix_v2 python bench.py --workload svf --mix 1 --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-outmix rc=$?
