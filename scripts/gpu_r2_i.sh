# round 2, run I: K4s with the straight-line logarithm and the mel tiles shared by halves; K2 write-back wait relaxed, L2 evict-first A/B;
# mix-down of K1 / K2 on the fp64 tensor cores (headline bench with e2e and mixdown)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/i_pytest.log | head -30; tail -3 gpurun_out/i_pytest.log
one() {  # workload tag steps
  timeout 300 python bench.py --workload $1 --steps $3 --warmup 5 --no-cpu --no-extras > gpurun_out/i_bench_$1_$2.json 2> gpurun_out/i_bench_$1_$2.err; rc=$?
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/i_bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1 $2', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])
except Exception as e: print('$1 $2 rc=$rc', e, open('gpurun_out/i_bench_$1_$2.err').read()[-300:])"
}
one mfcc base 30
one svf base 100
python -c "
import json; d=json.loads(open('gpurun_out/i_bench_svf_base.json').read().strip().splitlines()[-1]); print('mixdown', d['mixdown']['value'], d['mixdown'].get('fp64_pipe_frac_per_gpu'), 'e2e', d['mixdown']['e2e']['value'])"
for v in "" l2h "" l2h; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  one delay "${v:-base}" 40
done
unset MXB_LIB_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_stream -s 3 -c 1 -f -o gpurun_out/prof_r02_stft_stream_v5 python bench.py --workload mfcc --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-stft rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:delay_bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_delay_bulk_v3 python bench.py --workload delay --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-delay rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_bank_outmix_dmma python bench.py --workload svf --mix 1 --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-outmix rc=$?
