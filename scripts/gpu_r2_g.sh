# round 2, run G: K4s with 4-accumulator tensor-core tiles; K2 staging depth / CTA shape A/B (library variants built by
# scripts/build_variant_full.sh: s3 = 3 stages, s4 = 4 stages, t64s3 = 64-thread CTAs with 3 stages); K1 256-thread A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/g_pytest.log | head -30; tail -3 gpurun_out/g_pytest.log
one() {  # workload tag steps
  timeout 300 python bench.py --workload $1 --steps $3 --warmup 5 --no-cpu --no-extras > gpurun_out/g_bench_$1_$2.json 2> gpurun_out/g_bench_$1_$2.err; rc=$?
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/g_bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1 $2', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])
except Exception as e: print('$1 $2 rc=$rc', e, open('gpurun_out/g_bench_$1_$2.err').read()[-300:])"
}
one mfcc base 30
for v in "" s3 s4 t64s3 "" s3; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  one delay "${v:-base}" 40
done
unset MXB_LIB_PATH
# the delay-line tests once with each staging depth (same tests, other library)
for v in s3 s4 t64s3; do
  MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so timeout 600 python -m pytest tests/test_gpu_bank.py -m gpu -q -k "delay or ring" 2>&1 | tail -1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_stream -s 3 -c 1 -f -o gpurun_out/prof_r02_stft_stream_v4 python bench.py --workload mfcc --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-stft rc=$?
for v in s3 s4; do
MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:delay_bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_delay_bulk_$v python bench.py --workload delay --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-delay-$v rc=$?
done
for v in "" b256 "" b256; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  one svf "${v:-base}" 100
done
unset MXB_LIB_PATH
