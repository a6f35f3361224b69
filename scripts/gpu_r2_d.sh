# round 2, run D: everything again after per-sample triggers / MOD+ENV / K2 14-warp CTAs; delay bench + capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/d_pytest.log | head -30; tail -3 gpurun_out/d_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --workload delay --steps 40 --warmup 5 --no-cpu > gpurun_out/d_bench_delay.json 2> gpurun_out/d_bench_delay.err; echo "delay rc=$?"; tail -c 300 gpurun_out/d_bench_delay.err
python -c "
import json; d=json.loads(open('gpurun_out/d_bench_delay.json').read().strip().splitlines()[-1]); print('delay', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:delay_bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_delay_bulk_v2 python bench.py --workload delay --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1; echo ncu-delay rc=$?
timeout 300 python bench.py --workload mfcc --steps 30 --warmup 5 --no-cpu > gpurun_out/d_bench_mfcc.json 2> gpurun_out/d_bench_mfcc.err; echo "mfcc rc=$?"; tail -c 300 gpurun_out/d_bench_mfcc.err
python -c "
import json; d=json.loads(open('gpurun_out/d_bench_mfcc.json').read().strip().splitlines()[-1]); print('mfcc', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_stream -s 3 -c 1 -f -o gpurun_out/prof_r02_stft_stream_v3 python bench.py --workload mfcc --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1; echo ncu-stft rc=$?
# A/B: 256-thread CTAs for K1 (4 KB per CTA-step) against the 128-thread build
for v in "" b256; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  for w in svf svf; do
    timeout 200 python bench.py --workload $w --no-cpu --no-extras --steps 100 --warmup 5 2>/dev/null | python -c "import sys,json; L=sys.stdin.read().splitlines(); J=[l for l in L if l.startswith(chr(123))]; print('variant=${v:-base} $w', (lambda d:(d['value'], round(d['roofline']['frac'],4), d['roofline']['launch_ms_median'], d['e2e']['value'], d['mixdown']['value']))(json.loads(J[-1])) if J else L[-3:])"
  done
done
unset MXB_LIB_PATH
