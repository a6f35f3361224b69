# round 2, run O: K2 write-back by plain coalesced stores (wbstg), and without the copy engine at all (notma: 16-byte cp.async loads), against the copy-engine pipeline
mkdir -p gpurun_out
one() {  # workload tag steps
  timeout 300 python bench.py --workload $1 --steps $3 --warmup 5 --no-cpu --no-extras > gpurun_out/o_bench_$1_$2.json 2> gpurun_out/o_bench_$1_$2.err; rc=$?
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/o_bench_$1_$2.json').read().strip().splitlines()[-1]); print('$1 $2', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])
except Exception as e: print('$1 $2 rc=$rc', e, open('gpurun_out/o_bench_$1_$2.err').read()[-300:])"
}
for v in "" wbstg notma "" wbstg notma; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  one delay "${v:-base}" 40
done
for v in wbstg notma; do MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so timeout 600 python -m pytest tests/test_gpu_bank.py -m gpu -q -k "delay or ring" 2>&1 | tail -1; done
MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_wbstg.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:delay_bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_delay_wbstg python bench.py --workload delay --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-delay rc=$?
MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_notma.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:delay_bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_delay_notma python bench.py --workload delay --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-delay-notma rc=$?
