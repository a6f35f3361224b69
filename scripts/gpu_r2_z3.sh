# round 2, run Z3: K2 out + mix with an 8-row mix tile (16 warps/SM by shared memory) against the 16-row tile (12 warps/SM), device-timed
# (--mix 1: the headline value is the out + mix kernel) and end to end; parity tests on the variant; memcheck of the modulated K2 kernels
mkdir -p gpurun_out
for i in 1 2; do
for v in "" mix8; do
  if [ -n "$v" ]; then export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_$v.so; else unset MXB_LIB_PATH; fi
  timeout 300 python bench.py --workload delay --mix 1 --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('delay+mix ${v:-mix16}', d['value'], round(d['roofline']['frac'],4), d['roofline'].get('launch_ms_median'), 'e2e', d['e2e']['value'])"
done
done
export MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_mix8.so
timeout 400 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "modulated or delay or trigger or config2 or per_sample" > gpurun_out/z3_pytest_mix8.log 2>&1; tail -3 gpurun_out/z3_pytest_mix8.log
unset MXB_LIB_PATH
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "modulated_frequency_and_cutoff_with_a_delay_line" > gpurun_out/z3_memcheck_k2_mod.log 2>&1; tail -4 gpurun_out/z3_memcheck_k2_mod.log
# the banks with a per-sample frequency stream, timed
timeout 300 python bench.py --workload modulated --steps 20 --warmup 5 > gpurun_out/z3_bench_modulated.json 2> gpurun_out/z3_bench_modulated.err; tail -c 1800 gpurun_out/z3_bench_modulated.json; tail -3 gpurun_out/z3_bench_modulated.err
