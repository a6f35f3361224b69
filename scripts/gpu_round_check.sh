# One command for everything measured on one B200 (run with: gpurun --timeout 1200 -- bash scripts/gpu_round_check.sh).
# Every step is bounded by its own timeout; results land in gpurun_out/.
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in svf biquad delay mfcc; do timeout 240 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"; done
timeout 240 python bench.py --workload svf --mix 1 --no-cpu > gpurun_out/bench_svf_mix.json 2>/dev/null
timeout 240 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null
NCU="timeout 240 ncu --set full --clock-control none --import-source on"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
$NCU -k regex:bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_f_svf python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
$NCU -k regex:bank_kernel -s 10 -c 1 -f -o gpurun_out/prof_f_mixdown python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
$NCU -k regex:bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_f_svf_mix python bench.py --steps 5 --warmup 3 --no-cpu --mix 1 > /dev/null 2>&1
$NCU -k regex:delay_bank_kernel -s 5 -c 1 -f -o gpurun_out/prof_f_delay python bench.py --workload delay --steps 5 --warmup 4 --no-cpu > /dev/null 2>&1
$NCU -k regex:stft_kernel -s 3 -c 1 -f -o gpurun_out/prof_f_stft python bench.py --workload mfcc --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -12
