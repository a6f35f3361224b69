set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in svf biquad delay mfcc; do python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 600 gpurun_out/bench_$w.err; done
python bench.py --workload svf --mix 1 --no-cpu > gpurun_out/bench_svf_mix.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
NCU="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
$NCU -k regex:bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_f_svf python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
$NCU -k regex:bank_kernel -s 10 -c 1 -f -o gpurun_out/prof_f_mixdown python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
$NCU -k regex:bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_f_svf_mix python bench.py --steps 5 --warmup 3 --no-cpu --mix 1 > /dev/null 2>&1
$NCU -k regex:delay_bank_kernel -s 3 -c 1 -f -o gpurun_out/prof_f_delay python bench.py --workload delay --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
$NCU -k regex:stft_kernel -s 3 -c 1 -f -o gpurun_out/prof_f_stft python bench.py --workload mfcc --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -12
