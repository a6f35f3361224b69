# quick single-GPU check: every command bounded by its own timeout
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in mfcc svf delay; do timeout 120 python bench.py --workload $w --no-cpu --steps 50 --warmup 5 > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; echo "$w rc=$?"; done
timeout 120 python bench.py --workload svf --mix 1 --no-cpu --steps 50 --warmup 5 > gpurun_out/q_svf_mix.json 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 3 -c 1 -f -o gpurun_out/prof_stft8 python bench.py --workload mfcc --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1; echo ncu rc=$?
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mix_reduce -c 6 --csv --log-file gpurun_out/launches_reduce.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
