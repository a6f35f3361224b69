# round 2, run N (final, 1 GPU): the whole GPU suite, smoke, the default bench line (all configurations) and the reference arm exactly as the
# driver runs them, the launch list of the default command, compute-sanitizer passes over the kernels with mbarriers / named barriers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/n_pytest.log | head -30; tail -3 gpurun_out/n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/n_bench_default.json 2> gpurun_out/n_bench_default.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/n_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/n_bench_default.json').read().strip().splitlines()[-1])
print('svf', d['value'], round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
print('mixdown', d['mixdown']['value'], 'e2e', d['mixdown']['e2e']['value'])
for k,w in d['workloads'].items(): print(k, w['value'], round(w['roofline']['frac'],4), 'e2e', w['e2e']['value'], 'cpu', w['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference > gpurun_out/n_bench_reference.json 2> gpurun_out/n_bench_reference.err; echo "bench reference rc=$?"; tail -c 400 gpurun_out/n_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/n_launches_default.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1; echo launch-list rc=$?
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_spectral.py -m gpu -q -k "stream_kernel_mfcc_only_whole_hops and 7-512" > gpurun_out/n_sanitizer_racecheck_stft.log 2>&1; tail -3 gpurun_out/n_sanitizer_racecheck_stft.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_spectral.py -m gpu -q -k "stream_kernel_mfcc_only_whole_hops and 7-512" > gpurun_out/n_sanitizer_memcheck_stft.log 2>&1; tail -3 gpurun_out/n_sanitizer_memcheck_stft.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_bank.py -m gpu -q -k "delayline_index_exact and 1024-False" > gpurun_out/n_sanitizer_memcheck_delay.log 2>&1; tail -3 gpurun_out/n_sanitizer_memcheck_delay.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_bank.py -m gpu -q -k "delayline_index_exact and 1024-False" > gpurun_out/n_sanitizer_synccheck_delay.log 2>&1; tail -3 gpurun_out/n_sanitizer_synccheck_delay.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_bank.py -m gpu -q -k "delayline_index_exact and 1024-False" > gpurun_out/n_sanitizer_racecheck_delay.log 2>&1; tail -3 gpurun_out/n_sanitizer_racecheck_delay.log
