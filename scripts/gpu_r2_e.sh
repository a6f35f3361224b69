# round 2, run E (gpurun --gpus 2): the peer-memory exchange across two GPUs -- tests (incl. the silent-peer time-out), smoke, a 2-rank
# bench with mixdown.check, and a compute-sanitizer pass over the two kernels that use mbarriers / named barriers.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q > gpurun_out/e_pytest_exchange.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/e_pytest_exchange.log | head; tail -3 gpurun_out/e_pytest_exchange.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err; echo "bench n2 rc=$?"; tail -c 300 gpurun_out/e_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/e_bench_n2.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['roofline']['frac'], 'mixdown', d['mixdown']['value'], d['mixdown'].get('check'))"
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_spectral.py -m gpu -q -k "stream_kernel_mfcc_only_whole_hops and 7-512" > gpurun_out/e_sanitizer_racecheck_stft.log 2>&1; tail -4 gpurun_out/e_sanitizer_racecheck_stft.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_bank.py -m gpu -q -k "delayline_index_exact and 1024-False" > gpurun_out/e_sanitizer_memcheck_delay.log 2>&1; tail -4 gpurun_out/e_sanitizer_memcheck_delay.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_bank.py -m gpu -q -k "delayline_index_exact and 1024-False" > gpurun_out/e_sanitizer_synccheck_delay.log 2>&1; tail -4 gpurun_out/e_sanitizer_synccheck_delay.log
