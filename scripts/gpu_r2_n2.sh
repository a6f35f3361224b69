# round 2, last two-GPU run (gpurun --gpus 2): the peer-memory exchange for banks AND voice patches (tests incl. the silent-peer time-out),
# smoke() with its 2-rank exchange, and a 2-rank bench with mixdown.check on the finished tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q > gpurun_out/n2_pytest_exchange.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/n2_pytest_exchange.log | head; tail -3 gpurun_out/n2_pytest_exchange.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"; tail -c 300 gpurun_out/n2_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'mixdown', d['mixdown']['value'], d['mixdown'].get('check'))"
