# round 2, 2 GPUs: the peer-memory exchange tests, the 2-rank smoke, the bench at N = 2 exactly as the driver launches it (+ the reference arm's rank handling)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/m2_bench_n2.json 2> gpurun_out/m2_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/m2_bench_n2.json').read().splitlines() if l.startswith('{')][-1])
print('n2', d['value'], round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], 'mixdown', d['mixdown']['value'], d['mixdown'].get('check'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -c 200
