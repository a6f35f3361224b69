# round 2, run F (gpurun --gpus 8): the scaling series 8 and 4 (the driver runs 1/2/4/8 at round end) with mixdown.check on every rank count
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/f_bench_n$n.json 2> gpurun_out/f_bench_n$n.err; echo "bench n$n rc=$?"; tail -c 300 gpurun_out/f_bench_n$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/f_bench_n$n.json').read().strip().splitlines()[-1]); print('n$n', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'mixdown', d['mixdown']['value'], d['mixdown'].get('check'))"
done
