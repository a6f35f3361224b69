# multi-GPU check of bench.py (run with: gpurun --gpus 8 -- bash scripts/gpu_scale_check.sh)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_exchange.py -q 2>&1 | tail -2
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29510+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  echo "N=$n rc=$?"; tail -c 400 gpurun_out/scale_$n.err
done
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
timeout 300 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 --collective nccl > gpurun_out/scale_8_nccl.json 2> gpurun_out/scale_8_nccl.err; echo "nccl rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 50 --warmup 5 --mix 1 > gpurun_out/scale_8_mix1.json 2> gpurun_out/scale_8_mix1.err; echo "mix1 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 30 --warmup 5 --workload mfcc > gpurun_out/scale_8_mfcc.json 2> gpurun_out/scale_8_mfcc.err; echo "mfcc rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29534 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/scale_2_ref.json 2> gpurun_out/scale_2_ref.err; echo "ref rc=$?"
