# multi-GPU check of bench.py (gpurun --gpus 8 -- bash scripts/gpu_scale_check.sh); every command bounded
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_gpu_exchange.py -q 2>&1 | tail -2
for n in 8 4 2; do
  timeout 200 $TR --nproc-per-node $n --master-port $((29510+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  echo "N=$n rc=$?"
done
timeout 120 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
timeout 200 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 --collective nccl > gpurun_out/scale_8_nccl.json 2> gpurun_out/scale_8_nccl.err; echo "nccl rc=$?"
