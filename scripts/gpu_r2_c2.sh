# round 2, run C2: maxiChorus stage on the GPU (both executors), whole patch file again
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_patch.py -m gpu -q > gpurun_out/c2_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/c2_pytest.log | head; tail -15 gpurun_out/c2_pytest.log
