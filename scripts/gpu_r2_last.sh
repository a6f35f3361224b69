# round 2, the last GPU seconds: the whole GPU suite on the tree as committed (default library: 8-row K2 mix tile, straight-line increments, stream look-ahead)
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/last_pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/last_pytest_gpu.log | head; tail -2 gpurun_out/last_pytest_gpu.log
