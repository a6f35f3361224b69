# the last GPU seconds of round 2: the golden replay of the modulated cases (incl. the two new ones on chains with a delay line)
timeout 40 python -m pytest tests/test_gpu_bank.py -m gpu -q -x -k "golden" 2>&1 | tail -3
