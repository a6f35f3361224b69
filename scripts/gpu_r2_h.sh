# round 2, run H: occupancy probe of the streaming STFT kernel (occ1 = 1 CTA per SM) and FFT-only timing; spectral tests again
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectral.py -m gpu -q 2>&1 | tail -2
timeout 200 python scripts/stft_probe.py 2>&1 | tail -1
MXB_LIB_PATH=$PWD/maximilian_b200/lib_exp/libmaxib200_occ1.so timeout 200 python scripts/stft_probe.py 2>&1 | tail -1
timeout 300 python bench.py --workload mfcc --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('mfcc', d['value'], round(d['roofline']['frac'],4))"
