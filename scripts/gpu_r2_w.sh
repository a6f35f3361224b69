# round 2, run W: full-size parity of the new legs, sanitizer passes over the new kernels, K7s with the L1 prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_patch.py -m gpu -q -k "full_size or istft" > gpurun_out/w_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/w_pytest.log | head; tail -8 gpurun_out/w_pytest.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_spectral.py -m gpu -q -x -k "istft_fused_overlap_add and 512" > gpurun_out/w_racecheck_istft1024.log 2>&1; tail -4 gpurun_out/w_racecheck_istft1024.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_patch.py -m gpu -q -x -k "fused_equals_interpreted" > gpurun_out/w_memcheck_patch.log 2>&1; tail -4 gpurun_out/w_memcheck_patch.log
timeout 300 python bench.py --workload spectral --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
for k in ('analysis_mags_phases','resynthesis','analysis_octave_bark'): print(k, d[k]['value'], round(d[k]['roofline']['frac'],4), d[k]['ms_per_step'])"
