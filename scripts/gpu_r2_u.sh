# round 2, run U: K7s (one-kernel maxiIFFT) parity + A/B against the two-kernel path, the cheaper warp mix reduction of the patches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_patch.py tests/test_cpp_dropin.py -m gpu -q > gpurun_out/u_pytest.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/u_pytest.log | head -20; tail -12 gpurun_out/u_pytest.log
for tp in "" 1; do
  MXB_ISTFT_TWO_PASS=$tp timeout 300 python bench.py --workload spectral --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
for k in ('analysis_mags_phases','resynthesis','analysis_octave_bark'): print('two_pass=${tp:-0}', k, d[k]['value'], round(d[k]['roofline']['frac'],4), d[k]['ms_per_step'])"
done
timeout 300 python bench.py --workload patch --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('patch', d['value'], round(d['roofline']['frac'],4), d['ms_per_step'], 'interp', d['interpreter']['value'], 'e2e', d['e2e']['value'])"
