# round 2, run M: K4s with only the last read of a sample steered past L1 (DRAM traffic check); spectral tests; launch list of the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectral.py tests/test_cpp_dropin.py -m gpu -q 2>&1 | tail -2
timeout 300 python bench.py --workload mfcc --steps 30 --warmup 5 --no-cpu --no-extras > gpurun_out/m_bench_mfcc.json 2> gpurun_out/m_bench_mfcc.err
python -c "
import json; d=json.loads(open('gpurun_out/m_bench_mfcc.json').read().strip().splitlines()[-1]); print('mfcc', d['value'], round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_stream -s 3 -c 1 -f -o gpurun_out/prof_r02_stft_stream_v8 python bench.py --workload mfcc --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1; echo ncu-stft rc=$?
