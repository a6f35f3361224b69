# round 2, run R: voice patches compiled per stage list (K8f): parity both ways, the polysynth bench leg, one ncu capture of the generated kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_patch.py tests/test_cpp_dropin.py -m gpu -q -x > gpurun_out/r_pytest_patch.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/r_pytest_patch.log | head -20; tail -15 gpurun_out/r_pytest_patch.log
timeout 400 python bench.py --workload patch --steps 20 --warmup 3 > gpurun_out/r_bench_patch.json 2> gpurun_out/r_bench_patch.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r_bench_patch.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r_bench_patch.json").read().strip().splitlines()[-1])
    print("patch fused", d["value"], "frac", d["roofline"]["frac"], "ms", d["ms_per_step"], "interp", d["interpreter"]["value"], "x", d["interpreter"]["fused_speedup"],
          "e2e", d["e2e"]["value"], d["e2e"]["frac_of_resident"], "compile_s", d["compile_seconds"], "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mxb_fused_patch -c 1 -o gpurun_out/r_fused_patch python bench.py --workload patch --steps 3 --warmup 3 --no-cpu > gpurun_out/r_ncu.log 2>&1; tail -3 gpurun_out/r_ncu.log
