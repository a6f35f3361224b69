# round 2, run A: parity (incl. the new full-size oracle-slice tests), smoke, the default bench line with every workload.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/a_pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/a_pytest_gpu.log | head -30; tail -3 gpurun_out/a_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; tail -2 gpurun_out/a_smoke.log
timeout 600 python bench.py > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/a_bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/a_bench_default.json").read().strip().splitlines()[-1])
    print("headline", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_resident"))
    for k, v in d.get("workloads", {}).items():
        print(k, v["value"], v["roofline"]["frac"], "e2e", v["e2e"]["value"], "cpu", v.get("cpu_baseline", {}).get("value"))
    print("mixdown", d["mixdown"]["value"], d["mixdown"]["e2e"]["value"]); print("cpu", d["cpu_baseline"])
except Exception as e:
    print("parse failed", e)
PY
