# round 2, run T: the default bench line (all workloads) with its wall time, and the reference arm
mkdir -p gpurun_out
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/t_bench_default.json 2> gpurun_out/t_bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"; grep -v "^\s" gpurun_out/t_bench_default.err | tail -5
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/t_bench_default.json").read().strip().splitlines()[-1])
    print("headline", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_resident"))
    for k, v in d.get("workloads", {}).items():
        if "value" in v:
            print(k, v["value"], v["roofline"]["frac"], "e2e", v["e2e"]["value"], "cpu", v.get("cpu_baseline", {}).get("value"))
        else:
            for kk, vv in v.items():
                if isinstance(vv, dict) and "value" in vv:
                    print(k, kk, vv["value"], vv["roofline"]["frac"], vv["ms_per_step"])
    print("mixdown", d["mixdown"]["value"], d["mixdown"]["e2e"]["value"]); print("cpu", d["cpu_baseline"])
except Exception as e:
    print("parse failed", e)
PY
t0=$(date +%s); timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/t_bench_ref.json 2> gpurun_out/t_bench_ref.err; echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"
