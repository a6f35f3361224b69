# round 2, run S: everything after the fused patches went in -- full parity suite, smoke, min-blocks A/B of the generated kernel, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s_pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)" gpurun_out/s_pytest_gpu.log | head -30; tail -3 gpurun_out/s_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; tail -2 gpurun_out/s_smoke.log
for n in 0 7 8; do
  MXB_PATCH_MINBLOCKS=$n timeout 300 python bench.py --workload patch --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('patch minblocks=$n', d['value'], round(d['roofline']['frac'],4), d['ms_per_step'], 'interp', d['interpreter']['value'], 'e2e', d['e2e']['value'])"
done
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/s_bench_default.json 2> gpurun_out/s_bench_default.err; echo "bench rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/s_bench_default.err; grep -v "^\s" gpurun_out/s_bench_default.err | tail -5
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/s_bench_default.json").read().strip().splitlines()[-1])
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
