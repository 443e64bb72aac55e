# round 2: planner co-residency (global top phase), long-row planner, full bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python scripts/gpu_devplan_time.py 5 4096 3 2>&1 | tail -5
( time timeout 1200 python bench.py > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -5 gpurun_out/r02f_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02f_bench_default.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d["parity"], "clocks", d["clocks"])
print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("includes", "copy_ceiling", "variants")})
print("variants", d["e2e"].get("variants"))
for k, v in d.get("configs", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error")}, "frac", (v.get("roofline") or {}).get("frac"), "parity", v.get("parity"))
print("cpu", d.get("cpu_baseline"))
PY
