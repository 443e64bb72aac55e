mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( time timeout 1200 python bench.py > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -3 gpurun_out/r02g_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02g_bench_default.json"))
print("value", d["value"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e"]["frac_of_copy_ceiling"], {k: v["value"] for k, v in d["e2e"]["variants"].items()})
for k, v in d.get("configs", {}).items():
    print(k, v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), v.get("error"))
PY
bash scripts/gpu_profile.sh r02g > /dev/null 2>&1
ls gpurun_out | grep r02g
