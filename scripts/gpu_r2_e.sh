# round 2: new tests (streaming forms, batcher), stream variants, full bench with config 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for so in scl-deepfake-audio-detection_b200/lib/librawboost_b200.so scl-deepfake-audio-detection_b200/lib/var_w*.so; do
  echo "=== $so"
  RAWBOOST_B200_LIB=$PWD/$so timeout 300 python scripts/gpu_isd_probe.py 4096 1024 2>&1 | grep -v "Warning\|torch copy"
done | tee gpurun_out/r02e_isd_stream_variants.log
( time timeout 1200 python bench.py > gpurun_out/r02e_bench_default.json 2> gpurun_out/r02e_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -5 gpurun_out/r02e_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02e_bench_default.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d["parity"], "clocks", d["clocks"])
print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("includes", "copy_ceiling", "variants")})
print("ceiling", d["e2e"].get("copy_ceiling"))
print("variants", d["e2e"].get("variants"))
for k, v in d.get("configs", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error")}, "frac", (v.get("roofline") or {}).get("frac"), "parity", v.get("parity"))
print("cpu", d.get("cpu_baseline"))
PY
