# round 2: streaming-kernel variants (sub-tile pipelining), the restructured bench (all sub-records), the reference arm
mkdir -p gpurun_out
for so in scl-deepfake-audio-detection_b200/lib/librawboost_b200.so scl-deepfake-audio-detection_b200/lib/var_v*.so; do
  echo "=== $so"
  RAWBOOST_B200_LIB=$PWD/$so timeout 300 python scripts/gpu_isd_probe.py 4096 1024 2>&1 | grep -v "Warning\|torch copy"
done | tee gpurun_out/r02d_isd_stream_variants.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time timeout 1200 python bench.py > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -5 gpurun_out/r02d_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02d_bench_default.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", d["parity"])
print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("includes", "copy_ceiling", "variants")})
print("ceiling", d["e2e"].get("copy_ceiling"))
for k, v in d.get("configs", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error")}, "frac", (v.get("roofline") or {}).get("frac"), "parity", v.get("parity"))
print("cpu", d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_reference.json 2>&1; cut -c1-300 gpurun_out/r02d_bench_reference.json
