mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02j_bench_2gpu.json 2> gpurun_out/r02j_bench_2gpu.err ) 2>&1 | grep real; echo "rc=$?"
tail -4 gpurun_out/r02j_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_bench_2gpu.json"))
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "scaling")}, d["config"]["workload"], d["roofline"]["frac"], d["parity"])
print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("includes", "copy_ceiling", "variants")})
print("ceiling", d["e2e"]["copy_ceiling"]["value"], d["e2e"]["copy_ceiling"]["aggregate_gb_s_each_way"])
print("variants", {k: v["value"] for k, v in d["e2e"]["variants"].items()})
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02j_ref_2gpu.json 2>/dev/null ) 2>&1 | grep real; cut -c1-200 gpurun_out/r02j_ref_2gpu.json
