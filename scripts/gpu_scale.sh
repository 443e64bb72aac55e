# bench.py on N GPUs of one box exactly as the driver launches it (N > 1: BASELINE config 5, global batch 65536, strong scaling)
# usage: bash scripts/gpu_scale.sh <tag> <N> [steps]
tag=$1; n=$2; steps=${3:-10}
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps $steps --warmup 3 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err ) 2>&1 | grep real
echo "rc=$?"; tail -3 gpurun_out/${tag}_bench_${n}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_${n}gpu.json"))
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "scaling")}, d["roofline"]["frac"], d["parity"])
e = d["e2e"]
print("e2e", e["value"], e["ms_per_step"], "frac of ceiling", e["frac_of_copy_ceiling"], "ceiling", e["copy_ceiling"]["value"], e["copy_ceiling"]["aggregate_gb_s_each_way"], "GB/s each way")
print("variants", {k: round(v["value"]) for k, v in e["variants"].items()})
PY
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $n --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_${n}gpu.json 2>/dev/null ) 2>&1 | grep real; cut -c1-260 gpurun_out/${tag}_bench_reference_${n}gpu.json
