mkdir -p gpurun_out
python scripts/gpu_pcie.py 2>&1 | tee gpurun_out/pcie.log
timeout 900 python -m pytest tests/test_gpu_devplan.py -m gpu -x -q 2>&1 | tail -3
for ch in 0 592 1184; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --host-chunk $ch --e2e-steps 5 > gpurun_out/bench_chunk$ch.log 2>&1; echo "bench chunk=$ch rc=$?"
  tail -1 gpurun_out/bench_chunk$ch.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'],'copy+kern',d['e2e']['copy_and_kernels_only_value'])" || tail -5 gpurun_out/bench_chunk$ch.log
done
