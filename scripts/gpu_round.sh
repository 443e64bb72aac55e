# full GPU check: tests, smoke, bench (default + chunk sweep)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
for ch in 0 148 296; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --host-chunk $ch > gpurun_out/bench_chunk$ch.log 2>&1; echo "bench chunk=$ch rc=$?"
  tail -1 gpurun_out/bench_chunk$ch.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'],'copy+kern',d['e2e']['copy_and_kernels_only_value'])" || tail -5 gpurun_out/bench_chunk$ch.log
done
