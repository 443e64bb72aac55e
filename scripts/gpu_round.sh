# full GPU check: tests, smoke, default bench, per-algo values, ncu evidence
tag=${1:-r02v1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>&1; echo "ref rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_reference.json
bash scripts/gpu_algos.sh
bash scripts/gpu_profile.sh $tag
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_devplan_launches.csv python scripts/gpu_devplan_time.py 5 4096 2 > /dev/null 2>&1
python scripts/gpu_timeline.py 0 > gpurun_out/${tag}_timeline.log 2>&1; tail -9 gpurun_out/${tag}_timeline.log
python scripts/gpu_config4.py 8192 5 > gpurun_out/${tag}_config4.json 2>&1; cut -c1-300 gpurun_out/${tag}_config4.json
