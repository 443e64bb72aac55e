# Full GPU check of a build: parity tests, smoke, the bench line (all sub-records), the reference arm, sanitizers, ncu evidence.
# usage (on the GPU box, via gpurun): bash scripts/gpu_round.sh <tag>      -> everything lands in gpurun_out/<tag>_*
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log; tail -2 gpurun_out/${tag}_smoke.log
# ncu first: the bench line quotes roofline.traffic from profiles/traffic.json, keyed by the build id, so the captures of THIS build
# are taken and summarised before the line is produced (the summaries travel back through gpurun_out/)
bash scripts/gpu_profile.sh $tag > gpurun_out/${tag}_profile.log 2>&1
python scripts/ncu_summarise.py $tag fir_algo5_b4096:fir_bank_kernel:5:4096 fir_algo3_b1024:fir_bank_kernel:3:1024 \
    stream_algo2_b1024:norm_stream_kernel:2:1024 stream_algo2_b4096:norm_stream_kernel:2:4096 >> gpurun_out/${tag}_profile.log 2>&1
cp profiles/traffic.json gpurun_out/${tag}_traffic.json; cp profiles/${tag}_*_ncu_full.md gpurun_out/
timeout 1200 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_default.json
timeout 900 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_reference.json
timeout 300 python scripts/gpu_fir_sweep.py 2048 2>&1 | grep -v Warning > gpurun_out/${tag}_fir_sweep.log
timeout 300 python scripts/gpu_isd_probe.py 4096 1024 2>&1 | grep -v Warning > gpurun_out/${tag}_stream_probe.log
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python scripts/gpu_sanitize.py"
  timeout 900 compute-sanitizer --tool $tool python scripts/gpu_sanitize.py 2>&1 | grep -E "sanitize script done|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -8
done > gpurun_out/${tag}_compute_sanitizer.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_devplan_launches.csv python scripts/gpu_devplan_time.py 5 4096 2 > /dev/null 2>&1
ls gpurun_out | grep ${tag}
