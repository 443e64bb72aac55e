# round 2, first GPU call: parity of the new streaming ISD / normWav path, its variants, a quick bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -15 gpurun_out/r02c_pytest_gpu.log
for so in scl-deepfake-audio-detection_b200/lib/librawboost_b200.so scl-deepfake-audio-detection_b200/lib/var_u*.so; do
  echo "=== $so"
  RAWBOOST_B200_LIB=$PWD/$so timeout 300 python scripts/gpu_isd_probe.py 4096 1024 2>&1 | grep -v Warning
done | tee gpurun_out/r02c_isd_stream_variants.log
timeout 600 python bench.py --algo 2 --batch 1024 --no-e2e --no-cpu > gpurun_out/r02c_bench_algo2_b1024.json 2> gpurun_out/r02c_bench_algo2_b1024.err; cut -c1-400 gpurun_out/r02c_bench_algo2_b1024.json
