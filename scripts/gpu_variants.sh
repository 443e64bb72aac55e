# run the FIR sweep for every library variant in lib/
mkdir -p gpurun_out
for so in scl-deepfake-audio-detection_b200/lib/*.so; do
  echo "=== $so"
  RAWBOOST_B200_LIB=$PWD/$so timeout 300 python scripts/gpu_fir_sweep.py ${1:-2048} 2>&1 | grep -v Warning
done | tee gpurun_out/variants.log
