mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "filter_fir" 2>&1 | tail -2
for so in librawboost_b200 var_rul3 var_rul10; do
  export RAWBOOST_B200_LIB=$PWD/scl-deepfake-audio-detection_b200/lib/$so.so
  timeout 300 python scripts/gpu_ssi_probe.py 5:4096 1:4096 3:4096 3:1024 2>&1 | grep -v Warning
  timeout 300 python scripts/gpu_fir_sweep.py 2048 2>&1 | grep "filter_fir"
done | tee gpurun_out/r02m_register_usage_level.log
