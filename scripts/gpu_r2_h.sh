mkdir -p gpurun_out
echo "== TMA streamer"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
RAWBOOST_B200_LIB=$PWD/scl-deepfake-audio-detection_b200/lib/var_dbg.so python scripts/gpu_stream_debug.py 1024 | tail -12
for so in librawboost_b200; do echo "=== tma $so"; RAWBOOST_B200_LIB=$PWD/scl-deepfake-audio-detection_b200/lib/$so.so timeout 300 python scripts/gpu_isd_probe.py 4096 1024 2>&1 | grep -v "Warning\|torch copy"; done | tee gpurun_out/r02h_stream_tma.log
