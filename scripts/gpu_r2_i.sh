mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:norm_stream_tma -s 3 -c 1 -f -o gpurun_out/r02i_tma_algo2_b1024 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs --algo 2 --batch 1024 > gpurun_out/r02i_a.log 2>&1
ls -la gpurun_out | grep r02i
