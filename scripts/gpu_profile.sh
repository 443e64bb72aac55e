# ncu evidence for the shipped build: launch list of the default bench step + full captures of the dominant kernels.
# usage: bash scripts/gpu_profile.sh <tag>     (reports land in gpurun_out/; summarise them here with scripts/ncu_summarise.py)
tag=$1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_algo5_b4096.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/${tag}_launches_bench.log 2>&1
cap() {  # name kernel-regex bench-args...
  name=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/${tag}_${name} \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs "$@" > gpurun_out/${tag}_${name}_bench.log 2>&1
}
cap fir_algo5_b4096 fir_bank --algo 5 --batch 4096
cap fir_algo3_b1024 fir_bank --algo 3 --batch 1024
cap stream_algo2_b1024 norm_stream --algo 2 --batch 1024
cap stream_algo2_b4096 norm_stream --algo 2 --batch 4096
ls -la gpurun_out/ | grep ${tag}
