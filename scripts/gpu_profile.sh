# ncu evidence: launch list of the default bench step + full captures of the dominant kernel of algo 5 / 3 / 2.
# usage: bash scripts/gpu_profile.sh <tag>
tag=$1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_algo5_b4096.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_launches_bench.log 2>&1
cap() {  # name kernel-regex bench-args...
  name=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/${tag}_${name} \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/${tag}_${name}_bench.log 2>&1
}
cap fir_algo5_b4096 fir_bank --algo 5 --batch 4096
cap fir_algo3_b1024 fir_bank --algo 3 --batch 1024
cap isd_algo2_b1024 isd_fused --algo 2 --batch 1024
ls -la gpurun_out/ | grep ${tag}
