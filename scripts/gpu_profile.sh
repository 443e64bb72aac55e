# ncu evidence for one bench configuration: launch list (shares of the step) + one full capture of the FIR kernel.
# usage: bash scripts/gpu_profile.sh <tag> [bench args...]
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fir_bank -s 3 -c 1 -f -o gpurun_out/${tag}_fir \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/${tag}_fir_bench.log 2>&1
ls -la gpurun_out/
