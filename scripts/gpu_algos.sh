# value (device-resident) for every BASELINE algo at its config size; no e2e / cpu legs
mkdir -p gpurun_out
for cfg in "1 4096" "5 4096" "3 1024" "2 1024" "3 4096" "2 4096" "4 1024" "8 1024"; do
  set -- $cfg
  timeout 300 python bench.py --algo $1 --batch $2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/algo$1_b$2.log 2>&1
  echo "algo $1 batch $2 rc=$?"; tail -1 gpurun_out/algo$1_b$2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['bound'], r['frac'], r.get('kernel_share_of_step'), d['gpu_launches'])"
done
