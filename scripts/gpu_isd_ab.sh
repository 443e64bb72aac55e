for so in scl-deepfake-audio-detection_b200/lib/var_isd_old.so scl-deepfake-audio-detection_b200/lib/librawboost_b200.so; do
  for rep in 1 2; do for b in 1024 4096; do RAWBOOST_B200_LIB=$PWD/$so python bench.py --algo 2 --batch $b --no-e2e --no-cpu --steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so'[-22:], 'b$b', round(d['value']), round(d['ms_per_step'],4))"; done; done
done
