"""Device-resident timing of single algos for one library build (RAWBOOST_B200_LIB), through bench.py's own Bench.resident.
usage: python scripts/gpu_ssi_probe.py [algo:batch ...]   (default 3:1024 3:4096 5:4096)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

cases = [tuple(map(int, c.split(":"))) for c in (sys.argv[1:] or ["3:1024", "3:4096", "5:4096"])]
a = argparse.Namespace(length=64600)
b = bench.Bench(a)
print("lib", os.environ.get("RAWBOOST_B200_LIB", "(default)"))
for algo, B in cases:
    best = None
    for _ in range(3):
        r = b.resident(algo, B, 64600, 10, 3, parity_n=4)
        if best is None or r["ms_per_step"] < best["ms_per_step"]:
            best = r
    print(f"algo {algo} B={B:5d}  {best['ms_per_step']:8.4f} ms  {best['value'] / 1e3:9.1f} k utt/s  frac {best['roofline']['frac']:.4f}  "
          f"parity {best['parity']['max_abs_vs_oracle']:.2e} ok={best['parity']['ok']}", flush=True)
b.pool.terminate()
