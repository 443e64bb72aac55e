"""FIR-bank kernel sweep on the GPU: achieved TFLOP/s of rb_filter_fir vs tap count, and of the LnL bank.
usage: python scripts/gpu_fir_sweep.py [B]   (library chosen by RAWBOOST_B200_LIB)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scl_deepfake_audio_detection_b200 import _lib, plans, workload
from scl_deepfake_audio_detection_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
L = 64600
eng = Engine(0)
lib = eng.lib
x = torch.rand(B, L, device="cuda") * 2 - 1
ln = torch.full((B,), L, dtype=torch.int32, device="cuda")
y = torch.zeros_like(x)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


print("lib", _lib.LIB_PATH)
for K in (51, 131, 271, 491, 503, 1001):
    taps = torch.randn(B * K, device="cuda") / K
    off = torch.arange(0, B + 1, dtype=torch.int32, device="cuda") * K
    ms = timeit(lambda: eng.filter_fir(x, ln, taps, off, out=y))
    if K == 271:  # correctness spot-check of this build against numpy (closed form of filterFIR)
        xs, ts, ys = x[3].cpu().numpy().astype(np.float64), taps[3 * K:4 * K].cpu().numpy().astype(np.float64), y[3].cpu().numpy()
        ref = np.convolve(xs, ts)[(K + 1) // 2:(K + 1) // 2 + L]
        print("   max-abs error vs numpy:", float(np.abs(ys - ref).max()))
    print(f"filter_fir K={K:5d}  {ms:8.3f} ms  {2.0 * B * L * K / ms / 1e9:7.2f} TFLOP/s")

args = workload.default_args()
pool = None
bp = plans.draw_batch([L] * 256, 16000, args, 5, seeds=[workload.seed_for(u) for u in range(256)])
# tile the 256 drawn plans over B utterances
reps = B // 256
import copy
taps = np.concatenate([bp.lnl_taps] * reps)
sizes = np.tile(np.diff(bp.lnl_tap_off), reps)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
isd_sizes = np.tile(np.diff(bp.isd_off), reps)
big = plans.BatchPlan(B=B, ld=L, lengths=np.full(B, L, np.int32), n_f=5, lnl_taps=taps, lnl_tap_off=off,
                      isd_off=np.concatenate([[0], np.cumsum(isd_sizes)]).astype(np.int32), isd_idx=np.concatenate([bp.isd_idx] * reps),
                      isd_fr=np.concatenate([bp.isd_fr] * reps), g_sd=2.0)
dp = eng.upload_plan(big)
for algo in (1, 5):
    lib.rb_profile_read(None, None, 1)
    lib.rb_profile_enable(1)
    ms = timeit(lambda: eng.process(algo, x, ln, dp, out=y))
    lib.rb_profile_enable(0)
    fm, fn_ = C.c_double(0), C.c_uint64(0)
    lib.rb_profile_read(C.byref(fm), C.byref(fn_), 1)
    print(f"algo {algo}: step {ms:8.3f} ms  {B / ms * 1e3:9.0f} utt/s   fir kernel {fm.value / fn_.value:8.3f} ms  "
          f"{big.fir_flops() / (fm.value / fn_.value) / 1e9:7.2f} TFLOP/s")
