"""Device time of the HBM-bound operators (normWav, algo 2) against a plain device copy. usage: gpu_isd_probe.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_deepfake_audio_detection_b200 import workload
from scl_deepfake_audio_detection_b200.engine import Engine
eng = Engine(0); args = workload.default_args()
L = 64600
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
for B in [int(a) for a in sys.argv[1:]] or [4096, 1024]:
    x = torch.empty((B, L), device="cuda").normal_(0, 0.1); x[1::2] *= 9   # odd rows: peak ~4 -> rescaled
    ln = torch.full((B,), L, dtype=torch.int32, device="cuda")
    dp = eng.draw_device_plan(ln, list(range(B)), 16000, args, 2, L)
    n_imp = int(eng.download_plan(dp).isd_off[-1])
    y = torch.empty_like(x)
    gb = 2 * B * L * 4 / 1e9
    gb2 = gb + 12 * n_imp / 1e9
    xq = torch.empty((B, L), device="cuda").normal_(0, 0.1)   # no row above 1: the pure streaming pass
    for name, fn, g in (("torch copy", lambda: y.copy_(x), gb), ("normwav quiet rows", lambda: eng.normwav(xq, ln, False, out=y), gb),
                        ("algo 2 quiet rows", lambda: eng.process(2, xq, ln, dp, out=y), gb2),
                        ("normwav always=0", lambda: eng.normwav(x, ln, False, out=y), gb),
                        ("normwav always=1", lambda: eng.normwav(x, ln, True, out=y), gb),
                        ("algo 2", lambda: eng.process(2, x, ln, dp, out=y), gb2)):
        ms = t(fn); print(f"B={B:5d} {name:18s} {ms:.3f} ms  {g / ms:.2f} TB/s algorithmic ({g / ms / 6.5555 * 100:.0f} % of 6555 GB/s)")
