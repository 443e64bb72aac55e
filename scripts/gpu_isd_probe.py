import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_deepfake_audio_detection_b200 import workload
from scl_deepfake_audio_detection_b200.engine import Engine
eng = Engine(0); args = workload.default_args()
B, L = 4096, 64600
x = torch.empty((B, L), device="cuda").normal_(0, 0.1); x[1::2] *= 20
ln = torch.full((B,), L, dtype=torch.int32, device="cuda")
seeds = list(range(B))
dp = eng.draw_device_plan(ln, seeds, 16000, args, 2, L)
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
gb = 2 * B * L * 4 / 1e9
for name, fn in (("torch copy", lambda: y.copy_(x)), ("normwav always=0", lambda: eng.normwav(x, ln, False, out=y)),
                 ("normwav always=1", lambda: eng.normwav(x, ln, True, out=y)), ("algo 2", lambda: eng.process(2, x, ln, dp, out=y))):
    ms = t(fn); print(f"{name:18s} {ms:.3f} ms  {gb / ms:.2f} TB/s (2*B*L*4 bytes)")
