"""Per-row timeline of the TMA streaming kernel (library built with -DRB_STREAM_DEBUG): when each row was claimed by a finisher
team, when its last tile arrived, when the team was done."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_deepfake_audio_detection_b200 import workload
from scl_deepfake_audio_detection_b200.engine import Engine
eng = Engine(0); args = workload.default_args()
B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 64600
x = torch.empty((B, L), device="cuda").normal_(0, 0.1)
ln = torch.full((B,), L, dtype=torch.int32, device="cuda")
dp = eng.draw_device_plan(ln, list(range(B)), 16000, args, 2, L)
y = torch.empty_like(x)
for _ in range(3):
    eng.process(2, x, ln, dp, out=y)
torch.cuda.synchronize()
ws = eng._ws
base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
raw = ws[base + (B + 1) * 8: base + (B + 1) * 8 + B * 32].cpu().numpy().view(np.uint64).reshape(B, 4).astype(np.int64)
t0 = raw[:, 0].min()
claim, ready, done, cta = (raw[:, 0] - t0) / 1e3, (raw[:, 1] - t0) / 1e3, (raw[:, 2] - t0) / 1e3, raw[:, 3]
print("rows", B, "kernel span us", done.max())
for r in list(range(0, B, max(1, B // 32))) + [B - 1]:
    print(f"row {r:5d} cta {cta[r]:4d} claimed {claim[r]:8.1f} ready {ready[r]:8.1f} done {done[r]:8.1f}  wait {ready[r]-claim[r]:7.1f} work {done[r]-ready[r]:7.1f}")
print("work us: mean %.1f median %.1f max %.1f; ready-time of last row %.1f" % ((done - ready).mean(), np.median(done - ready), (done - ready).max(), ready.max()))
