"""BASELINE config 4: SCL multi-view generation -- 4 RawBoost (algo 5) views per bona fide sample as in
asvspoof_2019_augall_3 (1 on the anchor + 1 on each of the 3 vocoded copies), then the shared crop to 64000 samples and the
assembly of the 8 views per item in the model's layout [V, 64000]. Device-resident timing (CUDA events), G items per step.

usage: python scripts/gpu_config4.py [items=8192] [steps=5]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scl_deepfake_audio_detection_b200 import multiview, workload
from scl_deepfake_audio_detection_b200.engine import Engine

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
L, TRIM, V = 64600, 64000, 8
eng = Engine(0)
args = workload.default_args()
dev = eng.device
B = 4 * G  # RawBoost utterances: per item vocoded 1..3 then the anchor (the loader's order)
# synthetic waveforms generated on the device (speech-level gaussian / loud uniform alternating), ~8.5 GB at G=8192
gen = torch.Generator(device=dev).manual_seed(1)
x = torch.empty((B, L), dtype=torch.float32, device=dev)
x.normal_(0, 0.1, generator=gen).clamp_(-1, 1)
x[1::2].uniform_(-0.9, 0.9, generator=gen)
ln = torch.full((B,), L, dtype=torch.int32, device=dev)
seeds = torch.from_numpy(np.array([workload.seed_for(u) for u in range(B)], dtype=np.uint32).view(np.int32)).to(dev)
rs = np.random.RandomState(4)
starts = torch.tensor([int(rs.rand() * (L - TRIM)) for _ in range(G)], dtype=torch.int32, device=dev)
y = torch.empty_like(x)
views = torch.empty((G * V, L), dtype=torch.float32, device=dev)
vlen = torch.full((G * V,), L, dtype=torch.int32, device=dev)
out = torch.empty((G, V, TRIM), dtype=torch.float32, device=dev)


def step():
    dp = eng.draw_device_plan(ln, seeds, workload.SAMPLE_RATE, args, 5, L)   # plans drawn on the device from the seeds
    eng.process(5, x, ln, dp, out=y)
    # view order of the Dataset: anchor, augmented anchor, vocoded 1..3, augmented vocoded 1..3
    xs, ys, vw = x.view(G, 4, L), y.view(G, 4, L), views.view(G, V, L)
    vw[:, 0] = xs[:, 3]
    vw[:, 1] = ys[:, 3]
    vw[:, 2:5] = xs[:, 0:3]
    vw[:, 5:8] = ys[:, 0:3]
    multiview.assemble(eng, views, vlen, V, starts, TRIM, True, multiview.LAYOUT_MODEL, out=out)


for _ in range(2):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for i in range(steps):
    step()
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
med = float(np.median(ms))
print(json.dumps({"config": "BASELINE config 4: 4 RawBoost views per bona fide sample + shared crop + view assembly [V=8, 64000]",
                  "items_per_step": G, "rawboost_utterances_per_step": B, "ms_per_step_median": med, "ms_per_step": ms,
                  "items_per_s": G / med * 1e3, "augmented_utterances_per_s": B / med * 1e3,
                  "includes": "device plan draw from seeds + algo-5 kernels + view gather (torch copies) + rb_multiview_assemble",
                  "out_checksum": float(out[0, 1, :8].abs().sum().item())}))
