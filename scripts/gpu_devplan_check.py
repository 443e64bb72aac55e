"""Device planner vs the native host planner (itself pinned to numpy by tests/test_host_logic.py): exactness + timing."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from scl_deepfake_audio_detection_b200 import workload
from scl_deepfake_audio_detection_b200.engine import Engine
from scl_deepfake_audio_detection_b200.native_planner import NativePlanner

eng = Engine(0)
args = workload.default_args()
host = NativePlanner(threads=0, pinned=False)
rng = np.random.RandomState(0)
for algo in (5, 2, 1, 3, 4, 6, 7, 8):
    for case in ("full", "ragged"):
        B = 24
        lengths = [64600] * B if case == "full" else [1, 2, 3, 31, 32, 33, 37, 600, 2560, 2561, 4097, 16000, 65536] + list(rng.randint(1, 65536, 11))
        ld = (max(lengths) + 3) // 4 * 4
        seeds = [workload.seed_for(1000 * algo + u) for u in range(B)]
        ref = host.draw(lengths, 16000, args, algo, seeds=seeds, ld=ld, copy=True)
        ln = torch.tensor(lengths, dtype=torch.int32, device="cuda")
        dp = eng.draw_device_plan(ln, seeds, 16000, args, algo, ld)
        torch.cuda.synchronize()
        got = eng.download_plan(dp)
        msgs = []
        for name in ("lnl_tap_off", "isd_off", "isd_idx", "isd_fr", "ssi_tap_off", "ssi_snr_db"):
            r, g = getattr(ref, name), getattr(got, name)
            if r is None:
                continue
            ok = g is not None and r.shape == g.shape and np.array_equal(r, g)
            msgs.append(f"{name}:{'EXACT' if ok else 'DIFF'}")
        for name in ("lnl_taps", "ssi_taps", "ssi_noise"):
            r, g = getattr(ref, name), getattr(got, name)
            if r is None:
                continue
            if g is None or r.shape != g.shape:
                msgs.append(f"{name}:SHAPE")
                continue
            d = np.abs(r.astype(np.float64) - g.astype(np.float64))
            rel = d / np.maximum(np.abs(r.astype(np.float64)), 1e-30)
            msgs.append(f"{name}:maxabs={d.max():.2e},nbitdiff={(r != g).mean():.2e},maxulp~{(d / np.maximum(np.spacing(np.abs(r)), 1e-45)).max():.1f}")
        print(algo, case, " ".join(msgs), flush=True)

# timing at bench size
for algo, B in ((5, 4096), (2, 4096), (3, 1024), (1, 4096)):
    lengths = [64600] * B
    ln = torch.tensor(lengths, dtype=torch.int32, device="cuda")
    seeds = torch.from_numpy(np.array([workload.seed_for(u) for u in range(B)], dtype=np.uint32).view(np.int32)).cuda()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dp = eng.draw_device_plan(ln, seeds, 16000, args, algo, 64600)
        e1.record()
        torch.cuda.synchronize()
        print(f"algo {algo} B={B} device plan draw {e0.elapsed_time(e1):.3f} ms", flush=True)
