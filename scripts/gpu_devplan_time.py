import sys
import numpy as np, torch
sys.path.insert(0, ".")
from scl_deepfake_audio_detection_b200 import workload
from scl_deepfake_audio_detection_b200.engine import Engine
eng = Engine(0)
args = workload.default_args()
algo, B = int(sys.argv[1]), int(sys.argv[2])
ln = torch.full((B,), 64600, dtype=torch.int32, device="cuda")
seeds = torch.from_numpy(np.array([workload.seed_for(u) for u in range(B)], dtype=np.uint32).view(np.int32)).cuda()
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dp = eng.draw_device_plan(ln, seeds, 16000, args, algo, 64600)
    e1.record()
    torch.cuda.synchronize()
    print(f"algo {algo} B={B} device plan draw {e0.elapsed_time(e1):.3f} ms", flush=True)
