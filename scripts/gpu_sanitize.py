"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_deepfake_audio_detection_b200 import multiview, plans, workload
from scl_deepfake_audio_detection_b200.engine import Engine

eng = Engine(0)
args = workload.default_args()
lengths = [1, 37, 2561, 4097, 9000, 20000]
rs = np.random.RandomState(0)
waves = [(0.7 * rs.standard_normal(n)).astype(np.float32) for n in lengths]
seeds = list(range(100, 100 + len(lengths)))
for algo in range(0, 9):
    bp = plans.draw_batch(lengths, 16000, args, algo if algo else 5, seeds=seeds)
    x, ln = eng.pack_waveforms(waves, ld=bp.ld)
    y = eng.process(algo, x, ln, eng.upload_plan(bp))
    if algo:
        dp = eng.draw_device_plan(ln, seeds, 16000, args, algo, bp.ld)
        y2 = eng.process(algo, x, ln, dp)
    torch.cuda.synchronize()
    xh = x.cpu().numpy()
    eng.set_host_chunk(4)
    eng.process_host_seeded(algo, xh, np.array(lengths, np.int32), seeds, 16000, args)
    eng.process_host(algo, xh, bp)
    print("algo", algo, "ok", float(y.abs().max()))
eng.normwav(x, ln, True)
taps = torch.randn(6 * 700, device="cuda") / 700
off = torch.arange(0, 7, dtype=torch.int32, device="cuda") * 700
eng.filter_fir(x, ln, taps, off)
out, olen = multiview.assemble(eng, x, ln, 3, [0, 5], 3000, True, multiview.LAYOUT_MODEL)
out, olen = multiview.assemble(eng, x, ln, 3, [0, 5], 3000, False, multiview.LAYOUT_ITEM)
torch.cuda.synchronize()
print("sanitize script done")
