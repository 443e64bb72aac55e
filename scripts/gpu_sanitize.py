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
# round 2: in-place normWav, in-place view assembly with labels, PCM16 / device-sink streaming, rows beyond 65536 samples
z = x.clone()
eng.normwav(z, ln, False, out=z)
G, nvoc = 1, 2
rows = torch.from_numpy(multiview.item_view_rows(G, nvoc).reshape(-1)).cuda()
lab = torch.from_numpy(multiview.item_labels(nvoc))
multiview.assemble_ex(eng, x[:3].contiguous(), y[:3].contiguous(), rows, ln[:3].contiguous(), 2 * (nvoc + 1), [3], 2000, True,
                      multiview.LAYOUT_MODEL, view_label=lab)
B8, L8 = 6, 4096
pcm = rs.randint(-20000, 20000, size=(B8, L8)).astype(np.int16)
l8 = np.array([4096, 4000, 37, 1, 2561, 4096], np.int32)
s8 = np.arange(6, dtype=np.uint32)
sink = torch.zeros((B8, L8), device="cuda")
for algo in (5, 2):
    eng.wait_host(eng.submit_host_ex(algo, pcm, "pcm16", l8, s8, 16000, args, out=sink))
    eng.wait_host(eng.submit_host_ex(algo, pcm, "pcm16", l8, s8, 16000, args, out=np.zeros((B8, L8), np.float32)))
    eng.process_device_seeded(algo, sink.clone(), torch.from_numpy(l8).cuda(), torch.from_numpy(s8.view(np.int32)).cuda(), 16000, args)
long_lengths = [70001, 300, 65537]
lw = [(0.5 * rs.standard_normal(n)).astype(np.float32) for n in long_lengths]
xl, lnl_ = eng.pack_waveforms(lw)
for algo in (2, 5):
    dpl = eng.draw_device_plan(lnl_, [7, 8, 9], 16000, args, algo, xl.shape[1])
    eng.process(algo, xl, lnl_, dpl)
torch.cuda.synchronize()
print("sanitize script done")
