"""Does the device planner run beside the FIR-bank kernel? Times algo-5 filtering (stream A) and a device plan draw (stream B)
alone and together. usage: gpu_overlap_probe.py [B=4096]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scl_deepfake_audio_detection_b200 import workload
from scl_deepfake_audio_detection_b200.engine import Engine
eng = Engine(0); args = workload.default_args()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = 64600
x = torch.empty((B, L), device="cuda").normal_(0, 0.1)
ln = torch.full((B,), L, dtype=torch.int32, device="cuda")
seeds = torch.arange(B, dtype=torch.int32, device="cuda")
dp = eng.draw_device_plan(ln, seeds, 16000, args, 5, L)
y = torch.empty_like(x)
import ctypes as C
from scl_deepfake_audio_detection_b200 import _lib
a = _lib.args_struct(args, 16000)
need = int(eng.lib.rb_devplan_bytes(C.byref(a), 5, B, L))
store = torch.empty(need + 256, dtype=torch.uint8, device="cuda")
sptr = (store.data_ptr() + 255) // 256 * 256
def draw(stream):
    s = _lib.RbPlan()
    rc = eng.lib.rb_devplan_draw(C.byref(a), 5, B, L, C.c_void_p(ln.data_ptr()), C.c_void_p(seeds.data_ptr()), C.c_void_p(sptr), need, C.byref(s), C.c_void_p(stream.cuda_stream))
    assert rc == 0
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
def run(fir, plan, prio):
    sa = torch.cuda.Stream(priority=-1)
    sb = torch.cuda.Stream(priority=prio)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    sa.wait_event(e0); sb.wait_event(e0)
    keep = []
    for _ in range(3):
        if fir:
            with torch.cuda.stream(sa):
                eng.process(5, x, ln, dp, out=y)
        if plan:
            draw(sb)
    with torch.cuda.stream(sa): e1.record()
    with torch.cuda.stream(sb): e2.record()
    torch.cuda.synchronize()
    return max(e0.elapsed_time(e1), e0.elapsed_time(e2)) / 3, e0.elapsed_time(e1) / 3, e0.elapsed_time(e2) / 3
for name, f, p, prio in (("fir alone", 1, 0, 0), ("plan alone", 0, 1, 0), ("both, planner low priority", 1, 1, 0), ("both, planner high priority", 1, 1, -1)):
    run(f, p, prio)
    t = run(f, p, prio)
    print(f"{name:30s} {t[0]:7.3f} ms per step (fir stream {t[1]:7.3f}, plan stream {t[2]:7.3f})")
