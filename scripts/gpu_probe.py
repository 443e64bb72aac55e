"""FP32-pipe probes: FFMA / FFMA2 register-resident chains and the FIR loop's operand pattern (rb_probe_fp32 modes)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scl_deepfake_audio_detection_b200 import _lib

lib = _lib.load()
sink = torch.zeros(4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
names = {0: "FFMA  a*acc+b (a,b fixed)", 1: "FFMA2 a*acc+b (a,b fixed)", 2: "FFMA2 FIR pattern: tap shared by 10, window, acc",
         3: "FFMA2 three distinct register pairs"}
for mode in (0, 1, 2, 3):
    best = 0.0
    fl = C.c_double(0)
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.rb_probe_fp32(mode, 4000, sink.data_ptr(), C.byref(fl), C.c_void_p(st))
        e1.record()
        e1.synchronize()
        assert rc == 0
        if i:
            best = max(best, fl.value / e0.elapsed_time(e1) / 1e9)
    print(f"mode {mode}: {best:7.2f} TFLOP/s   {names[mode]}")
