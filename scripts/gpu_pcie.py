import torch, time
n = 1058406400 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.zeros(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    if both:
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for both in (False, True, True):
    dt = run(both); print("both" if both else "h2d only", f"{dt*1e3:.2f} ms", f"{n*4/dt/1e9:.1f} GB/s per direction")
