"""PCIe floor for the e2e numbers: pinned H2D / D2H copy rates alone and overlapped (run on the GPU box)."""
import time, torch
def rate(fn, nbytes, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return nbytes / dt / 1e9, dt * 1e3
for mb in (4, 32, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def h2d():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    def both():
        h2d(); d2h()
    print(f"{mb} MiB  H2D {rate(h2d, n)[0]:.1f} GB/s  D2H {rate(d2h, n)[0]:.1f} GB/s  both {rate(both, 2*n)[0]:.1f} GB/s total, {rate(both, 2*n)[1]:.3f} ms")
