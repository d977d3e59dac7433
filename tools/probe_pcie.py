"""Measure pinned D2H / H2D copy bandwidth and the host's core count on the GPU box (plumbing probe)."""
import os, time, torch
print("cores", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
n = 2 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for name, (dst, src) in {"d2h": (h, d), "h2d": (d, h)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(name, "GB/s", 4 * n / dt / 1e9)
t0 = time.perf_counter(); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); print("pin 2GiB s", time.perf_counter() - t0)
