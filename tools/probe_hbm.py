"""HBM probes on the GPU box: copy (read+write), write-only (fill / memset) and read-only bandwidth, to put
the write-bound kernels' roofline fractions in context (MEASURED_PEAKS.json is a COPY figure)."""
import torch, time
dev = torch.device("cuda:0")
n = 4 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
tc = t(lambda: b.copy_(a)); print("copy  GB/s (r+w)", 2 * n / tc / 1e9)
tf = t(lambda: a.fill_(7)); print("fill  GB/s (w)  ", n / tf / 1e9)
tz = t(lambda: a.zero_()); print("zero  GB/s (w)  ", n / tz / 1e9)
af = a.view(torch.float32)
tr = t(lambda: af.sum()); print("sum   GB/s (r)  ", n / tr / 1e9)
