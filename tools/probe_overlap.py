"""Does the ALU-bound heights kernel overlap with the store-bound emit kernel when they run on two streams?
Two independent plans (separate buffers), so there are no dependencies: t(both) vs t(emit) + t(heights)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import herbolution_b200 as hb
ctx = hb.Context(0, seed=0)
regA = hb.Region(-128, -128, 256, 256, -3, 6)
regB = hb.Region(1000, 1000, 256, 256, -3, 6)
pa, pb = ctx.region_plan(regA), ctx.region_plan(regB)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
for p, s in ((pa, sa), (pb, sb)):
    p.enqueue(hb.STAGE_ALL_MESH, s.cuda_stream)
torch.cuda.synchronize()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    sa.synchronize(); sb.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def emit(): pa.enqueue(hb.STAGE_EMIT, sa.cuda_stream)
def heights(): pb.enqueue(hb.STAGE_HEIGHTS | hb.STAGE_COUNT, sb.cuda_stream)
def both(): emit(); heights()
def both2(): heights(); emit()
import time
for name, fn in (("emit", emit), ("heights+count", heights), ("both (emit first)", both), ("both (heights first)", both2)):
    # host-side wall clock with full sync (events on the legacy stream would not see the side streams)
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(f"{name:24s} {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms")
