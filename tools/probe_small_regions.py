import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import herbolution_b200 as hb
ctx = hb.Context(0, seed=0)
def sink(*a): pass
for name, reg in (("loader batch 33x33x11", hb.Region(-16, -16, 33, 33, -5, 11)), ("16x16x6 (config 1)", hb.Region(-8, -8, 16, 16, -3, 6)), ("4x4x6", hb.Region(0, 0, 4, 4, -3, 6))):
    for mode, fn in (("expanded", lambda: ctx.generate_mesh_region(reg, sink=sink)), ("packed sink", lambda: ctx.generate_mesh_region_packed(reg, sink=sink)), ("expanded+ids", lambda: ctx.generate_mesh_region(reg, want_ids=True, sink=sink))):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(10): tot = fn()
        dt = (time.perf_counter() - t0) / 10
        print(f"{name:24s} {mode:14s} {dt*1e3:8.3f} ms  {tot['chunks']/dt/1e6:7.3f} M chunks/s  {tot['quads']} quads")
