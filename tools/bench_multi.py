"""BASELINE config 5 (1024 x 1024 chunk columns x 6) from ONE process through hb_multi (include/herbgpu.h): every
visible GPU takes an x-slab, the serialised sink counts what it receives.  Wall clock around the blocking call.
    python tools/bench_multi.py [--columns 1024] [--reps 2]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import herbolution_b200 as hb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--devices", type=int, default=0, help="0 = all visible")
    a = ap.parse_args()
    n = a.columns
    region = hb.Region(-(n // 2), -(n // 2), n, n, -3, 6)
    m = hb.MultiContext(devices=None if a.devices == 0 else list(range(a.devices)), seed=0)
    stats = {"span": 0, "chunks": 0, "calls": 0}

    def sink(first, off, cnt, inst, ids):
        stats["span"] += len(inst)
        stats["chunks"] += len(cnt)
        stats["calls"] += 1

    m.generate_mesh_region(hb.Region(-8, -8, 16 * m.n_devices, 16, -3, 6), sink=sink)  # warm-up: contexts, pools, plans
    best = None
    for _ in range(a.reps):
        stats.update(span=0, chunks=0, calls=0)
        t0 = time.perf_counter()
        tot = m.generate_mesh_region(region, sink=sink)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(json.dumps({"api": "hb_multi_generate_mesh_region (one process)", "n_gpus": m.n_devices, "columns": n,
                      "chunks": tot["chunks"], "quads": tot["quads"], "seconds": best,
                      "chunks_per_s": tot["chunks"] / best, "instance_gb_delivered": stats["span"] * 100 / 1e9,
                      "host_write_gbs": stats["span"] * 100 / best / 1e9, "sink_calls": stats["calls"],
                      "slabs": [m.slab(region, i) for i in range(m.n_devices)]}))
    m.close()


if __name__ == "__main__":
    main()
