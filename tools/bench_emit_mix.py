"""Emit-kernel rate by terrain layer mix: which part of k_mesh_region<EMIT> (region-bottom slab fast path, buried
skip, general tile path on surface chunks) runs at what fraction of the HBM write peak.  CUDA events, device-resident.

    python tools/bench_emit_mix.py [--columns 256]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import herbolution_b200 as hb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    ctx = hb.Context(0, seed=0)
    st = torch.cuda.Stream()
    n = a.columns
    out = []
    for name, cy0, ncy in [("all (config 4)", -3, 6), ("deep only: slab + region-top faces", -3, 2), ("surface layers -1..0", -1, 2),
                           ("surface layers -2..1", -2, 4), ("air only", 3, 2)]:
        reg = hb.Region(-n // 2, -n // 2, n, n, cy0, ncy)
        plan = ctx.region_plan(reg)
        for _ in range(3):
            plan.enqueue(hb.STAGE_ALL_MESH, st.cuda_stream)
        res = plan.result(st.cuda_stream)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(st)
        for _ in range(a.reps):
            plan.enqueue(hb.STAGE_EMIT, st.cuda_stream)
        ev[1].record(st)
        st.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / a.reps
        out.append({"layers": name, "chunks": res["n_chunks"], "quads": res["quads"], "emit_ms": ms,
                    "GBs": res["quads"] * 100 / (ms * 1e-3) / 1e9 if ms > 0 else None,
                    "quads_per_chunk": res["quads"] / res["n_chunks"]})
        plan.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
