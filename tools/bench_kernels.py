"""Secondary measurements: BASELINE configs 2 (batched worldgen, 4096 chunks) and 3 (meshing of 4096
pre-generated chunks) through the device-pointer C ABI (hb_dev_generate / hb_dev_mesh), CUDA events on a
dedicated stream.  torch is used only to own device memory and events.

    python tools/bench_kernels.py [--reps 20] [--scale 1]     (scale s: 32s x 32s columns)
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import herbolution_b200 as hb  # noqa: E402
from herbolution_b200 import _lib  # noqa: E402
from herbolution_b200._lib import check  # noqa: E402


def measure(ctx, scale: int = 1, reps: int = 20) -> dict:
    """configs 2 and 3 (+ codec) on `ctx`'s device: 32*scale x 32*scale columns x 4 layers (scale 1 = BASELINE's 4096 chunks)."""
    class A:  # noqa: D401
        pass
    a = A()
    a.reps, a.scale = reps, scale
    lib = _lib.load()
    side = 32 * a.scale
    region = hb.Region(-side // 2, -side // 2, side, side, -2, 4)
    pts = region.chunk_points()
    n = len(pts)
    cols = np.stack(np.meshgrid(np.arange(side) - side // 2, np.arange(side) - side // 2, indexing="ij"), -1).reshape(-1, 2).astype(np.int32)
    col_of = (np.arange(n) // region.ncy).astype(np.int32)  # chunk index = (ix*ncz+iz)*ncy+iy
    dev = torch.device(f"cuda:{ctx.device}")
    st = torch.cuda.Stream(dev)
    s = C.c_void_p(st.cuda_stream)

    def dbuf(arr):
        return torch.from_numpy(arr.view(np.uint8).reshape(-1)).to(dev)

    d_cols, d_pts, d_colof = dbuf(cols), dbuf(pts), dbuf(col_of)
    d_heights = torch.empty(len(cols) * 1024 * 4, dtype=torch.uint8, device=dev)
    d_ids = torch.empty(n * 32768 * 2, dtype=torch.uint8, device=dev)
    d_cubes = torch.empty(n * 32768 * 8, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def timed(fn):
        for _ in range(3):
            fn()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.reps):
            fn()
        e1.record(st)
        st.synchronize()
        return e0.elapsed_time(e1) / a.reps

    gen_h = lambda: check(lib.hb_dev_generate(ctx._h, s, p(d_cols), len(cols), None, None, 0, p(d_heights), None, None, None))  # noqa: E731
    gen_i = lambda: check(lib.hb_dev_generate(ctx._h, s, p(d_cols), len(cols), p(d_pts), p(d_colof), n, p(d_heights), None, p(d_ids), None))  # noqa: E731
    gen_c = lambda: check(lib.hb_dev_generate(ctx._h, s, p(d_cols), len(cols), p(d_pts), p(d_colof), n, p(d_heights), None, None, p(d_cubes)))  # noqa: E731
    t_h, t_i, t_c = timed(gen_h), timed(gen_i), timed(gen_c)
    out = {"config2": {"chunks": n, "columns": len(cols), "ms_heights": t_h, "ms_heights_plus_fill_ids": t_i,
                       "ms_fill_ids": t_i - t_h, "fill_ids_GBs": n * 65536 / ((t_i - t_h) * 1e-3) / 1e9,
                       "ms_fill_cubes": t_c - t_h, "fill_cubes_GBs": n * 262144 / ((t_c - t_h) * 1e-3) / 1e9,
                       "chunks_per_s_ids": n / (t_i * 1e-3)}}

    # config 3: mesh the generated ids
    nbr = ctx.build_neighbor_index(pts)
    d_nbr = dbuf(nbr)
    ws = torch.empty(lib.hb_dev_mesh_workspace_bytes(n), dtype=torch.uint8, device=dev)
    d_off = torch.empty(n * 8, dtype=torch.uint8, device=dev)
    d_cnt = torch.empty(n * 4, dtype=torch.uint8, device=dev)
    d_tot = torch.zeros(4, dtype=torch.int64, device=dev)
    check(lib.hb_dev_mesh(ctx._h, s, p(d_pts), n, p(d_ids), p(d_nbr), p(ws), None, 0, p(d_off), p(d_cnt), p(d_tot)))
    st.synchronize()
    span, quads = int(d_tot[0]), int(d_tot[1])
    d_out = torch.empty(max(span, 4) * 100, dtype=torch.uint8, device=dev)
    mesh_count = lambda: check(lib.hb_dev_mesh(ctx._h, s, p(d_pts), n, p(d_ids), p(d_nbr), p(ws), None, 0, p(d_off), p(d_cnt), p(d_tot)))  # noqa: E731
    mesh_all = lambda: check(lib.hb_dev_mesh(ctx._h, s, p(d_pts), n, p(d_ids), p(d_nbr), p(ws), p(d_out), span, p(d_off), p(d_cnt), p(d_tot)))  # noqa: E731
    t_cnt, t_all = timed(mesh_count), timed(mesh_all)
    st.synchronize()
    assert int(d_tot[2]) == 0
    algo_bytes = n * 78608 + quads * 100
    out["config3"] = {"chunks": n, "quads": quads, "ms_pack_count_scan": t_cnt, "ms_total": t_all, "ms_emit": t_all - t_cnt,
                      "chunks_per_s": n / (t_all * 1e-3), "algorithmic_bytes": algo_bytes,
                      "algorithmic_GBs": algo_bytes / (t_all * 1e-3) / 1e9,
                      "pack_read_GBs_if_all_pack": n * 65536 / (t_cnt * 1e-3) / 1e9}
    # codec: run-length encode the same ids (server/src/chunk/codec.rs)
    d_roff = torch.empty(n * 8, dtype=torch.uint8, device=dev)
    d_rcnt = torch.empty(n * 4, dtype=torch.uint8, device=dev)
    check(lib.hb_dev_rle_encode(ctx._h, s, p(d_ids), n, p(ws), None, 0, p(d_roff), p(d_rcnt), p(d_tot)))
    st.synchronize()
    rspan = int(d_tot[0])
    d_rle = torch.zeros(max(rspan, 4) * 3, dtype=torch.uint8, device=dev)
    rle_count = lambda: check(lib.hb_dev_rle_encode(ctx._h, s, p(d_ids), n, p(ws), None, 0, p(d_roff), p(d_rcnt), p(d_tot)))  # noqa: E731
    rle_all = lambda: check(lib.hb_dev_rle_encode(ctx._h, s, p(d_ids), n, p(ws), p(d_rle), rspan, p(d_roff), p(d_rcnt), p(d_tot)))  # noqa: E731
    t_rc, t_ra = timed(rle_count), timed(rle_all)
    st.synchronize()
    assert int(d_tot[2]) == 0
    out["codec"] = {"chunks": n, "records": rspan, "encoded_MB": rspan * 3 / 1e6, "ratio": n * 65536 / (rspan * 3),
                    "ms_count_scan": t_rc, "ms_total": t_ra, "chunks_per_s": n / (t_ra * 1e-3),
                    "read_GBs": 2 * n * 65536 / (t_ra * 1e-3) / 1e9}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--scale", type=int, default=1)
    a = ap.parse_args()
    ctx = hb.Context(0, seed=0)
    print(json.dumps(measure(ctx, a.scale, a.reps)))
    ctx.close()


if __name__ == "__main__":
    main()
