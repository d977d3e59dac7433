"""Timings of the device-resident world (hb_world_*) on the reference's own runtime workload, with the literal
CPU restatement (oracle ChunkMap, 1 thread) beside it on a bounded sample.

    python tools/bench_world.py [--radius 16] [--cpu-radius 6] [--no-cpu]

Workload: ChunkLoader::update (server/src/entity/components.rs:26-46): fill_rhombus(centre, 16) = 4 235 chunks
loaded at once, then the loader moves one chunk in +x (unload the trailing shell, load the leading one), then
dig/place edits.  Each phase: server load / set_cube -> sync (CubeUpdate wire) -> client remesh.
Host wall-clock around the blocking C-ABI calls (they synchronise internally).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def fill_rhombus(c, radius):
    """components.rs:61-73"""
    out = []
    for x in range(-radius, radius + 1):
        for y in range(-(radius // 3), radius // 3 + 1):
            for z in range(-radius, radius + 1):
                if abs(x) + abs(y) + abs(z) <= radius:
                    out.append((c[0] + x, c[1] + y, c[2] + z))
    return out


def timed(fn, *a):
    t = time.perf_counter()
    r = fn(*a)
    return (time.perf_counter() - t) * 1e3, r


def surface_edits(n, rng, span):
    import oracle as O
    w = O.world(0)
    cache = {}
    xyz = np.zeros((n, 3), dtype=np.int32)
    for i in range(n):
        wx, wz = int(rng.integers(-span, span)), int(rng.integers(-span, span))
        k = (wx >> 5, wz >> 5)
        if k not in cache:
            cache[k] = O.heights_from_noise(O.noise_tile2(w, *k)).reshape(32, 32)
        xyz[i] = (wx, int(cache[k][wz & 31, wx & 31]) - 1 - int(rng.integers(0, 3)), wz)
    return xyz, rng.choice(np.array([0, 0, 1, 2, 3], dtype=np.uint16), size=n)


def gpu_run(radius):
    import herbolution_b200 as hb
    ctx = hb.Context(0, seed=0)
    res = {}
    first = fill_rhombus((0, 0, 0), radius)
    second = fill_rhombus((1, 0, 0), radius)
    s1, s2 = set(first), set(second)
    pts = hb.context.as_chunk_pts  # list -> hb_chunk_pt array, outside the timed calls
    with hb.World(ctx, len(first) + 1024) as w:
        # warm-up on a throw-away world so allocation / first-launch costs are not in the numbers
        with hb.World(ctx, 64) as tmp:
            tmp.load(first[:32]); tmp.sync(); tmp.fetch_updates(); tmp.remesh()
        for rep in range(2):  # second pass re-loads into warm scratch buffers
            if rep:
                w.unload(list(s1 | s2))
            first_a = pts(first)
            t_load, _ = timed(w.load, first_a)
            t_sync, (nd, ne) = timed(w.sync)
            t_fetch, upd = timed(w.fetch_updates)
            t_mesh, out = timed(w.remesh_dev)
            res["loader_update"] = {"chunks": len(first), "load_ms": t_load, "sync_ms": t_sync, "fetch_updates_ms": t_fetch,
                                    "remesh_dev_ms": t_mesh, "dirty": nd, "entries": ne, "wire_MB": ne * 12 / 1e6,
                                    "quads": int(out["counts"].sum()),
                                    "chunks_per_s_load_sync_remesh": len(first) / ((t_load + t_sync + t_mesh) * 1e-3)}
            # move one chunk in +x
            gone, new = pts(list(s1 - s2)), pts(list(s2 - s1))
            t_un, _ = timed(w.unload, gone)
            t_load, _ = timed(w.load, new)
            t_sync, (nd, ne) = timed(w.sync)
            t_mesh, out = timed(w.remesh_dev)
            res["loader_step"] = {"unloaded": len(gone), "loaded": len(new), "unload_ms": t_un, "load_ms": t_load, "sync_ms": t_sync,
                                  "remesh_dev_ms": t_mesh, "dirty": nd, "quads": int(out["counts"].sum())}
        rng = np.random.default_rng(7)
        res["edits"] = []
        w.reserve_remesh(1 << 20)  # the host sizes its pinned mesh buffer once, outside the timed calls (105 MB)
        for n in (1, 64, 4096):
            xyz, mats = surface_edits(n, rng, 32 * (radius // 2))
            t_set, _ = timed(w.set_cubes, xyz, mats)
            t_sync, (nd, ne) = timed(w.sync)
            t_mesh, out = timed(w.remesh)
            res["edits"].append({"edits": n, "set_cubes_ms": t_set, "sync_ms": t_sync, "remesh_host_ms": t_mesh, "dirty": nd,
                                 "entries": ne, "quads": int(out["counts"].sum()),
                                 "edit_to_mesh_ms": t_set + t_sync + t_mesh})
    ctx.close()
    return res


def cpu_run(radius):
    import oracle as O
    w = O.world(0)
    m = O.ChunkMapOracle(w)
    pts = fill_rhombus((0, 0, 0), radius)
    t_load, _ = timed(lambda: [m.load(p) for p in pts])
    t_drain, upd = timed(m.drain)
    entries = sum(v.shape[0] for v in upd.values())
    t_apply, _ = timed(m.apply, upd)
    t_mesh, quads = timed(lambda: sum(m.remesh(p).shape[0] for p in upd))
    rng = np.random.default_rng(7)
    xyz, mats = surface_edits(64, rng, 32 * max(1, radius // 2))
    t_set, _ = timed(lambda: [m.set_cube(p, int(k)) for p, k in zip(xyz, mats)])
    t_d2, upd2 = timed(m.drain)
    m.apply(upd2)
    t_m2, _ = timed(lambda: [m.remesh(p) for p in upd2])
    m.close()
    return {"kind": "port", "cores": 1, "chunks": len(pts), "load_ms": t_load, "drain_ms": t_drain, "apply_ms": t_apply,
            "remesh_ms": t_mesh, "entries": entries, "wire_MB": entries * 12 / 1e6, "quads": int(quads),
            "chunks_per_s_load_sync_remesh": len(pts) / ((t_load + t_drain + t_apply + t_mesh) * 1e-3),
            "edits64": {"set_cube_ms": t_set, "drain_ms": t_d2, "remesh_ms": t_m2, "dirty": len(upd2),
                        "edit_to_mesh_ms": t_set + t_d2 + t_m2}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--radius", type=int, default=16)
    ap.add_argument("--cpu-radius", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    out = {"gpu": gpu_run(a.radius)}
    if not a.no_cpu:
        out["cpu_oracle"] = cpu_run(a.cpu_radius)
    print(json.dumps(out))
