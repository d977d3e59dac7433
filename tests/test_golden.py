"""The committed golden fixtures (tests/golden/oracle_digests.json, made by tests/golden/make_fixtures.py) against
(1) the oracle on the CPU -- regression anchors for the checker itself -- and (2) the CUDA path on the GPU, directly,
without the oracle in between."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
from make_fixtures import world_script  # noqa: E402

D = json.load(open(os.path.join(GOLD, "oracle_digests.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


# ---------------------------------------------------------------- oracle regression (CPU)
def test_oracle_noise_kind_digests():
    for e in D["noise_kinds"]:
        t = O.noise_tile2(O.world(e["seed"], kind=e["kind"]), *e["col"])
        assert sha(t.tobytes()) == e["sha256"], e


def test_oracle_codec_digests():
    for e in D["codec"]:
        enc = O.rle_encode(O.generate_ids(O.world(e["seed"]), [e["pt"]])[0])
        assert enc.shape[0] == e["bytes"] and sha(enc.tobytes()) == e["sha256"], e


def test_oracle_world_digests():
    for e in D["world"]:
        m = O.ChunkMapOracle(O.world(e["seed"]))
        pts, edits = world_script(e["seed"])
        for p in pts:
            m.load(p)
        m.apply(m.drain())
        for pos, mat in edits:
            m.set_cube(pos, mat)
        upd = m.drain()
        m.apply(upd)
        assert sha(b"".join(m.cubes(p).tobytes() for p in pts)) == e["cubes_sha256"]
        assert [list(p) for p in sorted(upd)] == e["dirty"]
        h = hashlib.sha256()
        for p in sorted(upd):
            h.update(m.remesh(p).tobytes())
        assert h.hexdigest() == e["mesh_sha256"]
        m.close()


# ---------------------------------------------------------------- CUDA path against the fixtures (GPU)
@pytest.mark.gpu
def test_gpu_noise_ids_mesh_match_golden(ctx):
    import herbolution_b200 as hb
    try:
        for e in D["noise"]:
            ctx.set_world(e["seed"])
            noise = ctx.noise_tiles([e["col"]])[0]
            assert sha(noise.tobytes()) == e["sha256"], e
        for e in D["ids"]:
            ctx.set_world(e["seed"])
            assert sha(ctx.generate([e["pt"]])["ids"][0].tobytes()) == e["sha256"], e
        for e in D["mesh"]:
            ctx.set_world(e["seed"])
            parts = []
            ctx.generate_mesh_region(hb.Region(*e["region"]), sink=lambda first, off, cnt, inst, ids: parts.append(
                (first, cnt.copy(), [inst[int(o):int(o) + int(n)].copy() for o, n in zip(off, cnt)])))
            parts.sort(key=lambda t: t[0])
            counts = np.concatenate([p[1] for p in parts])
            inst = np.concatenate([c for p in parts for c in p[2]])
            assert int(counts.sum()) == e["quads"]
            assert sha(counts.astype(np.uint32).tobytes()) == e["counts_sha256"]
            assert sha(inst["model"].tobytes() + inst["ao"].tobytes()) == e["geometry_sha256"]
            assert sha(inst["rgba"].tobytes()) == e["colour_sha256"]
        for e in D["noise_kinds"]:
            ctx.set_world(e["seed"])
            ctx.set_noise_kind(e["kind"])
            assert sha(ctx.noise_tiles([e["col"]])[0].tobytes()) == e["sha256"], e
    finally:
        ctx.set_noise_kind(O.NOISE_FBM)
        ctx.set_world(0)


@pytest.mark.gpu
def test_gpu_codec_matches_golden(ctx):
    for e in D["codec"]:
        ctx.set_world(e["seed"])
        enc = ctx.rle_encode(ctx.generate([e["pt"]])["ids"])
        assert int(enc["sizes"][0]) == e["bytes"]
        assert sha(enc["bytes"][:e["bytes"]].tobytes()) == e["sha256"], e
    ctx.set_world(0)


@pytest.mark.gpu
def test_gpu_world_matches_golden(ctx):
    import herbolution_b200 as hb
    for e in D["world"]:
        ctx.set_world(e["seed"])
        pts, edits = world_script(e["seed"])
        with hb.World(ctx, 16) as world:
            world.load(pts)
            world.sync()
            world.remesh()
            world.set_cubes([p for p, _ in edits], [m for _, m in edits])
            assert sha(b"".join(world.read_chunk(p).tobytes() for p in pts)) == e["cubes_sha256"]
            n_dirty, n_entries = world.sync()
            assert n_entries == e["update_positions"]
            out = world.remesh()
            got = {(int(p["x"]), int(p["y"]), int(p["z"])): out["instances"][int(o):int(o) + int(n)]
                   for p, o, n in zip(out["pts"], out["offsets"], out["counts"])}
            assert [list(p) for p in sorted(got)] == e["dirty"]
            h = hashlib.sha256()
            for p in sorted(got):
                h.update(got[p].tobytes())
            assert sum(v.shape[0] for v in got.values()) == e["quads"] and h.hexdigest() == e["mesh_sha256"]
    ctx.set_world(0)
