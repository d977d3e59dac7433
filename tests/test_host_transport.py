"""The host half of the packed transports, on the CPU (no GPU, no context): hb_host_fill_ids and hb_host_expand_quads are
product code (csrc/expand.cu) that rebuilds block ids from heights tiles and Instance3d records from 4-byte quads.
Checked against the oracle: ids vs generate(), records vs generate_mesh() byte for byte."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from herbolution_b200 import _lib

STONE = [(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0), (0.7, 0.7, 0.7, 1.0)]


def aligned(n, dtype):
    raw = np.zeros(n * np.dtype(dtype).itemsize + 64, dtype=np.uint8)
    o = (-raw.ctypes.data) & 63
    return raw[o:o + n * np.dtype(dtype).itemsize].view(dtype)


@pytest.mark.parametrize("seed", [0, 5])
def test_host_fill_ids_equals_generate(seed):
    L = _lib.load()
    w = O.world(seed)
    for (cx, cz) in [(0, 0), (-3, 7), (100, -100)]:
        noise = O.noise_tile2(w, cx, cz)                       # index x + z*32
        h = O.heights_from_noise(noise).reshape(32, 32).T       # -> [x][z]
        heights = np.ascontiguousarray(h, dtype=np.int32).reshape(-1)
        top = int(heights.max()) // 32
        for cy in (top - 2, top - 1, top, top + 1, -40, 40):
            out = aligned(32768, np.uint16)
            assert L.hb_host_fill_ids(C.c_void_p(heights.ctypes.data), cy, C.c_void_p(out.ctypes.data)) == 0
            ref = O.generate_ids(w, [(cx, cy, cz)])[0]
            assert (out == ref).all(), (seed, cx, cy, cz)
    # extreme heights do not overflow the layer arithmetic
    heights = np.full(1024, 2**31 - 1, dtype=np.int32)
    out = aligned(32768, np.uint16)
    assert L.hb_host_fill_ids(C.c_void_p(heights.ctypes.data), -67108864, C.c_void_p(out.ctypes.data)) == 0 and (out == 1).all()
    heights[:] = -2**31
    assert L.hb_host_fill_ids(C.c_void_p(heights.ctypes.data), 67108863, C.c_void_p(out.ctypes.data)) == 0 and (out == 0).all()
    assert L.hb_host_fill_ids(None, 0, C.c_void_p(out.ctypes.data)) == _lib.HB_ERR_INVALID


def pack_from_instances(inst, ids, pos):
    """Derive the 4-byte quads (L | face << 15 | ao << 18 | (material - 1) << 26) from an Instance3d stream."""
    mats = np.zeros((6, 12), dtype=np.float32)
    for f in range(6):
        O.lib().hbo_face_matrix(f, mats[f].ctypes.data)
    ao_vals = [np.float32(O.lib().hbo_ao_value(k)) for k in range(4)]
    out = np.zeros(len(inst), dtype=np.uint32)
    for q, rec in enumerate(inst):
        model = rec["model"]
        face = next(f for f in range(6) if model[:12].tobytes() == mats[f].tobytes())
        x, y, z = (int(model[12 + k]) - 32 * int(pos[k]) for k in range(3))
        L = x * 1024 + z * 32 + y
        ao = 0
        for v in range(4):
            ao |= ao_vals.index(np.float32(rec["ao"][v])) << (2 * v)
        out[q] = L | (face << 15) | (ao << 18) | ((int(ids[L]) - 1) << 26)
    return out


@pytest.mark.parametrize("palette", [None, [STONE, [(0.25, 0.5, 0.75, 1.0)], [(0.1 * k, 0.2, 0.3, 1.0) for k in range(8)]]])
def test_host_expand_quads_equals_generate_mesh(palette):
    L = _lib.load()
    rng = np.random.default_rng(4)
    pos = (5, -2, -7)
    shell = []
    for i in range(27):
        a = np.zeros(32768, dtype=np.uint16)
        sel = rng.choice(32768, size=900 if i == 13 else 6000, replace=False)
        a[sel] = rng.integers(1, 4, size=len(sel))
        shell.append(None if i in (1, 25) else a)
    pal = O.make_palette(palette) if palette else None
    want = O.mesh_ids(shell, pos, pal)
    assert len(want) > 3000
    quads = pack_from_instances(want, shell[13], pos)
    n = len(quads)
    out = aligned(((n + 3) & ~3), O.INSTANCE_DTYPE)
    if palette:
        ncol = np.array([len(c) for c in palette], dtype=np.int32)
        rgba = np.zeros((len(palette), 8, 4), dtype=np.float32)
        for i, cs in enumerate(palette):
            rgba[i, :len(cs)] = cs
        rc = L.hb_host_expand_quads(_lib.ChunkPtC(*pos), C.c_void_p(quads.ctypes.data), n, len(palette), C.c_void_p(ncol.ctypes.data),
                                    C.c_void_p(rgba.ctypes.data), C.c_void_p(out.ctypes.data))
    else:
        rc = L.hb_host_expand_quads(_lib.ChunkPtC(*pos), C.c_void_p(quads.ctypes.data), n, 0, None, None, C.c_void_p(out.ctypes.data))
    assert rc == 0
    assert out[:n].tobytes() == want.tobytes(), "expanded records differ from the oracle's generate_mesh"
    assert not out[n:].tobytes().strip(b"\0"), "gap after the last quad must be zero"
    assert L.hb_host_expand_quads(_lib.ChunkPtC(*pos), None, 4, 0, None, None, C.c_void_p(out.ctypes.data)) == _lib.HB_ERR_INVALID


def test_edit_conflict_groups_are_the_stencil_components():
    """hb_host_group_edits (the grouping hb_world_set_cubes uses): two edits belong to one group iff they are connected
    through overlapping 7-voxel stencils, i.e. through pairs at L1 distance <= 2; groups are numbered by their first
    edit.  Checked against a brute-force graph (scipy connected components) on scattered, clustered, duplicated and
    chained edits."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    L = _lib.load()
    rng = np.random.default_rng(12)
    cases = [rng.integers(-40, 40, size=(600, 3)),                                   # sparse: mostly singletons
             rng.integers(0, 9, size=(500, 3)),                                      # dense: few big components
             np.concatenate([np.stack([np.arange(50) * 2, np.zeros(50), np.zeros(50)], 1),  # a chain at distance 2
                             np.array([[200, 0, 0], [203, 0, 0], [200, 0, 0]]),             # distance 3 apart, one repeat
                             np.array([[31, 5, 5], [32, 5, 5], [32, 6, 6]])]),              # across a chunk border, diagonal
             np.array([[2**31 - 3, 0, -2**31 + 2]])]                                 # near the int32 ends
    for xyz in cases:
        xyz = np.ascontiguousarray(xyz, dtype=np.int32)
        n = len(xyz)
        got = np.zeros(n, dtype=np.uint32)
        ng = C.c_uint32(0)
        assert L.hb_host_group_edits(C.c_void_p(xyz.ctypes.data), n, C.c_void_p(got.ctypes.data), C.byref(ng)) == 0
        d = np.abs(xyz[:, None, :].astype(np.int64) - xyz[None, :, :].astype(np.int64)).sum(-1)
        i, j = np.nonzero(d <= 2)
        ncomp, lab = connected_components(coo_matrix((np.ones(len(i)), (i, j)), shape=(n, n)), directed=False)
        assert ng.value == ncomp
        # same partition, and groups numbered in order of first appearance
        seen = {}
        for k in range(n):
            seen.setdefault(int(lab[k]), len(seen))
            assert got[k] == seen[int(lab[k])], (k, got[k], lab[k])
    assert L.hb_host_group_edits(None, 0, None, C.byref(ng)) == 0 and ng.value == 0


def test_host_fill_ids_against_the_threshold_rule_in_numpy():
    """generator/mod.rs:85-91 written out in numpy (independent of the C oracle): with d = h - world_y, stone (1) for
    d > 6, dirt (2) for d > 1, grass (3) for d > 0, else air."""
    L = _lib.load()
    rng = np.random.default_rng(21)
    for cy in (-3, 0, 2):
        heights = rng.integers(cy * 32 - 12, cy * 32 + 44, size=1024).astype(np.int32)  # straddles the chunk and its thresholds
        out = aligned(32768, np.uint16)
        assert L.hb_host_fill_ids(C.c_void_p(heights.ctypes.data), cy, C.c_void_p(out.ctypes.data)) == 0
        d = heights.astype(np.int64)[:, None] - (cy * 32 + np.arange(32))[None, :]        # [column x*32+z][y]
        want = np.where(d > 6, 1, np.where(d > 1, 2, np.where(d > 0, 3, 0))).astype(np.uint16)
        assert (out.reshape(1024, 32) == want).all()
