"""The C oracle's mesher half (hbo_generate_mesh: seed hash, colour stream, face order, AO, Instance3d layout) against
tests/np_mesh.py, a second restatement written from the Rust sources alone.  CPU only.

Both restate fastrand's wyrand from its published algorithm; no reference-issued vector for it exists in this image
(see np_mesh.py), so agreement here tightens the pin to "two independent readings of the same sources", not further."""
import numpy as np

import np_mesh as NM
import oracle as O

STONE = [(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0), (0.7, 0.7, 0.7, 1.0)]   # server/src/chunk/material.rs:26-60
DIRT = [(0.4, 0.3, 0.2, 1.0), (0.5, 0.4, 0.3, 1.0), (0.6, 0.5, 0.4, 1.0)]
GRASS = [(0.1, 0.8, 0.1, 1.0), (0.2, 0.9, 0.2, 1.0), (0.3, 1.0, 0.3, 1.0)]
DEFAULT = [STONE, DIRT, GRASS]


def _shells(ids27, pos, flags_override=None):
    """ids27: 27 entries of u16[32768] or None -> (oracle shell of CUBE_DTYPE arrays, np_mesh shell)."""
    so, sn = [], []
    for i, ids in enumerate(ids27):
        if ids is None:
            so.append(None)
            sn.append(None)
            continue
        cubes = O.cubes_from_ids(ids)
        if flags_override is not None and i == 13:
            cubes["flags"] = flags_override
        dx, dy, dz = i // 9 - 1, (i // 3) % 3 - 1, i % 3 - 1
        so.append(cubes)
        sn.append({"position": (pos[0] + dx, pos[1] + dy, pos[2] + dz), "material": cubes["material"].copy(),
                   "flags": cubes["flags"].copy()})
    return so, sn


def _compare(ids27, pos, palette=DEFAULT, flags_override=None):
    so, sn = _shells(ids27, pos, flags_override)
    want = NM.generate_mesh(sn, palette)
    got = O.mesh_cubes_literal(so, pos, O.make_palette(palette))
    assert len(got) == len(want), (len(got), len(want))
    raw = got.tobytes()
    for q, rec in enumerate(want):
        assert raw[q * 100:(q + 1) * 100] == rec, f"quad {q} differs: oracle {got[q]} vs np_mesh {np.frombuffer(rec, dtype=O.INSTANCE_DTYPE)[0]}"
    return len(want)


def _idx(x, y, z):
    return x * 1024 + z * 32 + y


def test_siphash_and_wyrand_restatements_agree():
    import ctypes as C
    for p in [(0, 0, 0), (3, -4, 5), (-262143, 7, 262143), (1, 2, 3)]:
        assert O.lib().hbo_siphash13_chunk(*p) == NM.chunk_seed(p)
    # SipHash paper's reference vector is for SipHash-2-4; for 1-3 the independent check is CPython's own hash
    # (tests/test_oracle_gen_mesh.py::test_siphash13_live_against_cpython pins the oracle; here: np_mesh vs oracle)
    st = C.c_uint64(NM.chunk_seed((3, -4, 5)))
    rng = NM.Rng(NM.chunk_seed((3, -4, 5)))
    for _ in range(64):
        assert np.float32(O.lib().hbo_wyrand_f32(C.byref(st))) == rng.f32()


def test_face_models_and_ao_values():
    for f, name in enumerate(NM.FACES):
        m = np.zeros(12, dtype=np.float32)
        O.lib().hbo_face_matrix(f, m.ctypes.data)
        want = np.zeros((3, 4), dtype=np.float32)
        want[:, :3] = NM.face_model(name)
        assert m.tobytes() == want.tobytes(), name  # bit pattern incl. signed zeros
    sh = [None] * 27
    for k in range(4):
        # occlusion k -> AO factor; facial_ao on an empty shell gives k = 0, so evaluate the formula directly
        a = np.float32(float(k) / 3.0)
        r = np.float32(np.float32(-a) + np.float32(1.0))
        assert np.float32(O.lib().hbo_ao_value(k)) == (r if r > np.float32(0.15) else np.float32(0.15))
    assert NM.facial_ao(sh, "Up", (0, 0, 0)) == [np.float32(1.0)] * 4


def test_sparse_voxels_with_neighbours_and_missing_chunks():
    rng = np.random.default_rng(7)
    ids27 = []
    for i in range(27):
        if i != 13 and i % 4 == 1:
            ids27.append(None)  # unloaded neighbour
            continue
        a = np.zeros(32768, dtype=np.uint16)
        sel = rng.choice(32768, size=500 if i == 13 else 3000, replace=False)
        a[sel] = rng.integers(1, 4, size=len(sel))
        ids27.append(a)
    # make sure border voxels of the centre exist so that AO probes cross into neighbours and into absent chunks
    for (x, y, z) in [(0, 0, 0), (31, 31, 31), (0, 31, 15), (31, 0, 16), (15, 15, 0), (16, 16, 31)]:
        ids27[13][_idx(x, y, z)] = 2
    assert _compare(ids27, (3, -2, 5)) > 2000


def test_caves_all_neighbours_loaded():
    rng = np.random.default_rng(11)
    ids27 = []
    for i in range(27):
        a = np.ones(32768, dtype=np.uint16)
        if i == 13:
            g = a.reshape(32, 32, 32)  # [x][z][y]
            g[4:9, 4:9, 4:9] = 0       # a cave
            g[0:3, 10:14, 29:32] = 0   # a notch on the west / top border
            g[29:32, 0:2, 0:3] = 0     # and on the east / south / bottom corner
            g[12, 12, :] = 0           # a shaft
            g[13, 12, 5] = 3
        else:
            holes = rng.choice(32768, size=2000, replace=False)
            a[holes] = 0
        ids27.append(a)
    assert _compare(ids27, (-1, 0, 2)) > 300


def test_checkerboard_block_and_single_colour_material():
    a = np.zeros(32768, dtype=np.uint16)
    for x in range(10, 18):
        for z in range(10, 18):
            for y in range(10, 18):
                if (x + y + z) & 1:
                    a[_idx(x, y, z)] = 1 + (x % 3)
    ids27 = [None] * 27
    ids27[13] = a
    pal = [STONE, [(0.25, 0.5, 0.75, 1.0)], [(0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0, 0.5)]]
    assert _compare(ids27, (0, 0, 0), pal) == 256 * 6


def test_arbitrary_flag_bits_and_opaque_gate():
    """generate_mesh reads the faces from cube.flags (bit 0 gates, bits 1..6 are the faces, higher bits ignored):
    feed flags the set path would never produce."""
    rng = np.random.default_rng(3)
    a = np.zeros(32768, dtype=np.uint16)
    sel = rng.choice(32768, size=400, replace=False)
    a[sel] = rng.integers(1, 4, size=len(sel))
    flags = rng.integers(0, 1 << 12, size=32768, dtype=np.uint32)
    ids27 = [np.ones(32768, dtype=np.uint16) if i in (4, 22) else None for i in range(27)]
    ids27[13] = a
    assert _compare(ids27, (100, -3, -100), flags_override=flags) > 300
