"""CPU checks of the oracle's server ChunkMap / CubeUpdate / client restatement (oracle/hb_oracle.c hbo_map_*):
the literal incremental path must agree with the closed forms the GPU kernels implement."""
import numpy as np

import oracle as O

NORMAL = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]


def closed_form_cubes(w, pos, loaded):
    """PaletteCube image of `pos` after generation + cull_shared_faces with the loaded face neighbours:
    in-chunk closed form (hbo_cubes_from_ids), then a border face goes away iff both sides are solid."""
    ids = O.generate_ids(w, [pos])[0]
    cubes = O.cubes_from_ids(ids)
    flags = cubes["flags"].reshape(32, 32, 32).copy()  # [x][z][y]
    solid = (ids != 0).reshape(32, 32, 32)
    for f, d in enumerate(NORMAL):
        q = (pos[0] + d[0], pos[1] + d[1], pos[2] + d[2])
        if q not in loaded:
            continue
        other = (O.generate_ids(w, [q])[0] != 0).reshape(32, 32, 32)
        axis = {0: 0, 1: 2, 2: 1}[f >> 1]  # array axis of x / y / z in [x][z][y]
        mine = [slice(None)] * 3
        theirs = [slice(None)] * 3
        mine[axis] = 31 if f % 2 == 0 else 0
        theirs[axis] = 0 if f % 2 == 0 else 31
        both = solid[tuple(mine)] & other[tuple(theirs)]
        sl = flags[tuple(mine)]
        sl[both] &= ~np.uint32(1 << (f + 1))
        flags[tuple(mine)] = sl
    cubes["flags"] = flags.reshape(-1)
    return cubes


def test_load_equals_closed_form_any_order():
    w = O.world(5)
    pts = [(x, y, z) for x in (3, 4) for y in (-1, 0) for z in (-2, -1)]
    rng = np.random.default_rng(1)
    for _ in range(2):
        rng.shuffle(pts)
        m = O.ChunkMapOracle(w)
        for p in pts:
            m.load(p)
        for p in pts:
            assert m.cubes(p).tobytes() == closed_form_cubes(w, p, set(pts)).tobytes(), p
        m.close()


def test_generation_update_set_is_solid_plus_neighbours():
    """updated_positions after generate() as a set = solid voxels and their in-chunk 6-neighbours (what the fill
    kernel writes to the touched bitmap)."""
    w = O.world(0)
    m = O.ChunkMapOracle(w)
    m.load((0, 0, 0))
    upd = m.drain()[(0, 0, 0)]
    solid = (O.generate_ids(w, [(0, 0, 0)])[0] != 0).reshape(32, 32, 32)
    want = solid.copy()
    for ax in range(3):
        for sh in (1, -1):
            r = np.roll(solid, sh, axis=ax)
            idx = [slice(None)] * 3
            idx[ax] = 0 if sh == 1 else 31
            r[tuple(idx)] = False
            want |= r
    x, y, z = (upd["pos"] >> 11) & 31, (upd["pos"] >> 6) & 31, (upd["pos"] >> 1) & 31
    got = np.zeros((32, 32, 32), dtype=bool)
    got[x, z, y] = True
    assert (got == want).all()
    assert upd.shape[0] > want.sum()  # the Vec keeps duplicates
    m.close()


def test_duplicates_are_invisible_to_the_client():
    w = O.world(0)
    m = O.ChunkMapOracle(w)
    for p in [(0, -1, 0), (0, 0, 0), (1, 0, 0)]:
        m.load(p)
    m.set_cube((31, 0, 0), 0)
    m.set_cube((31, 0, 0), 2)
    for pos, entries in m.drain().items():
        a = np.zeros(O.CHUNK_VOLUME, dtype=O.CUBE_DTYPE)
        O.lib().hbo_apply_updates(a.ctypes.data, np.ascontiguousarray(entries).ctypes.data, entries.shape[0])
        _, first = np.unique(entries["pos"], return_index=True)
        dedup = np.ascontiguousarray(entries[np.sort(first)])
        b = np.zeros(O.CHUNK_VOLUME, dtype=O.CUBE_DTYPE)
        O.lib().hbo_apply_updates(b.ctypes.data, dedup.ctypes.data, dedup.shape[0])
        assert a.tobytes() == b.tobytes() == m.cubes(pos).tobytes()
    m.close()


def test_set_cube_dig_place_roundtrip_and_border_insert():
    w = O.world(0)
    m = O.ChunkMapOracle(w)
    m.load((0, -1, 0))
    m.load((1, -1, 0))
    before = m.cubes((0, -1, 0))
    L = 5 * 1024 + 7 * 32 + 9
    assert before["material"][L] == 1 and before["flags"][L] == 1  # buried stone: no faces
    m.set_cube((5, -32 + 9, 7), 0)
    dug = m.cubes((0, -1, 0))
    assert dug["material"][L] == 0 and dug["flags"][L] == 1
    # add_neighboring_faces walks faces.not(): all six neighbours gain the face towards the hole
    for f, d in enumerate(NORMAL):
        Ln = (5 + d[0]) * 1024 + (7 + d[2]) * 32 + (9 + d[1])
        assert (dug["flags"][Ln] >> 1) & 63 == 1 << (f ^ 1)
    m.set_cube((5, -32 + 9, 7), 1)
    assert m.cubes((0, -1, 0)).tobytes() == before.tobytes()
    # border dig: the neighbour chunk's voxel gains the face unconditionally (map.rs:113)
    m.set_cube((31, -20, 4), 0)
    nb = m.cubes((1, -1, 0))
    Lb = 0 * 1024 + 4 * 32 + 12
    assert (nb["flags"][Lb] >> 1) & 63 == 0b10  # WEST
    m.close()


def test_unload_keeps_neighbour_faces_culled():
    w = O.world(0)
    m = O.ChunkMapOracle(w)
    m.load((0, -1, 0))
    alone = m.cubes((0, -1, 0))
    m.load((1, -1, 0))
    both = m.cubes((0, -1, 0))
    assert alone.tobytes() != both.tobytes()
    m.unload((1, -1, 0))
    assert m.cubes((0, -1, 0)).tobytes() == both.tobytes()
    m.close()
