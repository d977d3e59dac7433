"""GPU parity for the device-resident world (hb_world_*) against the literal
restatement of the reference's server ChunkMap + CubeUpdate wire + client remesh (tests/oracle.py
ChunkMapOracle over oracle/hb_oracle.c hbo_map_*).

Bar: PaletteCube arrays bit-exact after every load / edit batch; the CubeUpdate entries equal the
reference's updated_positions as a SET with the same final cubes (the reference's Vec keeps ~7x duplicates
that apply_updates_from_server cannot tell apart -- asserted here by applying both); the remesh set equals
the set of chunks that received a CubeUpdate; every remeshed chunk's Instance3d stream bit-exact, in order.
"""
import numpy as np
import pytest

import herbolution_b200 as hb
import oracle as O
from test_gpu_parity import assert_instances_equal

pytestmark = pytest.mark.gpu


def key(p):
    return (int(p["x"]), int(p["y"]), int(p["z"]))


def split_updates(upd):
    return {key(p): upd["entries"][int(o):int(o) + int(n)] for p, o, n in zip(upd["pts"], upd["offsets"], upd["counts"])}


def split_meshes(out):
    return {key(p): out["instances"][int(o):int(o) + int(n)] for p, o, n in zip(out["pts"], out["offsets"], out["counts"])}


def check_sync(world, ora):
    """One server tick's sync_with_client on both sides.  Returns the set of chunks that received a CubeUpdate."""
    want = ora.drain()
    n_dirty, n_entries = world.sync()
    got = split_updates(world.fetch_updates())
    assert set(got) == set(want) and n_dirty == len(want)
    total = 0
    for pos, w in want.items():
        g = got[pos]
        total += g.shape[0]
        # same positions as a set, ascending voxel index (x, z, y), no duplicates
        x, y, z = hb.world.vec3u5_unpack(g["pos"])
        L = x * 1024 + z * 32 + y
        assert (np.diff(L) > 0).all(), pos
        assert set(g["pos"].tolist()) == set(w["pos"].tolist()), pos
        assert (g["pad0"] == 0).all()
        # every entry carries the cube as it is now; the reference's duplicates carry the same cube
        now = ora.cubes(pos)
        assert (g["material"] == now["material"][L]).all() and (g["flags"] == now["flags"][L]).all() and (g["pad"] == 0).all()
        # the client cannot tell the two encodings apart
        a = ora.client[pos].copy()
        e = np.ascontiguousarray(g)
        O.lib().hbo_apply_updates(a.ctypes.data, e.ctypes.data, e.shape[0])
        b = ora.client[pos].copy()
        O.lib().hbo_apply_updates(b.ctypes.data, np.ascontiguousarray(w).ctypes.data, w.shape[0])
        assert a.tobytes() == b.tobytes(), pos
    assert total == n_entries
    ora.apply(want)
    return set(want)


def check_cubes(world, ora):
    for _, pos in ora.loaded():
        got = world.read_chunk(pos)
        want = ora.cubes(pos)
        assert got.tobytes() == want.tobytes(), f"chunk {pos}: {(got != want).sum()} cubes differ"


def check_remesh(world, ora, expect):
    got = split_meshes(world.remesh())
    loaded = {pos for _, pos in ora.loaded()}
    assert set(got) == (expect & loaded)
    for pos, inst in got.items():
        assert_instances_equal(inst, ora.remesh(pos), f"chunk {pos}")
    return got


def surface_block(cx0, cz0, n, ys=(-1, 0)):
    return [(cx0 + i, y, cz0 + j) for i in range(n) for j in range(n) for y in ys]


@pytest.mark.parametrize("seed", [0, 5])
def test_world_load_sync_remesh(ctx, seed):
    """Bulk load in two batches (old/new and new/new cull pairs), then the full server->client->mesh loop."""
    ctx.set_world(seed)
    ora = O.ChunkMapOracle(O.world(seed))
    pts = surface_block(-2, 3, 3)
    rng = np.random.default_rng(seed)
    rng.shuffle(pts)
    with hb.World(ctx, 64) as world:
        for batch in (pts[:7], pts[7:] + pts[:3]):  # second batch repeats loaded chunks: ignored (map.rs:69-75)
            world.load(batch)
            for p in batch:
                ora.load(p)
        assert len(world) == len(pts)
        check_cubes(world, ora)
        dirty = check_sync(world, ora)
        assert dirty == set(pts)
        check_remesh(world, ora, dirty)
        # nothing pending any more
        assert world.sync() == (0, 0)
        assert world.remesh()["pts"].shape[0] == 0
    ora.close()
    ctx.set_world(0)


def test_world_incremental_load_remeshes_neighbours(ctx):
    """A chunk loaded next to existing ones culls their border faces: they get CubeUpdates and are remeshed."""
    ora = O.ChunkMapOracle(O.world(0))
    with hb.World(ctx, 16) as world:
        first = [(0, -1, 0), (0, 0, 0)]
        world.load(first)
        for p in first:
            ora.load(p)
        check_remesh(world, ora, check_sync(world, ora))
        second = [(1, -1, 0), (0, -2, 0)]
        world.load(second)
        for p in second:
            ora.load(p)
        check_cubes(world, ora)
        dirty = check_sync(world, ora)
        assert (0, -1, 0) in dirty  # solid chunk below the surface: both new neighbours cull against it
        check_remesh(world, ora, dirty)
    ora.close()


def edit_script(rng, w, n):
    """Digs and places around the surface of chunks (0..1, -1..0, 0..1), biased to chunk borders and repeats."""
    edits = []
    hcache = {}
    for _ in range(n):
        wx, wz = int(rng.integers(-3, 67)), int(rng.integers(-3, 67))
        if rng.random() < 0.4:  # force a chunk border column
            wx = int(rng.choice([-1, 0, 31, 32, 63, 64]))
        if rng.random() < 0.4:
            wz = int(rng.choice([-1, 0, 31, 32, 63, 64]))
        cx, cz = wx >> 5, wz >> 5
        if (cx, cz) not in hcache:
            hcache[(cx, cz)] = O.heights_from_noise(O.noise_tile2(w, cx, cz)).reshape(32, 32)  # [z][x]
        h = int(hcache[(cx, cz)][wz & 31, wx & 31])
        wy = h + int(rng.integers(-4, 3))
        if rng.random() < 0.15:
            wy = int(rng.choice([-33, -32, -1, 0, 31, 32]))
        mat = int(rng.choice([0, 0, 0, 1, 2, 3]))
        edits.append(((wx, wy, wz), mat))
        if rng.random() < 0.3:  # touch the same or an adjacent voxel again
            d = [0, 0, 0]
            d[int(rng.integers(0, 3))] = int(rng.integers(-1, 2))
            edits.append(((wx + d[0], wy + d[1], wz + d[2]), int(rng.choice([0, 1, 3]))))
    return edits


@pytest.mark.parametrize("seed,batch", [(0, 1), (0, 37), (5, 400)])
def test_world_edits(ctx, seed, batch):
    """ChunkMap::set_cube in order (dig, place, replace, no-op, borders incl. unloaded sides), batch by batch."""
    ctx.set_world(seed)
    w = O.world(seed)
    ora = O.ChunkMapOracle(w)
    # chunk (1, *, 1) column missing on purpose: edits at x in [32,64), z in [32,64) hit an unloaded chunk whose
    # loaded neighbours still get the border insert (map.rs:85-140)
    pts = [(cx, cy, cz) for cx in (0, 1) for cz in (0, 1) for cy in (-2, -1, 0, 1) if not (cx == 1 and cz == 1)]
    rng = np.random.default_rng(100 + seed)
    with hb.World(ctx, 32) as world:
        world.load(pts)
        for p in pts:
            ora.load(p)
        check_remesh(world, ora, check_sync(world, ora))
        edits = edit_script(rng, w, 1200 if batch > 1 else 60)
        for i in range(0, len(edits), batch):
            chunk = edits[i:i + batch]
            world.set_cubes([e[0] for e in chunk], [e[1] for e in chunk])
            for pos, mat in chunk:
                ora.set_cube(pos, mat)
            if batch == 1 or (i // batch) % 3 == 0:
                check_cubes(world, ora)
                check_remesh(world, ora, check_sync(world, ora))
        check_cubes(world, ora)
        check_remesh(world, ora, check_sync(world, ora))
    ora.close()
    ctx.set_world(0)


def test_world_border_quirk_and_unloaded_owner(ctx):
    """Digging a border voxel inserts the facing face on the neighbour chunk's voxel even when that voxel is air
    (flags gain faces on an air cube), and an edit whose own chunk is not loaded still does so."""
    ora = O.ChunkMapOracle(O.world(0))
    with hb.World(ctx, 8) as world:
        pts = [(0, 2, 0), (1, 2, 0)]  # all-air layers high above the terrain
        world.load(pts)
        for p in pts:
            ora.load(p)
        check_sync(world, ora)
        world.remesh()
        edits = [((31, 70, 5), 1), ((31, 70, 5), 0), ((32, 70, 5), 2), ((-1, 70, 5), 3), ((64, 70, 9), 0)]
        for pos, mat in edits:
            world.set_cube(pos, mat)
            ora.set_cube(pos, mat)
        check_cubes(world, ora)
        got = world.read_chunk((1, 2, 0))
        L = 0 * 1024 + 5 * 32 + (70 & 31)
        assert got["material"][L] == 2 and got["flags"][L] & 1
        dirty = check_sync(world, ora)
        assert dirty == {(0, 2, 0), (1, 2, 0)}
        check_remesh(world, ora, dirty)
    ora.close()


def test_world_unload_reload(ctx):
    ora = O.ChunkMapOracle(O.world(0))
    with hb.World(ctx, 4) as world:
        pts = [(0, -1, 0), (1, -1, 0), (0, -1, 1)]
        world.load(pts)
        for p in pts:
            ora.load(p)
        check_remesh(world, ora, check_sync(world, ora))
        world.set_cube((31, -20, 3), 0)
        ora.set_cube((31, -20, 3), 0)
        world.unload([(1, -1, 0), (9, 9, 9)])  # the second is not loaded: ignored
        ora.unload((1, -1, 0))
        assert len(world) == 2
        with pytest.raises(hb.HerbGpuError):
            world.read_chunk((1, -1, 0))
        check_cubes(world, ora)  # neighbours keep the culled faces (map.rs:230-235 does not restore them)
        dirty = check_sync(world, ora)
        assert dirty == {(0, -1, 0)}
        check_remesh(world, ora, dirty)
        # the freed slot is reused; the reload culls against the survivors again
        world.load([(1, -1, 0), (2, -1, 0)])
        ora.load((1, -1, 0))
        ora.load((2, -1, 0))
        check_cubes(world, ora)
        check_remesh(world, ora, check_sync(world, ora))
        with pytest.raises(hb.HerbGpuError) as e:
            world.load([(5, 5, 5)])
        assert e.value.code == hb._lib.HB_ERR_CAPACITY and len(world) == 4
    ora.close()


def test_world_errors_and_two_call_remesh(ctx):
    import ctypes as C
    lib = hb._lib.load()
    with pytest.raises(hb.HerbGpuError):
        hb.World(ctx, 0)
    with hb.World(ctx, 4) as world:
        world.load(np.zeros((0, 3), dtype=np.int64))
        assert world.sync() == (0, 0)
        with pytest.raises(hb.HerbGpuError) as e:
            world.load([(0, 0, 2**20)])
        assert e.value.code == hb._lib.HB_ERR_RANGE
        world.load([(0, -2, 0)])
        with pytest.raises(hb.HerbGpuError) as e:
            world.set_cubes([(0, -40, 0)], [9])
        assert e.value.code == hb._lib.HB_ERR_INVALID
        n_dirty, n_entries = world.sync()
        assert n_dirty == 1 and n_entries == 32768  # a solid chunk: every voxel was set
        # fetch with too little room reports HB_ERR_CAPACITY and can be retried
        pts = np.zeros(1, dtype=hb.CHUNK_PT_DTYPE)
        off, cnt = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint32)
        small = np.zeros(10, dtype=hb.CHUNK_CUBE_DTYPE)
        nr, ne = C.c_size_t(0), C.c_uint64(0)
        assert lib.hb_world_fetch_updates(world._h, hb._lib.ptr(pts), 1, hb._lib.ptr(off), hb._lib.ptr(cnt),
                                          hb._lib.ptr(small), 10, C.byref(nr), C.byref(ne)) == hb._lib.HB_ERR_CAPACITY
        assert nr.value == 1 and ne.value == 32768
        assert world.fetch_updates()["entries"].shape[0] == 32768
        # remesh: sizing call keeps the remesh set
        n, total = C.c_size_t(0), C.c_uint64(0)
        assert lib.hb_world_remesh(world._h, None, 0, None, 0, None, None, C.byref(n), C.byref(total)) == hb._lib.HB_ERR_CAPACITY
        assert n.value == 1 and total.value == 6 * 1024  # lone solid chunk: six full border slabs
        out = world.remesh()
        assert out["counts"].tolist() == [6 * 1024]
        assert world.remesh()["pts"].shape[0] == 0


def test_world_remesh_dev_matches_host(ctx):
    pts = surface_block(10, -4, 2)
    with hb.World(ctx, 16) as a, hb.World(ctx, 16) as b:
        a.load(pts)
        b.load(pts)
        a.sync()
        b.sync()
        host = a.remesh()
        dev = b.remesh_dev()
        assert dev["span"] == host["instances"].shape[0]
        assert (dev["offsets"] == host["offsets"]).all() and (dev["counts"] == host["counts"]).all()
        inst = ctx.download(dev["d_instances"], hb.INSTANCE_DTYPE, dev["span"])
        assert inst.tobytes() == host["instances"].tobytes()


def test_server_chunk_map_mirror_and_extreme_coordinates(ctx):
    """The reference-named mirror (queue_load / update / set_cube / sync_with_client / remesh) at negative chunk
    coordinates and at the +-262143 coordinate limit."""
    ora = O.ChunkMapOracle(O.world(0))
    m = hb.ServerChunkMap(ctx, 16)
    lim = 262143
    pts = [(-5, -1, -9), (-4, -1, -9), (-5, 0, -9), (lim, 0, -lim), (lim, -1, -lim), (lim - 1, 0, -lim)]
    for p in pts:
        m.queue_load(p)
        ora.load(p)
    m.update()
    assert len(m) == len(pts)
    edits = [((-5 * 32 + 31, -3, -9 * 32), 0), ((-4 * 32, -3, -9 * 32), 0), ((-5 * 32 + 31, -3, -9 * 32), 3),
             ((lim * 32 + 31, 5, -lim * 32), 2), ((lim * 32, 5, -lim * 32 + 31), 1), ((lim * 32, -1, -lim * 32 + 7), 0),
             ((lim * 32, 0, -lim * 32 + 7), 0)]
    for pos, mat in edits:
        m.set_cube(pos, mat)
        ora.set_cube(pos, mat)
    want = ora.drain()
    got = m.sync_with_client()
    assert set(got) == set(want)
    for pos in want:
        assert set(got[pos]["pos"].tolist()) == set(want[pos]["pos"].tolist())
        assert m.world.read_chunk(pos).tobytes() == ora.cubes(pos).tobytes()
    ora.apply(want)
    meshes = m.remesh()
    assert set(meshes) == set(want)
    for pos, inst in meshes.items():
        assert_instances_equal(inst, ora.remesh(pos), f"chunk {pos}")
    m.queue_unload(pts[0])
    m.update()
    assert len(m) == len(pts) - 1
    with pytest.raises(hb.HerbGpuError) as e:
        m.queue_load((lim + 1, 0, 0))
        m.update()
    assert e.value.code == hb._lib.HB_ERR_RANGE
    m.close()
    ora.close()


def test_readme_example_runs(ctx):
    """The usage snippet of README.md, kept honest."""
    gen = ctx.generate([(0, 0, 0), (0, -1, 0)], want_cubes=True)
    mesh = ctx.mesh([(0, 0, 0), (0, -1, 0)], gen["ids"])
    assert mesh["counts"].sum() > 0 and gen["cubes"].shape[0] == 2
    tot = ctx.generate_mesh_region(hb.Region(-8, -8, 16, 16, -3, 6), sink=lambda first, off, cnt, inst, ids: None)
    assert tot["chunks"] == 16 * 16 * 6 and tot["quads"] > 0
    world = hb.ServerChunkMap(ctx, capacity=64)
    world.queue_load((0, 0, 0))
    world.update()
    world.set_cube((3, 5, 7), 0)
    updates = world.sync_with_client()
    meshes = world.remesh()
    assert set(updates) == set(meshes) == {(0, 0, 0)}
    world.close()


def test_fetch_updates_after_unload_reports_what_it_wrote(ctx):
    """sync -> unload -> fetch: runs of chunks unloaded since the sync are skipped, and the call says how many runs
    and entries it actually wrote (no zero-padded (0,0,0) run can reach ServerChunkMap.sync_with_client)."""
    import ctypes as C
    ctx.set_world(0)
    lib = hb._lib.load()
    with hb.World(ctx, 8) as world:
        pts = [(0, -2, 0), (1, -2, 0), (2, -2, 0)]
        world.load(pts)
        n_dirty, n_entries = world.sync()
        assert n_dirty == 3
        world.unload([pts[1]])
        upd = world.fetch_updates()
        assert [key(p) for p in upd["pts"]] == [pts[0], pts[2]]
        assert upd["offsets"].shape[0] == 2 and upd["counts"].shape[0] == 2
        assert upd["entries"].shape[0] == int(upd["counts"].sum()) < n_entries
        # nothing left: the outputs report zero runs instead of leaving the caller's arrays untouched
        world.unload([pts[0], pts[2]])
        nr, ne = C.c_size_t(99), C.c_uint64(99)
        assert lib.hb_world_fetch_updates(world._h, None, 0, None, None, None, 0, C.byref(nr), C.byref(ne)) == hb._lib.HB_OK
        assert nr.value == 0 and ne.value == 0
        assert lib.hb_world_fetch_updates(world._h, None, 0, None, None, None, 0, None, None) == hb._lib.HB_ERR_INVALID


def test_world_edit_conflict_groups(ctx):
    """One batch of 4 000 edits is split into conflict groups (edits whose 7-voxel stencils overlap stay in order on one
    warp, the groups run concurrently): chains of adjacent digs, repeated edits of one voxel, diagonal neighbours,
    chunk-border pairs and scattered edits -- the cubes, the CubeUpdate sets and the remeshed chunks must equal the
    reference's strictly sequential ChunkMap::set_cube calls (server/src/chunk/map.rs:81-151)."""
    seed = 5
    ctx.set_world(seed)
    w = O.world(seed)
    ora = O.ChunkMapOracle(w)
    pts = [(cx, cy, cz) for cx in (0, 1) for cz in (0, 1) for cy in (-2, -1, 0, 1)]
    rng = np.random.default_rng(77)
    edits = edit_script(rng, w, 2400)
    # a tunnel: 200 consecutive voxels dug in order, then refilled in reverse (one long group)
    tunnel = [((5 + i % 50, -20 - i // 50, 9), 0) for i in range(200)]
    edits += tunnel + [(p, 2) for p, _ in reversed(tunnel)]
    # the same voxel toggled many times, interleaved with far-away edits
    for i in range(150):
        edits.append(((40, -12, 40), i % 4))
        edits.append(((int(rng.integers(0, 64)), int(rng.integers(-60, -30)), int(rng.integers(0, 64))), int(rng.integers(0, 4))))
    # pairs across chunk borders and diagonal neighbours (L1 distance 2: they share a stencil voxel)
    for i in range(100):
        y = -40 + i % 20
        edits += [((31, y, 3 + i % 25), 0), ((32, y, 3 + i % 25), 3), ((32, y + 1, 4 + i % 25), 0)]
    edits = edits[:4000]
    with hb.World(ctx, 32) as world:
        world.load(pts)
        for p in pts:
            ora.load(p)
        check_remesh(world, ora, check_sync(world, ora))
        world.set_cubes([e[0] for e in edits], [e[1] for e in edits])
        groups = int(world._lib.hb_world_last_edit_groups(world._h))
        assert 50 < groups < len(edits), groups  # neither one serial chain nor all independent
        for pos, mat in edits:
            ora.set_cube(pos, mat)
        check_cubes(world, ora)
        check_remesh(world, ora, check_sync(world, ora))
        # a batch that is one chain from the first edit to the last stays one group
        line = [((10 + i, -50, 20), 0) for i in range(40)]
        world.set_cubes([e[0] for e in line], [e[1] for e in line])
        assert int(world._lib.hb_world_last_edit_groups(world._h)) == 1
        for pos, mat in line:
            ora.set_cube(pos, mat)
        check_cubes(world, ora)
    ora.close()
    ctx.set_world(0)


def test_world_edit_batch_larger_than_a_sub_batch(ctx):
    """70 000 edits in one call: the library groups and applies them in sub-batches of 65 536, one after the other;
    the cubes equal the reference's sequential result."""
    ctx.set_world(0)
    w = O.world(0)
    ora = O.ChunkMapOracle(w)
    pts = [(cx, cy, cz) for cx in (0, 1) for cz in (0, 1) for cy in (-1, 0)]
    rng = np.random.default_rng(9)
    n = 70000
    xyz = np.stack([rng.integers(-2, 66, n), rng.integers(-34, 34, n), rng.integers(-2, 66, n)], axis=1).astype(np.int32)
    xyz[65000:] = xyz[:5000]  # the tail re-edits voxels of the first sub-batch: order across sub-batches matters
    mats = rng.integers(0, 4, n).astype(np.uint16)
    with hb.World(ctx, 16) as world:
        world.load(pts)
        for p in pts:
            ora.load(p)
        world.set_cubes(xyz, mats)
        assert int(world._lib.hb_world_last_edit_groups(world._h)) > 1000
        for p, m in zip(xyz.tolist(), mats.tolist()):
            ora.set_cube(tuple(p), m)
        check_cubes(world, ora)
        check_remesh(world, ora, check_sync(world, ora))
    ora.close()
