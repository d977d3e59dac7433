"""GPU parity tests: libherbgpu (through its C ABI) against the CPU oracle on identical inputs.

Bar: noise bit-exact (tolerance stated by north_star is 1e-5 relative; we assert 0 ulp and report),
block ids / PaletteCube images / quad sets bit-exact INCLUDING order.  The colour channel (wyrand,
unpinned -- see oracle/hb_oracle.h) is asserted separately from geometry.
"""
import numpy as np
import pytest

import herbolution_b200 as hb
import oracle as O

pytestmark = pytest.mark.gpu

SEEDS = [0, 5, 12345, -7]


def split_chunks(out):
    return [out["instances"][int(o):int(o) + int(n)] for o, n in zip(out["offsets"], out["counts"])]


def assert_instances_equal(got, want, what=""):
    assert got.shape == want.shape, f"{what}: quad count {got.shape[0]} != {want.shape[0]}"
    geo = ("model", "light", "ao")
    for f in geo:
        assert got[f].tobytes() == want[f].tobytes(), f"{what}: geometry field {f} differs"
    assert got["rgba"].tobytes() == want["rgba"].tobytes(), f"{what}: colour channel differs (geometry matched)"


def region_points(reg):
    cx0, cz0, ncx, ncz, cy0, ncy = reg
    return hb.Region(cx0, cz0, ncx, ncz, cy0, ncy).chunk_points()


# ---------------------------------------------------------------- noise
@pytest.mark.parametrize("seed", SEEDS)
def test_noise_tiles_bit_exact(ctx, seed):
    ctx.set_world(seed)
    rng = np.random.default_rng(seed & 0xFFFF)
    cols = np.concatenate([
        np.array([[0, 0], [-1, -1], [3, -5], [-512, 511], [262143, -262143], [-262143, 262143]]),
        rng.integers(-4000, 4000, size=(26, 2)),
    ]).astype(np.int32)
    noise, heights = ctx.noise_tiles(cols, want_heights=True)
    w = O.world(seed)
    max_rel = 0.0
    for i, (cx, cz) in enumerate(cols):
        ref = O.noise_tile2(w, int(cx), int(cz))
        rel = np.abs(noise[i] - ref) / np.maximum(np.abs(ref), 1e-30)
        max_rel = max(max_rel, float(rel.max()))
        assert noise[i].view(np.uint32).tolist() == ref.view(np.uint32).tolist(), f"column {cx},{cz} max rel {rel.max()}"
        h = O.heights_from_noise(ref).reshape(32, 32).T.ravel()  # noise index x+z*32 -> heights index x*32+z
        assert (heights[i] == h).all()
    assert max_rel <= 1e-5  # the tolerance north_star states; achieved: 0


def test_noise_custom_params(ctx):
    ctx.set_world(99, freq=0.0123, octaves=3, lacunarity=2.5, gain=0.45)
    w = O.world(99, freq=0.0123, octaves=3, lacunarity=2.5, gain=0.45)
    noise = ctx.noise_tiles([[7, -9], [100, 100]])
    for i, (cx, cz) in enumerate([(7, -9), (100, 100)]):
        assert noise[i].tobytes() == O.noise_tile2(w, cx, cz).tobytes()
    ctx.set_world(0)


def test_noise_range_error(ctx):
    with pytest.raises(hb.HerbGpuError) as e:
        ctx.noise_tiles([[262144, 0]])
    assert e.value.code == -4


def test_noise_fbm3d(ctx):
    for seed in (0, 5):
        ctx.set_world(seed)
        w = O.world(seed)
        got = ctx.noise_fbm3d((-17.0, 40.0, 3.0), (33, 9, 5), (0.01, 0.02, 0.013))
        ref = O.noise_tile3(w, (-17.0, 40.0, 3.0), (33, 9, 5), (0.01, 0.02, 0.013))
        assert got.tobytes() == ref.tobytes()
    with pytest.raises(hb.HerbGpuError):
        ctx.noise_fbm3d((0.5, 0, 0), (4, 4, 4), (0.1, 0.1, 0.1))
    ctx.set_world(0)


# ---------------------------------------------------------------- generation
@pytest.mark.parametrize("seed", [0, 5])
def test_generate_c2_full(ctx, seed):
    """BASELINE config 2: 4096 chunks (32x32 columns, cy in [-2,2)), ids bit-exact vs the oracle."""
    ctx.set_world(seed)
    pts = region_points((-16, -16, 32, 32, -2, 4))
    out = ctx.generate(pts, want_noise=True)
    w = O.world(seed)
    p3 = np.stack([pts["x"], pts["y"], pts["z"]], axis=1)
    ref = O.generate_ids(w, p3)
    mism = int((out["ids"] != ref).sum())
    assert mism == 0, f"{mism} voxels differ"
    # threshold-adjacent voxels (north_star asks for the count): |v*32 - round| tiny
    v32 = out["noise"].astype(np.float64) * 32.0
    near = int((np.abs(v32 - np.round(v32)) <= 1e-5 * 32).sum())
    print(f"seed {seed}: 0 id mismatches; {near} noise samples within 1e-5*32 of an integer height")
    ctx.set_world(0)


def test_generate_cubes_vs_literal_set_path(ctx):
    """PaletteCube image == state left by CubeMesh::set (server/src/chunk/mesh.rs:125-232)."""
    ctx.set_world(0)
    w = O.world(0)
    pts = [(0, 0, 0), (0, -1, 0), (1, 0, -1), (-3, 1, 2), (5, -1, 5), (5, 0, 5), (-8, -2, 7), (31, 1, -31)]
    out = ctx.generate(pts, want_cubes=True)
    for i, p in enumerate(pts):
        m = O.generate_literal(w, p)
        ref = O.mesh_cubes(m)
        assert out["cubes"][i].tobytes() == ref.tobytes(), f"chunk {p}"
        assert (out["ids"][i] == ref["material"]).all()


def test_generate_empty_and_dups(ctx):
    assert ctx.generate(np.zeros((0, 3), dtype=np.int64))["ids"].shape == (0, 32768)
    out = ctx.generate([(2, 0, 2), (2, 0, 2), (2, -1, 2)])
    assert (out["ids"][0] == out["ids"][1]).all()
    with pytest.raises(hb.HerbGpuError) as e:
        ctx.generate([(0, 0, 10**6)])
    assert e.value.code == -4


def test_generator_mirror(ctx):
    params = hb.GenerationParams(5, ctx=ctx)
    gen = hb.ChunkGenerator(params)
    want = [(0, 0, 0), (1, -1, 2)]
    for p in want:
        gen.request(p)
    meshes = list(gen.dequeue())
    assert [tuple(m.position) for m in meshes] == want
    ref = O.generate_ids(O.world(5), want)
    for m, r in zip(meshes, ref):
        assert (m.ids == r).all()
    assert list(gen.dequeue()) == []
    assert params.get_noise((3, 4)).tobytes() == O.noise_tile2(O.world(5), 3, 4).tobytes()
    ctx.set_world(0)


# ---------------------------------------------------------------- meshing
@pytest.mark.parametrize("seed", [0, 5])
def test_mesh_vs_literal_pipeline(ctx, seed):
    """gen (set path) + cull_shared_faces + generate_mesh from flags, vs hb_mesh from ids."""
    ctx.set_world(seed)
    reg = (-3, -3, 6, 6, -3, 6)
    ref = O.region_pipeline(O.world(seed), reg, literal=True, threads=8, want_ids=True)
    pts = region_points(reg)
    out = ctx.mesh(pts, ref["ids"])
    assert (out["counts"] == ref["counts"]).all()
    assert (out["offsets"] % 4 == 0).all()
    got = np.concatenate(split_chunks(out))
    assert_instances_equal(got, ref["instances"], "region")
    ctx.set_world(0)


def toggled_ids(reg, seed, p=0.01, nmat=3):
    ids = O.region_pipeline(O.world(seed), reg, literal=False, threads=8, want_instances=False, want_ids=True)["ids"]
    rng = np.random.default_rng(1234 + seed)
    flip = rng.random(ids.shape) < p
    mats = rng.integers(1, nmat + 1, size=ids.shape).astype(np.uint16)
    out = ids.copy()
    out[flip & (ids != 0)] = 0
    out[flip & (ids == 0)] = mats[flip & (ids == 0)]
    return out


def oracle_mesh_all(pts, ids, nbr, pal=None):
    res = []
    for i in range(len(pts)):
        shell = [ids[j] if j >= 0 else None for j in nbr[i]]
        res.append(O.mesh_ids(shell, (pts["x"][i], pts["y"][i], pts["z"][i]), pal))
    return res


def test_mesh_toggled_caves(ctx):
    """BASELINE config 3 variant: heightmap ids with 1% of voxels toggled (overhangs / caves)."""
    reg = (-2, 1, 4, 4, -2, 3)
    ids = toggled_ids(reg, 0)
    pts = region_points(reg)
    nbr = ctx.build_neighbor_index(pts)
    out = ctx.mesh(pts, ids, nbr)
    ref = oracle_mesh_all(pts, ids, nbr)
    for i, (g, r) in enumerate(zip(split_chunks(out), ref)):
        assert_instances_equal(g, r, f"chunk {i}")


def test_mesh_sparse_set_and_custom_neighbors(ctx):
    """Arbitrary chunk sets: missing neighbours are absent (faces emitted, AO empty)."""
    rng = np.random.default_rng(7)
    pts_all = region_points((0, 0, 3, 3, 0, 3))
    keep = rng.random(len(pts_all)) < 0.6
    keep[13] = True
    pts = pts_all[keep]
    ids = (rng.random((len(pts), 32768)) < 0.5).astype(np.uint16) * rng.integers(1, 4, size=(len(pts), 32768)).astype(np.uint16)
    nbr = ctx.build_neighbor_index(pts)
    out = ctx.mesh(pts, ids, nbr)
    ref = oracle_mesh_all(pts, ids, nbr)
    for i, (g, r) in enumerate(zip(split_chunks(out), ref)):
        assert_instances_equal(g, r, f"chunk {i}")


def test_mesh_edge_cases(ctx):
    one = np.array([(0, 0, 0)])
    # empty batch
    e = ctx.mesh(np.zeros((0, 3), dtype=np.int64), np.zeros((0, 32768), dtype=np.uint16))
    assert e["instances"].shape == (0,)
    # all air
    a = ctx.mesh(one, np.zeros((1, 32768), dtype=np.uint16))
    assert a["counts"][0] == 0 and a["instances"].shape == (0,)
    # all solid, alone: 6 * 1024 boundary faces
    s = ctx.mesh(one, np.full((1, 32768), 2, dtype=np.uint16))
    assert s["counts"][0] == 6144
    assert_instances_equal(split_chunks(s)[0], O.mesh_ids([None] * 13 + [np.full(32768, 2, np.uint16)] + [None] * 13, (0, 0, 0)))
    # 3x3x3 all solid: the centre chunk is fully buried
    pts = region_points((0, 0, 3, 3, 0, 3))
    b = ctx.mesh(pts, np.full((27, 32768), 1, dtype=np.uint16))
    assert b["counts"][13] == 0 and b["counts"].sum() == 54 * 1024
    # 3-D checkerboard: the maximum, 16384 voxels * 6 faces = 98304 quads (many descriptor rounds)
    L = np.arange(32768)
    x, z, y = L >> 10, (L >> 5) & 31, L & 31
    chk = (((x + y + z) & 1) == 0).astype(np.uint16) * 3
    c = ctx.mesh(one, chk[None, :])
    assert c["counts"][0] == 98304
    assert_instances_equal(split_chunks(c)[0], O.mesh_ids([None] * 13 + [chk] + [None] * 13, (0, 0, 0)))


def test_mesh_errors(ctx):
    ids = np.full((1, 32768), 9, dtype=np.uint16)  # material 9 with a 3-entry palette
    with pytest.raises(hb.HerbGpuError) as e:
        ctx.mesh([(0, 0, 0)], ids)
    assert e.value.code == -1
    with pytest.raises(hb.HerbGpuError) as e:
        ctx.mesh([(0, 0, 0)], np.zeros((1, 32768), np.uint16), np.full((1, 27), 5, dtype=np.int32))
    assert e.value.code == -1


def test_mesh_custom_materials(ctx):
    colors = [[(0.1, 0.2, 0.3, 1.0)], [(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0), (0.7, 0.7, 0.7, 1.0), (0.8, 0.8, 0.8, 0.5),
                                        (0.9, 0.9, 0.9, 0.25)]]
    ctx.set_materials(colors)
    try:
        rng = np.random.default_rng(3)
        ids = (rng.random((1, 32768)) < 0.3).astype(np.uint16) * rng.integers(1, 3, size=(1, 32768)).astype(np.uint16)
        out = ctx.mesh([(4, -2, 9)], ids)
        ref = O.mesh_ids([None] * 13 + [ids[0]] + [None] * 13, (4, -2, 9), O.make_palette(colors))
        assert_instances_equal(split_chunks(out)[0], ref)
    finally:
        ctx.set_materials([[(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0), (0.7, 0.7, 0.7, 1.0)],
                           [(0.4, 0.3, 0.2, 1.0), (0.5, 0.4, 0.3, 1.0), (0.6, 0.5, 0.4, 1.0)],
                           [(0.1, 0.8, 0.1, 1.0), (0.2, 0.9, 0.2, 1.0), (0.3, 1.0, 0.3, 1.0)]])


def test_mesher_mirror(ctx):
    ctx.set_world(0)
    reg = (0, 0, 2, 2, -1, 2)
    pts = region_points(reg)
    ids = ctx.generate(pts)["ids"]
    cm = hb.ChunkMap(ctx)
    for p, a in zip(pts, ids):
        cm.insert((p["x"], p["y"], p["z"]), a)
    assert cm.update() == len(pts)
    ref = O.region_pipeline(O.world(0), reg, literal=False, threads=4, want_ids=False)
    off = np.concatenate([[0], np.cumsum(ref["counts"].astype(np.int64))])
    for i, p in enumerate(pts):
        key = hb.ChunkPt(int(p["x"]), int(p["y"]), int(p["z"]))
        assert_instances_equal(cm.meshes[key], ref["instances"][off[i]:off[i + 1]], str(key))
    # single-chunk drop-in signature
    shell = hb.create_chunk_shell(cm.map, hb.ChunkPt(0, 0, 0))
    inst = []
    hb.generate_mesh(ctx, shell, inst)
    assert_instances_equal(inst[0], cm.meshes[hb.ChunkPt(0, 0, 0)])


@pytest.mark.parametrize("seed", [0, 5])
def test_mesh_c3_full(ctx, seed):
    """BASELINE config 3: 4096 pre-generated chunks incl. cross-chunk borders, bit-exact quad set."""
    ctx.set_world(seed)
    reg = (-16, -16, 32, 32, -2, 4)
    pts = region_points(reg)
    ids = ctx.generate(pts)["ids"]
    out = ctx.mesh(pts, ids)
    ref = O.region_pipeline(O.world(seed), reg, literal=False, threads=8)
    assert (out["counts"] == ref["counts"]).all()
    got = np.concatenate(split_chunks(out))
    assert_instances_equal(got, ref["instances"], "C3")
    ctx.set_world(0)


# ---------------------------------------------------------------- fused region
def collect_region(ctx, region, want_ids=False):
    chunks = {}

    def sink(first, off, cnt, inst, ids):
        for k in range(len(cnt)):
            o, n = int(off[k]), int(cnt[k])
            chunks[first + k] = (inst[o:o + n].copy(), None if ids is None else ids[k].copy())

    tot = ctx.generate_mesh_region(region, want_ids=want_ids, sink=sink)
    return tot, chunks


@pytest.mark.parametrize("seed,reg", [(0, (-3, -3, 6, 6, -3, 6)), (5, (10, -20, 5, 3, -1, 2)), (0, (0, 0, 1, 1, 0, 1)),
                                      (5, (3, 3, 1, 2, -20, 40)), (0, (-1, 0, 2, 1, -1, 1)), (0, (0, 0, 2, 2, 3, 2)),
                                      (0, (-2, -2, 3, 3, -3, 2)), (5, (0, 0, 2, 2, -2, 1))])  # region top below the terrain
def test_region_fused_vs_oracle(ctx, seed, reg):
    ctx.set_world(seed)
    ref = O.region_pipeline(O.world(seed), reg, literal=True, threads=8, want_ids=True)
    tot, chunks = collect_region(ctx, hb.Region(*reg), want_ids=True)
    n = reg[2] * reg[3] * reg[5]
    assert tot["chunks"] == n and tot["quads"] == ref["total"]
    off = np.concatenate([[0], np.cumsum(ref["counts"].astype(np.int64))])
    for i in range(n):
        inst, ids = chunks[i]
        assert_instances_equal(inst, ref["instances"][off[i]:off[i + 1]], f"chunk {i}")
        assert (ids == ref["ids"][i]).all()
    ctx.set_world(0)


def test_region_plan_slabs_equal_whole(ctx):
    """A region processed as x-slabs (how it is sharded across GPUs) equals the whole region."""
    ctx.set_world(0)
    region = hb.Region(-4, -2, 7, 4, -3, 6)
    _, whole = collect_region(ctx, region)
    ncz, ncy = region.ncz, region.ncy
    for (a, b) in [(0, 3), (3, 4), (4, 7)]:
        plan = ctx.region_plan(region, ix_begin=a, ix_end=b)
        plan.enqueue(hb.STAGE_ALL_MESH)
        res = plan.result()
        n = res["n_chunks"]
        assert n == (b - a) * ncz * ncy
        off = ctx.download(res["d_offsets"], np.uint64, n)
        cnt = ctx.download(res["d_counts"], np.uint32, n)
        inst = ctx.download(res["d_instances"], hb.INSTANCE_DTYPE, res["span"])
        for k in range(n):
            g = inst[int(off[k]):int(off[k]) + int(cnt[k])]
            assert_instances_equal(g, whole[a * ncz * ncy + k][0], f"slab {a}:{b} chunk {k}")
        plan.close()


def test_region_slabs_through_host_api_equal_whole(ctx):
    """BASELINE config 5 in miniature: the region is cut into x-slabs as herbolution_b200.sharding does per
    GPU; every slab is generated + meshed on its own (apron recomputed, no exchange) and streams back through the
    pinned sink.  The union equals the unsharded region chunk for chunk."""
    from herbolution_b200 import sharding
    ctx.set_world(5)
    region = hb.Region(-5, 2, 9, 3, -3, 6)
    _, whole = collect_region(ctx, region, want_ids=True)
    for world in (2, 4):
        got = {}

        def sink(first, off, cnt, inst, ids):
            for k in range(len(cnt)):
                assert first + k not in got
                got[first + k] = (inst[int(off[k]):int(off[k]) + int(cnt[k])].copy(), ids[k].copy())

        totals = [ctx.generate_mesh_region(region, want_ids=True, sink=sink, ix_begin=a, ix_end=b)
                  for a, b in (sharding.slab_bounds(region.ncx, world, r) for r in range(world))]
        assert sum(t["chunks"] for t in totals) == region.n_chunks == len(got)
        for i in range(region.n_chunks):
            assert_instances_equal(got[i][0], whole[i][0], f"world {world} chunk {i}")
            assert (got[i][1] == whole[i][1]).all()
    ctx.set_world(0)


def test_region_c4_properties(ctx):
    """BASELINE config 4 at a reduced but still multi-slab size (64x64 columns x 6): size-independent
    properties -- quad total equals the sum over independently meshed sub-blocks' interior-consistent
    counts is not available, so we check: (1) fused == hb_generate + hb_mesh on a random sample of
    columns with their full neighbourhoods, (2) every offset is 4-aligned and spans do not overlap."""
    ctx.set_world(0)
    region = hb.Region(-32, -32, 64, 64, -3, 6)
    counts = {}
    keep = {}
    rng = np.random.default_rng(11)
    sample = set(int(v) for v in rng.integers(0, region.n_chunks, size=64))

    def sink(first, off, cnt, inst, ids):
        assert (off % 4 == 0).all()
        ends = off + ((cnt.astype(np.uint64) + 3) // 4) * 4
        assert (ends[:-1] <= off[1:]).all() and (len(inst) == 0 or ends[-1] <= len(inst))
        for k in range(len(cnt)):
            counts[first + k] = int(cnt[k])
            if first + k in sample:
                keep[first + k] = inst[int(off[k]):int(off[k]) + int(cnt[k])].copy()

    tot = ctx.generate_mesh_region(region, sink=sink)
    assert tot["chunks"] == region.n_chunks == len(counts)
    assert tot["quads"] == sum(counts.values())
    pts = region.chunk_points()
    for c in sorted(sample):
        # gather the 27-neighbourhood inside the region, generate + mesh it with the general path
        ix, rem = divmod(c, region.ncz * region.ncy)
        iz, iy = divmod(rem, region.ncy)
        sel = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    jx, jy, jz = ix + dx, iy + dy, iz + dz
                    if 0 <= jx < region.ncx and 0 <= jy < region.ncy and 0 <= jz < region.ncz:
                        sel.append((jx * region.ncz + jz) * region.ncy + jy)
        sub = pts[sel]
        ids = ctx.generate(sub)["ids"]
        out = ctx.mesh(sub, ids)
        k = sel.index(c)
        g = out["instances"][int(out["offsets"][k]):int(out["offsets"][k]) + int(out["counts"][k])]
        assert_instances_equal(keep[c], g, f"chunk {c}")


# ---------------------------------------------------------------- literal, flags-driven mesher (hb_mesh_cubes)
def oracle_mesh_cubes_all(pts, cubes, nbr, pal=None):
    res = []
    for i in range(len(pts)):
        shell = [cubes[j] if j >= 0 else None for j in nbr[i]]
        res.append(O.mesh_cubes_literal(shell, (pts["x"][i], pts["y"][i], pts["z"][i]), pal))
    return res


def test_mesh_cubes_after_set_and_cull(ctx):
    """PaletteCube arrays exactly as the server leaves them: CubeMesh::set per voxel, then cull_shared_faces
    for every adjacent pair (load_provided), then one dig + one place through set() in a border voxel."""
    import ctypes as C
    w = O.world(0)
    reg = (-1, 2, 3, 2, -2, 3)
    pts = region_points(reg)
    meshes = [O.generate_literal(w, (p["x"], p["y"], p["z"])) for p in pts]
    nbr = ctx.build_neighbor_index(pts)
    for i in range(len(pts)):
        for f_shell in (22, 16, 14):  # +x, +y, +z neighbours: each adjacent pair once
            j = nbr[i][f_shell]
            if j >= 0:
                O.lib().hbo_cull_shared_faces(C.byref(meshes[i]), C.byref(meshes[j]))
    # incremental edits through the literal set() path (dig the top voxel of a column, place one in the air)
    for m in meshes[:4]:
        cubes = O.mesh_cubes(m)
        solid = np.flatnonzero(cubes["material"] != 0)
        if len(solid):
            L = int(solid[-1])
            O.lib().hbo_mesh_set(C.byref(m), L >> 10, L & 31, (L >> 5) & 31, 0)
        O.lib().hbo_mesh_set(C.byref(m), 31, 31, 31, 2)
    cubes = np.stack([O.mesh_cubes(m) for m in meshes])
    out = ctx.mesh_cubes(pts, cubes, nbr)
    ref = oracle_mesh_cubes_all(pts, cubes, nbr)
    assert out["counts"].sum() > 0
    for i, (g, r) in enumerate(zip(split_chunks(out), ref)):
        assert_instances_equal(g, r, f"chunk {i}")


def test_mesh_cubes_arbitrary_flags(ctx):
    """The faces drawn are whatever cube.flags says (even inconsistent with the materials, as
    ChunkMap::set_cube's unconditional insert_faces on border neighbours can leave them)."""
    rng = np.random.default_rng(21)
    pts = region_points((0, 0, 2, 1, 0, 2))
    n = len(pts)
    cubes = np.zeros((n, 32768), dtype=hb.CUBE_DTYPE)
    cubes["material"] = (rng.random((n, 32768)) < 0.35) * rng.integers(1, 4, size=(n, 32768))
    cubes["flags"] = rng.integers(0, 128, size=(n, 32768)) * (rng.random((n, 32768)) < 0.2)
    nbr = ctx.build_neighbor_index(pts)
    out = ctx.mesh_cubes(pts, cubes, nbr)
    ref = oracle_mesh_cubes_all(pts, cubes, nbr)
    for i, (g, r) in enumerate(zip(split_chunks(out), ref)):
        assert_instances_equal(g, r, f"chunk {i}")
    # flags on air voxels and flags without the opaque tag draw nothing
    cubes2 = np.zeros((1, 32768), dtype=hb.CUBE_DTYPE)
    cubes2["flags"] = 0x7F
    assert ctx.mesh_cubes([(0, 0, 0)], cubes2)["counts"][0] == 0
    cubes2["material"] = 1
    cubes2["flags"] = 0x7E
    assert ctx.mesh_cubes([(0, 0, 0)], cubes2)["counts"][0] == 0
    cubes2["flags"] = 0x7F
    assert ctx.mesh_cubes([(0, 0, 0)], cubes2)["counts"][0] == 6 * 32768


def test_mesh_cubes_equals_mesh_on_settled_world(ctx):
    """On a settled world (generate + cull of all pairs) flags-driven and recomputed visibility agree."""
    ctx.set_world(5)
    reg = (4, 4, 2, 2, -1, 2)
    pts = region_points(reg)
    ref = O.region_pipeline(O.world(5), reg, literal=True, threads=4, want_ids=True)
    a = ctx.mesh(pts, ref["ids"])
    assert (a["counts"] == ref["counts"]).all()
    ctx.set_world(0)


# ---------------------------------------------------------------- other noise kinds of the crate
@pytest.mark.parametrize("kind", [O.NOISE_RIDGE, O.NOISE_TURBULENCE, O.NOISE_GRADIENT])
def test_noise_kinds_bit_exact(ctx, kind):
    """Ridge / turbulence / gradient sampling (functions/{ridge,turbulence}_32.rs, noise/gradient.rs) through the same
    tile walks: 2-D tiles + heights, the 3-D tile, and block ids generated from those heights."""
    try:
        for seed in (0, 5):
            ctx.set_world(seed)
            ctx.set_noise_kind(kind)
            w = O.world(seed, kind=kind)
            cols = np.array([[0, 0], [-1, -1], [7, -3], [-512, 511], [262143, -262143]], dtype=np.int32)
            noise, heights = ctx.noise_tiles(cols, want_heights=True)
            for i, (cx, cz) in enumerate(cols):
                ref = O.noise_tile2(w, int(cx), int(cz))
                assert noise[i].view(np.uint32).tolist() == ref.view(np.uint32).tolist(), (kind, seed, cx, cz)
                assert (heights[i] == O.heights_from_noise(ref).reshape(32, 32).T.ravel()).all()
            got3 = ctx.noise_fbm3d((-5.0, 100.0, 3.0), (9, 5, 7), (0.013, 0.02, 0.011))
            ref3 = O.noise_tile3(w, (-5.0, 100.0, 3.0), (9, 5, 7), (0.013, 0.02, 0.011))
            assert got3.view(np.uint32).tolist() == ref3.view(np.uint32).tolist()
            pts = region_points((2, -2, 2, 2, 3, 4)) if kind == O.NOISE_RIDGE else region_points((2, -2, 2, 2, -1, 3))
            ids = ctx.generate(pts)["ids"]
            want = O.generate_ids(w, [(int(p["x"]), int(p["y"]), int(p["z"])) for p in pts])
            assert (ids == want).all()
            assert 0 < (ids != 0).sum() < ids.size  # the surface is inside the sampled layers
    finally:
        ctx.set_noise_kind(O.NOISE_FBM)
        ctx.set_world(0)


@pytest.mark.parametrize("kind", [O.NOISE_FBM, O.NOISE_RIDGE, O.NOISE_TURBULENCE, O.NOISE_GRADIENT])
def test_unfused_arithmetic_column_bit_exact(ctx, kind):
    """hb_set_fma_mode(0): the SSE2 / SSE4.1 / scalar column of the reference's runtime dispatch, where neg_mul_add is a
    rounded product and a subtraction (simd/ops/f32.rs:141-162).  Bit-exact against the oracle's unfused switch for
    2-D tiles, heights, the 3-D tile, block ids and a meshed region; and it really is a different column (some sample
    differs from the fused one)."""
    try:
        for seed in (0, 5):
            ctx.set_world(seed)
            ctx.set_noise_kind(kind)
            cols = np.array([[0, 0], [-1, -1], [7, -3], [-512, 511], [262143, -262143]], dtype=np.int32)
            fused_noise = ctx.noise_tiles(cols)
            ctx.set_fma_mode(False)
            w = O.world(seed, kind=kind, fused=0)
            noise, heights = ctx.noise_tiles(cols, want_heights=True)
            for i, (cx, cz) in enumerate(cols):
                ref = O.noise_tile2(w, int(cx), int(cz))
                assert noise[i].view(np.uint32).tolist() == ref.view(np.uint32).tolist(), (kind, seed, cx, cz)
                assert (heights[i] == O.heights_from_noise(ref).reshape(32, 32).T.ravel()).all()
            assert (noise.view(np.uint32) != fused_noise.view(np.uint32)).any()
            # the two CPU columns differ by rounding only: 1e-5 of the noise's unit scale (values near zero lose
            # relative, not absolute, agreement through cancellation between the octaves)
            assert float(np.abs(noise - fused_noise).max()) <= 1e-5
            got3 = ctx.noise_fbm3d((-5.0, 100.0, 3.0), (9, 5, 7), (0.013, 0.02, 0.011))
            ref3 = O.noise_tile3(w, (-5.0, 100.0, 3.0), (9, 5, 7), (0.013, 0.02, 0.011))
            assert got3.view(np.uint32).tolist() == ref3.view(np.uint32).tolist()
            if kind == O.NOISE_FBM:
                reg = (3, -4, 3, 2, -2, 4)
                got = {}

                def sink(first, off, cnt, inst, ids):
                    for k in range(len(cnt)):
                        got[first + k] = (inst[int(off[k]):int(off[k]) + int(cnt[k])].tobytes(), ids[k].copy())

                ctx.generate_mesh_region(hb.Region(*reg), want_ids=True, sink=sink)
                ref = O.region_pipeline(w, reg, literal=True, threads=2, want_ids=True)
                off = np.concatenate([[0], np.cumsum(ref["counts"].astype(np.int64))])
                for i in range(len(ref["counts"])):
                    assert (got[i][1] == ref["ids"][i]).all()
                    assert got[i][0] == ref["instances"][off[i]:off[i + 1]].tobytes()
            ctx.set_fma_mode(True)
    finally:
        ctx.set_fma_mode(True)
        ctx.set_noise_kind(O.NOISE_FBM)
        ctx.set_world(0)
    assert ctx._lib.hb_set_fma_mode(ctx._h, 2) == hb._lib.HB_ERR_INVALID


def test_noise_kind_errors(ctx):
    with pytest.raises(hb.HerbGpuError) as e:
        ctx.set_noise_kind(7)
    assert e.value.code == hb._lib.HB_ERR_INVALID


def test_region_config4_full_size(ctx):
    """BASELINE config 4 at its full size (256x256 columns x 6 = 393 216 chunks, 175 M quads, 17.5 GB of instances,
    device-resident plan): size-independent layout properties over all chunks, and bit-exact parity with the oracle
    on the interior of sampled 6x6-column patches (the patch's outer ring only supplies the neighbours)."""
    ctx.set_world(0)
    reg = (-128, -128, 256, 256, -3, 6)
    region = hb.Region(*reg)
    plan = ctx.region_plan(region)
    plan.enqueue(hb.STAGE_ALL_MESH)
    res = plan.result()
    n = region.n_chunks
    assert res["n_chunks"] == n
    counts = ctx.download(res["d_counts"], np.uint32, n).astype(np.int64)
    offsets = ctx.download(res["d_offsets"], np.uint64, n).astype(np.int64)
    spans = (counts + 3) // 4 * 4
    assert counts.sum() == res["quads"] and spans.sum() == res["span"]
    assert (offsets == np.concatenate([[0], np.cumsum(spans)[:-1]])).all()     # deterministic exclusive scan, no gaps
    # every column of the region bottom shows its 1024 Down faces at least
    per = counts.reshape(256, 256, 6)
    assert (per[:, :, 0] >= 1024).all()
    # a second run of the same plan reproduces the same counts (no atomics on the path)
    plan.enqueue(hb.STAGE_ALL_MESH)
    res2 = plan.result()
    assert (ctx.download(res2["d_counts"], np.uint32, n) == counts).all() and res2["quads"] == res["quads"]
    rng = np.random.default_rng(2024)
    w = O.world(0)
    # 6x6-column patches meshed by the oracle as regions of their own, ALL six layers (layer 0 = the region-bottom fast
    # path with 39 % of the benchmark's quads, layer 5 = the region top): three random interior patches, the region's
    # corner, and one patch on each of two borders.  A column is comparable iff every one of its 8 neighbour columns is
    # inside the patch (the oracle saw it) or outside the region (absent for both).
    spots = [(int(rng.integers(1, 249)), int(rng.integers(1, 249))) for _ in range(3)]
    spots += [(0, 0), (250, int(rng.integers(1, 249))), (int(rng.integers(1, 249)), 250)]
    compared = slab_chunks = top_chunks = 0
    for px, pz in spots:
        patch = (reg[0] + px, reg[1] + pz, 6, 6, reg[4], reg[5])
        ref = O.region_pipeline(w, patch, literal=True, threads=8)
        roff = np.concatenate([[0], np.cumsum(ref["counts"].astype(np.int64))])
        for ix in range(6):
            for iz in range(6):
                ok = True
                for dx in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        ax, az = px + ix + dx, pz + iz + dz
                        in_patch = 0 <= ix + dx < 6 and 0 <= iz + dz < 6
                        in_region = 0 <= ax < 256 and 0 <= az < 256
                        ok = ok and (in_patch or not in_region)
                if not ok:
                    continue
                for iy in range(6):
                    g = ((px + ix) * 256 + (pz + iz)) * 6 + iy
                    k = (ix * 6 + iz) * 6 + iy
                    assert counts[g] == ref["counts"][k], (px, pz, ix, iz, iy)
                    compared += 1
                    if counts[g]:
                        slab_chunks += iy == 0
                        top_chunks += iy == 5
                        inst = ctx.download(res2["d_instances"] + int(offsets[g]) * 100, hb.INSTANCE_DTYPE, int(counts[g]))
                        assert_instances_equal(inst, ref["instances"][roff[k]:roff[k + 1]], f"patch {px},{pz} chunk {ix},{iz},{iy}")
    assert compared >= 6 * 16 * 6 and slab_chunks >= 96  # every compared column has a bottom chunk with >= 1024 quads
    plan.close()


def test_concurrent_callers_share_one_context(ctx):
    """The reference calls its generator and mesher from many rayon workers at once (server/src/generator/mod.rs:36,
    client/src/world/chunk.rs:66-75); every hb_* entry point is thread-safe per ctx.  Eight threads interleave
    generate / mesh / noise / codec calls on one context; each result equals the single-threaded one."""
    import threading
    ctx.set_world(0)
    jobs = []
    for t in range(8):
        reg = (4 * t - 16, 3 - t, 2, 2, -1, 2)
        pts = region_points(reg)
        ids = ctx.generate(pts)["ids"]
        mesh = ctx.mesh(pts, ids)
        jobs.append((pts, ids, mesh, ctx.noise_tiles([[reg[0], reg[1]]]), ctx.rle_encode(ids)["bytes"]))
    errors = []

    def worker(k):
        try:
            pts, ids, mesh, noise, enc = jobs[k]
            for _ in range(4):
                got_ids = ctx.generate(pts)["ids"]
                assert (got_ids == ids).all()
                got = ctx.mesh(pts, got_ids)
                assert got["instances"].tobytes() == mesh["instances"].tobytes() and (got["offsets"] == mesh["offsets"]).all()
                assert ctx.noise_tiles([[int(pts[0]["x"]), int(pts[0]["z"])]]).tobytes() == noise.tobytes()
                assert ctx.rle_encode(got_ids)["bytes"].tobytes() == enc.tobytes()
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
