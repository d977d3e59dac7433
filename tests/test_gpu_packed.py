"""Packed-quad transport (include/herbgpu.h: hb_packed_quad, HB_TRANSPORT_*, hb_expand_quads): what the sink sees
must be byte-identical whichever way the quads crossed the link, and identical to the oracle's Instance3d stream.
All calls go through the C ABI."""
import numpy as np
import pytest

import herbolution_b200 as hb
import oracle as O

pytestmark = pytest.mark.gpu


def collect(ctx, region, want_ids=False, **kw):
    chunks, raw = {}, []

    def sink(first, off, cnt, inst, ids):
        raw.append((first, off.copy(), cnt.copy(), inst.tobytes()))
        for k in range(len(cnt)):
            chunks[first + k] = (inst[int(off[k]):int(off[k]) + int(cnt[k])].copy(), None if ids is None else ids[k].copy())

    tot = ctx.generate_mesh_region(region, want_ids=want_ids, sink=sink, **kw)
    return tot, chunks, raw


@pytest.fixture()
def transports(ctx):
    yield
    ctx.set_region_transport(hb.TRANSPORT_PACKED)
    ctx.set_host_threads(8)
    ctx.set_world(0)


@pytest.mark.parametrize("seed,reg", [(0, (-3, -3, 6, 6, -3, 6)), (5, (3, 3, 1, 2, -20, 40)), (0, (0, 0, 2, 2, 3, 2)),
                                      (5, (10, -20, 5, 3, -1, 2))])
def test_packed_transport_equals_instances_and_oracle(ctx, transports, seed, reg):
    ctx.set_world(seed)
    region = hb.Region(*reg)
    ctx.set_region_transport(hb.TRANSPORT_INSTANCES)
    tot_i, _, raw_i = collect(ctx, region)
    ctx.set_region_transport(hb.TRANSPORT_PACKED)
    tot_p, chunks_p, raw_p = collect(ctx, region)
    assert tot_i == tot_p
    assert len(raw_i) == len(raw_p)
    for (fa, oa, ca, ba), (fb, ob, cb, bb) in zip(raw_i, raw_p):
        assert fa == fb and (oa == ob).all() and (ca == cb).all()
        assert ba == bb, "expanded slab differs from the device-written instances (incl. the zero-filled gaps)"
    ref = O.region_pipeline(O.world(seed), reg, literal=True, threads=8)
    off = np.concatenate([[0], np.cumsum(ref["counts"].astype(np.int64))])
    for i in range(region.n_chunks):
        assert chunks_p[i][0].tobytes() == ref["instances"][off[i]:off[i + 1]].tobytes(), f"chunk {i}"


def test_packed_multi_slab_and_thread_counts(ctx, transports):
    """64x64 columns = several slabs through the double-buffered pipeline; 1, 3 and 16 host threads give the same bytes."""
    ctx.set_world(0)
    region = hb.Region(-32, -32, 64, 64, -3, 6)
    ctx.set_region_transport(hb.TRANSPORT_INSTANCES)
    tot_i, _, raw_i = collect(ctx, region)
    assert len(raw_i) > 2
    ctx.set_region_transport(hb.TRANSPORT_PACKED)
    for threads in (1, 3, 16):
        ctx.set_host_threads(threads)
        tot_p, _, raw_p = collect(ctx, region)
        assert tot_p == tot_i and len(raw_p) == len(raw_i)
        for (fa, oa, ca, ba), (fb, ob, cb, bb) in zip(raw_i, raw_p):
            assert fa == fb and (oa == ob).all() and (ca == cb).all() and ba == bb, f"{threads} host threads"


def test_packed_sink_and_expand_quads(ctx, transports):
    """hb_generate_mesh_region_packed hands out the 4-byte words; hb_expand_quads rebuilds each chunk's records."""
    ctx.set_world(5)
    ctx.set_materials([[(0.1 * k, 0.2, 0.3, 1.0) for k in range(1, 9)], [(0.5, 0.5, 0.5, 1.0)],
                       [(0.9, 0.1 * k, 0.0, 1.0) for k in range(5)]])
    region = hb.Region(-2, 1, 4, 3, -3, 6)
    try:
        _, want, _ = collect(ctx, region, want_ids=True)
        got = {}

        def psink(first, off, cnt, packed, ids):
            assert (off % 4 == 0).all()
            for k in range(len(cnt)):
                q = packed[int(off[k]):int(off[k]) + int(cnt[k])].copy()
                gap = packed[int(off[k]) + int(cnt[k]):int(off[k]) + ((int(cnt[k]) + 3) & ~3)]
                assert (gap == 0).all()
                got[first + k] = (q, ids[k].copy())

        tot = ctx.generate_mesh_region_packed(region, want_ids=True, sink=psink)
        assert tot["chunks"] == region.n_chunks
        pts = region.chunk_points()
        n_quads = 0
        for i in range(region.n_chunks):
            q, ids = got[i]
            assert (ids == want[i][1]).all()
            L, f, mat = q & 0x7FFF, (q >> 15) & 7, ((q >> 26) & 15) + 1
            assert (f < 6).all() and (q >> 30 == 0).all()
            assert (mat == ids[L]).all(), "material bits must be the block id of the voxel"
            key = L.astype(np.int64) * 8 + f
            assert (np.diff(key) > 0).all(), "reference emission order: ascending L, then face"
            inst = ctx.expand_quads((pts["x"][i], pts["y"][i], pts["z"][i]), q)
            assert inst.tobytes() == want[i][0].tobytes(), f"chunk {i}"
            n_quads += len(q)
        assert n_quads == tot["quads"]
    finally:
        ctx.set_materials([[(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0), (0.7, 0.7, 0.7, 1.0)],
                           [(0.4, 0.3, 0.2, 1.0), (0.5, 0.4, 0.3, 1.0), (0.6, 0.5, 0.4, 1.0)],
                           [(0.1, 0.8, 0.1, 1.0), (0.2, 0.9, 0.2, 1.0), (0.3, 1.0, 0.3, 1.0)]])


def test_plan_emit_packed_matches_emit(ctx, transports):
    """Device-resident plan: STAGE_EMIT_PACKED words expand to exactly what STAGE_EMIT wrote."""
    ctx.set_world(0)
    region = hb.Region(-4, -2, 5, 4, -3, 6)
    plan = ctx.region_plan(region)
    plan.enqueue(hb.STAGE_ALL_MESH | hb.STAGE_EMIT_PACKED)
    res = plan.result()
    n = res["n_chunks"]
    off = ctx.download(res["d_offsets"], np.uint64, n)
    cnt = ctx.download(res["d_counts"], np.uint32, n)
    inst = ctx.download(res["d_instances"], hb.INSTANCE_DTYPE, res["span"])
    packed = ctx.download(plan.packed_ptr(), np.uint32, res["span"])
    pts = region.chunk_points()
    for k in range(n):
        a, c = int(off[k]), int(cnt[k])
        e = ctx.expand_quads((pts["x"][k], pts["y"][k], pts["z"][k]), packed[a:a + c])
        assert e.tobytes() == inst[a:a + c].tobytes(), f"chunk {k}"
    plan.close()


def test_expand_quads_errors(ctx):
    lib = hb._lib.load()
    q = np.zeros(4, dtype=np.uint32)
    out = np.zeros(4 * 100 + 16, dtype=np.uint8)
    o = (-out.ctypes.data) & 15
    import ctypes as C
    assert lib.hb_expand_quads(ctx._h, hb._lib.ChunkPtC(0, 0, 0), hb._lib.ptr(q), 4, C.c_void_p(out.ctypes.data + o + 4)) == hb._lib.HB_ERR_INVALID
    assert lib.hb_expand_quads(ctx._h, hb._lib.ChunkPtC(0, 0, 0), None, 4, C.c_void_p(out.ctypes.data + o)) == hb._lib.HB_ERR_INVALID
    assert lib.hb_expand_quads(ctx._h, hb._lib.ChunkPtC(0, 0, 0), None, 0, None) == hb._lib.HB_OK
    with pytest.raises(hb.HerbGpuError):
        ctx.set_region_transport(7)
    with pytest.raises(hb.HerbGpuError):
        ctx.set_host_threads(0)


def test_multi_context_python_one_process():
    """hb_multi through the Python binding: all visible devices (one on the test box), serialised sink, same bytes as
    the single-ctx call; slabs cover the region."""
    import torch
    reg = hb.Region(-3, 2, 11, 5, -2, 4)
    with hb.MultiContext(devices=None, seed=5) as m:
        assert m.n_devices == min(torch.cuda.device_count(), 8)
        edges = [m.slab(reg, i) for i in range(m.n_devices)]
        assert edges[0][0] == 0 and edges[-1][1] == reg.ncx and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        got = {}

        def sink(first, off, cnt, inst, ids):
            for k in range(len(cnt)):
                got[first + k] = inst[int(off[k]):int(off[k]) + int(cnt[k])].tobytes()

        tot = m.generate_mesh_region(reg, sink=sink)
    c = hb.Context(0, seed=5)
    want = {}

    def sink1(first, off, cnt, inst, ids):
        for k in range(len(cnt)):
            want[first + k] = inst[int(off[k]):int(off[k]) + int(cnt[k])].tobytes()

    tot1 = c.generate_mesh_region(reg, sink=sink1)
    c.close()
    assert tot == tot1 and got == want and len(got) == 11 * 5 * 4


@pytest.mark.parametrize("seed,reg,kind", [(0, (-20, -6, 40, 12, -3, 6), 0), (5, (3, 3, 2, 2, -40, 80), 0), (5, (100, -100, 3, 3, 2, 3), 1),
                                           (0, (262140, -262143, 3, 2, -2, 4), 0)])
def test_block_ids_rebuilt_on_the_host_equal_the_device_fill(ctx, transports, seed, reg, kind):
    """want_ids with the packed transports: the heights tiles cross the link (4 KiB per column) and the host threads
    write the ids (csrc/expand.cu: fill_chunk_ids); with the instance transport k_fill writes them on the device and
    they are copied (64 KiB per chunk).  Same bytes, and the oracle's generate() agrees -- several slabs, 80 layers,
    another noise kind, the coordinate limit."""
    try:
        ctx.set_world(seed)
        ctx.set_noise_kind(kind)
        region = hb.Region(*reg)
        ctx.set_region_transport(hb.TRANSPORT_INSTANCES)
        _, dev, _ = collect(ctx, region, want_ids=True)
        ctx.set_region_transport(hb.TRANSPORT_PACKED)
        _, host, _ = collect(ctx, region, want_ids=True)
        got = {}

        def psink(first, off, cnt, packed, ids):
            for k in range(len(cnt)):
                got[first + k] = ids[k].copy()

        ctx.generate_mesh_region_packed(region, want_ids=True, sink=psink)
        assert len(dev) == len(host) == len(got) == region.n_chunks
        for i in range(region.n_chunks):
            assert (host[i][1] == dev[i][1]).all(), f"chunk {i}: host-rebuilt ids differ from the device fill"
            assert (got[i] == dev[i][1]).all(), f"chunk {i}: packed sink ids differ"
        pts = region.chunk_points()
        step = max(1, region.n_chunks // 40)
        w = O.world(seed, kind=kind)
        sel = list(range(0, region.n_chunks, step))
        ref = O.generate_ids(w, [(int(pts["x"][i]), int(pts["y"][i]), int(pts["z"][i])) for i in sel])
        for k, i in enumerate(sel):
            assert (host[i][1] == ref[k]).all(), f"chunk {i} vs oracle"
    finally:
        ctx.set_noise_kind(0)


def test_generate_ids_host_fill_equals_device_fill(ctx, transports):
    """hb_generate's block ids follow the transport too: heights tiles + host fill (default) vs k_fill + 64 KiB copies;
    duplicated points, several layers per column, an unaligned destination (falls back to the device fill)."""
    ctx.set_world(5)
    pts = hb.Region(-5, 7, 6, 5, -3, 6).chunk_points()
    pts = np.concatenate([pts, pts[:7]])
    ctx.set_region_transport(hb.TRANSPORT_INSTANCES)
    dev = ctx.generate(pts)["ids"]
    ctx.set_region_transport(hb.TRANSPORT_PACKED)
    host = ctx.generate(pts)["ids"]
    assert (dev == host).all()
    ref = O.generate_ids(O.world(5), [(int(p["x"]), int(p["y"]), int(p["z"])) for p in pts[::9]])
    assert (host[::9] == ref).all()
    # unaligned output pointer through the raw ABI
    import ctypes as C
    raw = np.zeros(len(pts) * 32768 + 8, dtype=np.uint16)
    view = raw[1:1 + len(pts) * 32768]
    assert view.ctypes.data % 16 != 0
    p2 = np.ascontiguousarray(pts)
    rc = ctx._lib.hb_generate(ctx._h, C.c_void_p(p2.ctypes.data), len(pts), C.c_void_p(view.ctypes.data), None, None)
    assert rc == 0 and (view.reshape(-1, 32768) == dev).all()
