"""hb_dev_mesh (device pointers, nothing synchronised): equals hb_mesh, and an instance buffer that is too small is
reported through d_total[2] without a single byte written past it (the pipelined emit kernel skips such chunks)."""
import numpy as np
import pytest

import herbolution_b200 as hb

pytestmark = pytest.mark.gpu


def _dev_mesh(ctx, torch, pts, ids, nbr, capacity, guard=4096):
    lib, dev = ctx._lib, torch.device("cuda", ctx.device)
    n = pts.shape[0]
    p = lambda t: t.data_ptr()  # noqa: E731
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(dev)  # noqa: E731
    d_pts, d_ids, d_nbr = up(pts), up(ids), up(nbr)
    ws = torch.zeros(lib.hb_dev_mesh_workspace_bytes(n), dtype=torch.uint8, device=dev)
    d_off = torch.zeros(n, dtype=torch.int64, device=dev)
    d_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
    d_tot = torch.zeros(4, dtype=torch.int64, device=dev)
    d_out = torch.full((capacity * 100 + guard,), 0xA5, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        hb._lib.check(lib.hb_dev_mesh(ctx._h, st.cuda_stream, p(d_pts), n, p(d_ids), p(d_nbr), p(ws), p(d_out), capacity, p(d_off),
                                      p(d_cnt), p(d_tot)))
    st.synchronize()
    return {"out": d_out.cpu().numpy(), "offsets": d_off.cpu().numpy().astype(np.uint64), "counts": d_cnt.cpu().numpy().astype(np.uint32),
            "total": d_tot.cpu().numpy()}


def test_dev_mesh_equals_host_mesh_and_respects_capacity(ctx):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    pts = np.array([(x, y, z) for x in range(3) for y in range(2) for z in range(3)], dtype=np.int32)
    n = pts.shape[0]
    ids = np.zeros((n, 32, 32, 32), dtype=np.uint16)
    for i in range(n):  # caves, an empty chunk, a solid one, a sparse one
        kind = i % 4
        if kind == 0:
            ids[i] = (rng.random((32, 32, 32)) < 0.55) * rng.integers(1, 4, (32, 32, 32))
        elif kind == 1:
            ids[i] = 0
        elif kind == 2:
            ids[i] = 1 + (i % 3)
        else:
            ids[i] = (rng.random((32, 32, 32)) < 0.02) * 2
    ids = ids.reshape(n, -1)
    nbr = ctx.build_neighbor_index(ctx_pts := hb.context.as_chunk_pts(pts))
    ref = ctx.mesh(pts, ids, nbr)
    span = int(ref["instances"].shape[0])
    assert span > 0

    full = _dev_mesh(ctx, torch, ctx_pts, ids, nbr, span)
    assert int(full["total"][0]) == span and int(full["total"][2]) == 0
    assert np.array_equal(full["offsets"], ref["offsets"]) and np.array_equal(full["counts"], ref["counts"])
    assert np.array_equal(full["out"][:span * 100], ref["instances"].view(np.uint8).reshape(-1))
    assert (full["out"][span * 100:] == 0xA5).all()

    # room for about half of the quads: the chunks that fit are written exactly, the others are skipped and flagged
    cap = span // 2
    part = _dev_mesh(ctx, torch, ctx_pts, ids, nbr, cap)
    assert int(part["total"][0]) == span and int(part["total"][2]) & 1
    assert (part["out"][cap * 100:] == 0xA5).all()
    want = ref["instances"].view(np.uint8).reshape(-1)
    for i in range(n):
        o, c = int(ref["offsets"][i]), int(ref["counts"][i])
        if c and o + ((c + 3) & ~3) <= cap:
            assert np.array_equal(part["out"][o * 100:(o + c) * 100], want[o * 100:(o + c) * 100])
