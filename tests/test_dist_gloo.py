"""world_size-2 CPU test (gloo) of the multi-GPU host logic: x-slab partition of a region, contiguous
chunk ranges per rank, weak-scaling regions that tile the global region, and the max/sum reductions
bench.py uses for timing.  There is no data-path collective to test: chunks are independent."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import herbolution_b200 as hb
    from herbolution_b200 import sharding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        region = hb.Region(-5, 3, 7, 4, -3, 6)  # 7 columns in x over 2 ranks: 4 + 3
        a, b = sharding.slab_bounds(region.ncx, world, rank)
        lo, hi = sharding.chunk_range(region, a, b)
        t = torch.tensor([a, b, lo, hi], dtype=torch.int64)
        gathered = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, t)
        bounds = [tuple(int(v) for v in g) for g in gathered]
        # slabs are contiguous, ordered, cover [0, ncx) exactly once; chunk ranges likewise
        assert bounds[0][0] == 0 and bounds[-1][1] == region.ncx
        assert bounds[0][2] == 0 and bounds[-1][3] == region.n_chunks
        for p, n in zip(bounds, bounds[1:]):
            assert p[1] == n[0] and p[3] == n[2]
        assert max(bb - aa for aa, bb, _, _ in bounds) - min(bb - aa for aa, bb, _, _ in bounds) <= 1
        # the slab's chunk points are the region's points restricted to ix in [a, b)
        pts = region.chunk_points()
        assert ((pts["x"][lo:hi] >= region.cx0 + a) & (pts["x"][lo:hi] < region.cx0 + b)).all()
        # weak-scaling regions sit side by side in x and tile the global region
        r = sharding.weak_scaling_region(8, rank, world)
        t2 = torch.tensor([r.cx0, r.cx0 + r.ncx], dtype=torch.int64)
        g2 = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(g2, t2)
        assert int(g2[0][1]) == int(g2[1][0]) and int(g2[0][0]) == -8 and int(g2[1][1]) == 8
        # timing / totals reductions
        assert sharding.reduce_max(1.0 + rank) == float(world)
        assert sharding.reduce_sum(float(hi - lo)) == float(region.n_chunks)
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_and_reductions():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_slab_bounds_properties():
    sys.path.insert(0, ROOT)
    from herbolution_b200 import sharding
    for ncx in (1, 2, 7, 256, 1024):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharding.slab_bounds(ncx, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == ncx
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.slab_bounds(4, 2, 2)
