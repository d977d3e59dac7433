"""Multi-GPU host logic: chunks are independent, so a region is partitioned into contiguous x-slabs of
chunk columns (all cy of a column stay together, sharing one noise tile); no collective touches the
data path.  torch.distributed is used only for the barrier / max-over-ranks timing / totals.

The reference has no distributed layer (its "server" is an in-process thread); the unit of
work it hands to its thread pool is one chunk (server/src/generator/mod.rs:36), which is what makes
this partition legal.
"""
from __future__ import annotations

from typing import Tuple

from ._lib import Region


def slab_bounds(ncx: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ix in [0, ncx) over `world` ranks: returns [ix_begin, ix_end)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(ncx, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def chunk_range(region: Region, ix_begin: int, ix_end: int) -> Tuple[int, int]:
    """Region chunk indices ((ix*ncz)+iz)*ncy+iy covered by the slab: a contiguous range."""
    per_ix = region.ncz * region.ncy
    return ix_begin * per_ix, ix_end * per_ix


def weak_scaling_region(columns: int, rank: int, world: int, cy0: int = -3, ncy: int = 6) -> Region:
    """The bench's weak-scaling layout: every rank owns its own columns x columns block, side by side in x,
    of a (columns*world) x columns region centred on the origin."""
    cx0 = -(columns * world) // 2 + rank * columns
    return Region(cx0, -(columns // 2), columns, columns, cy0, ncy)


def reduce_max(x: float, device=None) -> float:
    """max over ranks (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
