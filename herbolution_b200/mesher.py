"""Host-side mirror of the reference's client mesher (client/src/world/chunk.rs).

    generate_mesh(shell, instances)        # chunk.rs:176  -- one chunk, 27-entry shell
    ChunkMap.update()                      # chunk.rs:47-76 -- remesh every dirty chunk, batched

`ChunkData` is the client's copy of a chunk (chunk.rs:97-102); the face flags the reference reads
from `cube.flags` (maintained server-side by CubeMesh::set / cull_shared_faces) are recomputed on
the GPU from the block ids, which is equivalent once loading has settled.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from ._lib import CHUNK_VOLUME, CUBE_DTYPE, INSTANCE_DTYPE
from .context import Context
from .generator import ChunkPt


@dataclass
class ChunkData:
    """client/src/world/chunk.rs:97-102.  `cubes` is either u16[32768] material ids (faces are recomputed) or
    CUBE_DTYPE[32768] PaletteCubes as received from the server (faces are read from cube.flags: the literal
    generate_mesh, needed on the incremental dig/place path)."""
    position: ChunkPt
    cubes: np.ndarray  # index L = x*1024 + z*32 + y

    @property
    def has_flags(self) -> bool:
        return self.cubes.dtype == CUBE_DTYPE


def create_chunk_shell(chunks: Dict[ChunkPt, ChunkData], position: ChunkPt) -> List[Optional[ChunkData]]:
    """create_chunk_shell (chunk.rs:158-174): x outer, y, z inner; centre = index 13."""
    shell = []
    for x in (-1, 0, 1):
        for y in (-1, 0, 1):
            for z in (-1, 0, 1):
                shell.append(chunks.get(ChunkPt(position.x + x, position.y + y, position.z + z)))
    return shell


def generate_mesh(ctx: Context, shell: Sequence[Optional[ChunkData]], instances: list) -> None:
    """generate_mesh(shell, instances) (chunk.rs:176-223): appends the centre chunk's Instance3d
    records (one INSTANCE_DTYPE array) to `instances`."""
    center = shell[13]
    if center is None:
        raise ValueError("shell[13] (the centre chunk) must be present")  # the reference unwrap()s
    present = [(i, c) for i, c in enumerate(shell) if c is not None]
    pts = np.array([c.position for _, c in present], dtype=np.int64)
    if center.has_flags:
        if not all(c.has_flags for _, c in present):
            raise ValueError("either every chunk of the shell carries PaletteCubes or none does")
        out = ctx.mesh_cubes(pts, np.stack([c.cubes.reshape(CHUNK_VOLUME) for _, c in present]))
    else:
        ids = np.stack([np.asarray(c.cubes, dtype=np.uint16).reshape(CHUNK_VOLUME) for _, c in present])
        out = ctx.mesh(pts, ids)
    k = [i for i, _ in present].index(13)
    o, n = int(out["offsets"][k]), int(out["counts"][k])
    instances.append(out["instances"][o:o + n].copy())


class ChunkMap:
    """client ChunkMap (chunk.rs:32-83): holds ChunkData and remeshes dirty chunks in one batch."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.map: Dict[ChunkPt, ChunkData] = {}
        self.meshes: Dict[ChunkPt, np.ndarray] = {}
        self._dirty: set = set()

    def insert(self, position, cubes: np.ndarray) -> None:
        p = ChunkPt(*map(int, position))
        cubes = np.asarray(cubes)
        if cubes.dtype != CUBE_DTYPE:
            cubes = cubes.astype(np.uint16)
        self.map[p] = ChunkData(p, cubes.reshape(CHUNK_VOLUME))
        self._dirty.add(p)

    def update(self) -> int:
        """Remesh every dirty chunk (the reference spawns one task per chunk, chunk.rs:66-75); neighbours
        are taken from the whole map, absent ones are absent.  Returns the number of chunks remeshed."""
        if not self._dirty:
            return 0
        pts_list = list(self.map.keys())
        pts = np.array(pts_list, dtype=np.int64)
        blocks = np.stack([self.map[p].cubes for p in pts_list])
        out = self.ctx.mesh_cubes(pts, blocks) if blocks.dtype == CUBE_DTYPE else self.ctx.mesh(pts, blocks)
        for i, p in enumerate(pts_list):
            if p in self._dirty:
                o, n = int(out["offsets"][i]), int(out["counts"][i])
                self.meshes[p] = out["instances"][o:o + n].copy()
        done = len(self._dirty)
        self._dirty.clear()
        return done


def empty_instances() -> np.ndarray:
    return np.zeros(0, dtype=INSTANCE_DTYPE)
