"""World: the device-resident ChunkMap (hb_world_*), numpy-typed 1:1 wrappers.

Mirrors the reference's runtime loop on both sides of the wire:

    server  ChunkMap::queue_load / load_provided   -> World.load          (server/src/chunk/map.rs:69-75,203-228)
            ChunkMap::set_cube                     -> World.set_cubes     (map.rs:81-151)
            Chunk::sync_with_client                -> World.sync / fetch_updates   (server/src/chunk/mod.rs:35-67)
    client  ChunkMap::update (remesh what changed) -> World.remesh        (client/src/world/chunk.rs:47-76)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import CHUNK_CUBE_DTYPE, CHUNK_PT_DTYPE, CHUNK_VOLUME, CUBE_DTYPE, INSTANCE_DTYPE, ChunkPtC, check, ptr
from .context import Context, as_chunk_pts


def vec3u5_unpack(pos: np.ndarray):
    """vec3u5 bits -> (x, y, z) (lib/src/vector.rs:1181-1193)."""
    p = np.asarray(pos, dtype=np.uint16).astype(np.int64)
    return (p >> 11) & 31, (p >> 6) & 31, (p >> 1) & 31


class World:
    def __init__(self, ctx: Context, capacity: int):
        self._lib = _lib.load()
        self.ctx = ctx
        self.capacity = int(capacity)
        self._h = C.c_void_p()
        check(self._lib.hb_world_create(ctx._h, self.capacity, C.byref(self._h)))

    def __len__(self) -> int:
        return int(self._lib.hb_world_len(self._h))

    def load(self, positions) -> None:
        pts = as_chunk_pts(positions)
        check(self._lib.hb_world_load(self._h, ptr(pts), pts.shape[0]))

    def unload(self, positions) -> None:
        pts = as_chunk_pts(positions)
        check(self._lib.hb_world_unload(self._h, ptr(pts), pts.shape[0]))

    def set_cubes(self, cubes_xyz, materials) -> None:
        """Edits applied in order; cubes_xyz (n, 3) world cube coordinates, materials 0 = dig."""
        xyz = np.ascontiguousarray(cubes_xyz, dtype=np.int32).reshape(-1, 3)
        mats = np.ascontiguousarray(materials, dtype=np.uint16).reshape(-1)
        if mats.shape[0] != xyz.shape[0]:
            raise ValueError("one material per edit")
        check(self._lib.hb_world_set_cubes(self._h, ptr(xyz), ptr(mats), xyz.shape[0]))

    def set_cube(self, position, material: int) -> None:
        self.set_cubes([position], [material])

    def sync(self) -> tuple:
        """Returns (dirty chunks, CubeUpdate entries) drained by this sync."""
        n, e = C.c_size_t(0), C.c_uint64(0)
        check(self._lib.hb_world_sync(self._h, C.byref(n), C.byref(e)))
        self._last = (n.value, e.value)
        return self._last

    def fetch_updates(self) -> dict:
        """{"pts", "offsets", "counts", "entries" (CHUNK_CUBE_DTYPE)} of the last sync()."""
        n, e = getattr(self, "_last", (0, 0))
        pts = np.zeros(n, dtype=CHUNK_PT_DTYPE)
        offsets = np.zeros(n, dtype=np.uint64)
        counts = np.zeros(n, dtype=np.uint32)
        # the entries land in pinned memory (hb_host_alloc): one DMA at link speed instead of the pageable path.  The
        # buffer is kept and only ever grows (pinning 743 MB costs ~300 ms, the copy itself ~15 ms), so the returned
        # "entries" view is valid until the next fetch_updates() on this world.
        buf = getattr(self, "_pinned_entries", None)
        if buf is None or buf.shape[0] < e:
            self._pinned_entries = buf = self.ctx.pinned_empty(max(e + e // 2, 1 << 16), CHUNK_CUBE_DTYPE)
        entries = buf[:e]
        nr, ne = C.c_size_t(0), C.c_uint64(0)
        check(self._lib.hb_world_fetch_updates(self._h, ptr(pts), n, ptr(offsets), ptr(counts), ptr(entries), e,
                                               C.byref(nr), C.byref(ne)))
        # chunks unloaded since the sync are skipped: only the first nr runs / ne entries were written
        k = nr.value
        return {"pts": pts[:k], "offsets": offsets[:k], "counts": counts[:k], "entries": entries[:ne.value]}

    def remesh(self) -> dict:
        """{"pts", "offsets", "counts", "instances"} for every chunk that received a CubeUpdate since the last remesh.
        The instances land in a pinned buffer the World keeps (grow-only), so the arrays are views that stay valid until
        the next remesh(); with enough room the meshes are produced and copied in ONE pass (the two-call sizing pattern
        of hb_world_remesh is only taken when the buffer has to grow)."""
        cap = self.capacity
        pts = np.zeros(cap, dtype=CHUNK_PT_DTYPE)
        offsets = np.zeros(cap, dtype=np.uint64)
        counts = np.zeros(cap, dtype=np.uint32)
        buf = getattr(self, "_pinned_inst", None)
        if buf is None:
            self._pinned_inst = buf = self.ctx.pinned_empty(1 << 18, INSTANCE_DTYPE)  # 26 MB to start with
        n, total = C.c_size_t(0), C.c_uint64(0)
        rc = self._lib.hb_world_remesh(self._h, ptr(pts), cap, ptr(buf), buf.shape[0], ptr(offsets), ptr(counts),
                                       C.byref(n), C.byref(total))
        if rc == _lib.HB_ERR_CAPACITY:
            self._pinned_inst = buf = self.ctx.pinned_empty(int(total.value) + int(total.value) // 2 + 4, INSTANCE_DTYPE)
            rc = self._lib.hb_world_remesh(self._h, ptr(pts), cap, ptr(buf), buf.shape[0], ptr(offsets), ptr(counts),
                                           C.byref(n), C.byref(total))
        check(rc)
        k = n.value
        return {"pts": pts[:k], "offsets": offsets[:k], "counts": counts[:k], "instances": buf[:total.value]}

    def reserve_remesh(self, quads: int) -> None:
        """Size the pinned instance buffer of remesh() up front (pinning memory costs milliseconds per MB; a host sizes
        it once at start-up instead of on the first large remesh)."""
        buf = getattr(self, "_pinned_inst", None)
        if buf is None or buf.shape[0] < quads:
            self._pinned_inst = self.ctx.pinned_empty(int(quads) + 4, INSTANCE_DTYPE)

    def remesh_dev(self) -> dict:
        """Like remesh() but the instances stay in HBM: adds "d_instances" (raw device pointer) and "span"."""
        n, total = C.c_size_t(0), C.c_uint64(0)
        d = C.c_void_p()
        cap = self.capacity
        pts = np.zeros(cap, dtype=CHUNK_PT_DTYPE)
        offsets = np.zeros(cap, dtype=np.uint64)
        counts = np.zeros(cap, dtype=np.uint32)
        check(self._lib.hb_world_remesh_dev(self._h, ptr(pts), cap, ptr(offsets), ptr(counts), C.byref(n), C.byref(total),
                                            C.byref(d)))
        k = n.value
        return {"pts": pts[:k], "offsets": offsets[:k], "counts": counts[:k], "span": total.value, "d_instances": d.value}

    def read_chunk(self, position) -> np.ndarray:
        x, y, z = (int(v) for v in position)
        out = np.zeros(CHUNK_VOLUME, dtype=CUBE_DTYPE)
        check(self._lib.hb_world_read_chunk(self._h, ChunkPtC(x, y, z), ptr(out)))
        return out

    def close(self) -> None:
        if self._h:
            self._lib.hb_world_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ServerChunkMap:
    """The reference's names over a World: server ChunkMap (server/src/chunk/map.rs) + Chunk::sync_with_client
    (server/src/chunk/mod.rs:35-67) + the client's remesh of updated chunks (client/src/world/chunk.rs:47-76).

        queue_load / queue_unload   map.rs:65-75   (collected; flushed by update())
        update()                    map.rs:236-245 (load_provided, then unload_requested)
        set_cube(position, material) map.rs:81-151 (applied immediately, in call order)
        sync_with_client()          -> {chunk position: CHUNK_CUBE_DTYPE entries}   (CubeUpdate per chunk)
        remesh()                    -> {chunk position: INSTANCE_DTYPE array} for every chunk that got a CubeUpdate
    """

    def __init__(self, ctx: Context, capacity: int):
        self.world = World(ctx, capacity)
        self._load, self._unload = [], []

    def queue_load(self, position) -> None:
        self._load.append(tuple(int(v) for v in position))

    def queue_unload(self, position) -> None:
        self._unload.append(tuple(int(v) for v in position))

    def update(self) -> None:
        load, self._load = self._load, []
        unload, self._unload = self._unload, []
        if load:
            self.world.load(load)
        if unload:
            self.world.unload(unload)

    def set_cube(self, position, material: int) -> None:
        self.world.set_cube(position, material)

    def __len__(self) -> int:
        return len(self.world)

    def sync_with_client(self, want_wire: bool = True) -> dict:
        n, _ = self.world.sync()
        if not want_wire or n == 0:
            return {}
        u = self.world.fetch_updates()
        return {(int(p["x"]), int(p["y"]), int(p["z"])): u["entries"][int(o):int(o) + int(c)]
                for p, o, c in zip(u["pts"], u["offsets"], u["counts"])}

    def remesh(self) -> dict:
        out = self.world.remesh()  # views of the World's pinned buffer: copied, the mirror hands out owned meshes
        return {(int(p["x"]), int(p["y"]), int(p["z"])): out["instances"][int(o):int(o) + int(c)].copy()
                for p, o, c in zip(out["pts"], out["offsets"], out["counts"])}

    def close(self) -> None:
        self.world.close()
