"""Context: one hb_ctx (one GPU) with numpy-typed wrappers over the C ABI.

These are 1:1 wrappers (same argument meaning, same error behaviour as include/herbgpu.h); the
reference-shaped interface (ChunkGenerator, GenerationParams, generate_mesh) lives in
generator.py / mesher.py on top of this.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import (CHUNK_AREA, CHUNK_PT_DTYPE, CHUNK_VOLUME, CUBE_DTYPE, INSTANCE_DTYPE, MAX_COLORS, HerbGpuError,
                   Region, check, ptr)


def as_chunk_pts(positions) -> np.ndarray:
    """Accepts a CHUNK_PT_DTYPE array or an (n, 3) integer array / list of (x, y, z)."""
    a = np.asarray(positions)
    if a.dtype == CHUNK_PT_DTYPE:
        return np.ascontiguousarray(a)
    a = np.asarray(positions, dtype=np.int64).reshape(-1, 3)
    if a.size and (np.abs(a) > 2**31 - 1).any():
        raise HerbGpuError(_lib.HB_ERR_RANGE, "chunk coordinate does not fit i32")
    out = np.empty(a.shape[0], dtype=CHUNK_PT_DTYPE)
    out["x"], out["y"], out["z"] = a[:, 0], a[:, 1], a[:, 2]
    return out


class RegionPlan:
    """Device-resident fused gen+mesh pipeline for a slab of a region (hb_region_plan)."""

    def __init__(self, ctx: "Context", region: Region, ix_begin: int = 0, ix_end: Optional[int] = None,
                 quad_capacity: int = 0, want_ids: bool = False):
        self._lib = _lib.load()
        self.ctx = ctx
        self.region = region
        self._h = C.c_void_p()
        ix_end = region.ncx if ix_end is None else ix_end
        check(self._lib.hb_region_plan_create(ctx._h, C.byref(region), ix_begin, ix_end, quad_capacity, int(want_ids),
                                              C.byref(self._h)))

    def enqueue(self, stages: int = _lib.STAGE_ALL_MESH, stream: int = 0) -> None:
        check(self._lib.hb_region_plan_enqueue(self._h, stages, C.c_void_p(stream or None)))

    def launches(self, stages: int = _lib.STAGE_ALL_MESH) -> int:
        return self._lib.hb_region_plan_launches(self._h, stages)

    def result(self, stream: int = 0) -> dict:
        """Synchronises the stream; returns totals and raw device pointers."""
        n, q, span = C.c_uint64(), C.c_uint64(), C.c_uint64()
        p = [C.c_void_p() for _ in range(5)]
        check(self._lib.hb_region_plan_result(self._h, C.c_void_p(stream or None), C.byref(n), C.byref(q), C.byref(span),
                                              *[C.byref(x) for x in p]))
        return {"n_chunks": n.value, "quads": q.value, "span": span.value, "d_instances": p[0].value,
                "d_offsets": p[1].value, "d_counts": p[2].value, "d_ids": p[3].value, "d_heights": p[4].value}

    def packed_ptr(self) -> int:
        """Raw device pointer of the hb_packed_quad words written by STAGE_EMIT_PACKED (0 before the first one)."""
        p = C.c_void_p()
        check(self._lib.hb_region_plan_packed(self._h, C.byref(p)))
        return p.value or 0

    def close(self) -> None:
        if self._h:
            self._lib.hb_region_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _region_sink_cb(sink):
    """ctypes trampoline for hb_region_sink -> sink(first_chunk, offsets, counts, instances, ids or None)."""

    def _cb(_user, first, n, p_off, p_cnt, p_inst, span, p_ids):
        if sink is None:
            return
        off = np.frombuffer((C.c_uint64 * n).from_address(p_off), dtype=np.uint64) if n else np.zeros(0, np.uint64)
        cnt = np.frombuffer((C.c_uint32 * n).from_address(p_cnt), dtype=np.uint32) if n else np.zeros(0, np.uint32)
        inst = np.frombuffer((C.c_char * (span * 100)).from_address(p_inst), dtype=INSTANCE_DTYPE) if span else \
            np.zeros(0, INSTANCE_DTYPE)
        ids = None
        if p_ids:
            ids = np.frombuffer((C.c_uint16 * (n * CHUNK_VOLUME)).from_address(p_ids), dtype=np.uint16)
            ids = ids.reshape(n, CHUNK_VOLUME)
        sink(first, off, cnt, inst, ids)

    return _lib.REGION_SINK(_cb)


class MultiContext:
    """hb_multi: one process driving several GPUs behind one handle (include/herbgpu.h).  devices=None takes every
    visible device.  The region call splits the x range into one slab per device; the sink is serialised."""

    def __init__(self, devices=None, seed: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if devices is None:
            check(self._lib.hb_multi_create(None, 0, C.byref(self._h)))
        else:
            arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            check(self._lib.hb_multi_create(arr, len(devices), C.byref(self._h)))
        self.n_devices = int(self._lib.hb_multi_device_count(self._h))
        self.set_world(seed)

    def set_world(self, seed: int, freq: float = 0.001, octaves: int = 6, lacunarity: float = 2.0, gain: float = 0.6) -> None:
        check(self._lib.hb_multi_set_world(self._h, seed, freq, octaves, lacunarity, gain))
        self.seed = seed

    def set_noise_kind(self, kind: int) -> None:
        check(self._lib.hb_multi_set_noise_kind(self._h, int(kind)))

    def set_fma_mode(self, fused: bool) -> None:
        check(self._lib.hb_multi_set_fma_mode(self._h, 1 if fused else 0))

    def set_region_transport(self, transport: int) -> None:
        check(self._lib.hb_multi_set_region_transport(self._h, int(transport)))

    def set_host_threads(self, n_per_device: int) -> None:
        check(self._lib.hb_multi_set_host_threads(self._h, int(n_per_device)))

    def slab(self, region: Region, i: int):
        lo, hi = C.c_int32(0), C.c_int32(0)
        check(self._lib.hb_multi_slab(self._h, C.byref(region), i, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def generate_mesh_region(self, region: Region, want_ids: bool = False, sink: Optional[Callable[..., None]] = None) -> dict:
        cb = _region_sink_cb(sink)
        tc, tq = C.c_uint64(0), C.c_uint64(0)
        check(self._lib.hb_multi_generate_mesh_region(self._h, C.byref(region), int(want_ids), cb, None, C.byref(tc), C.byref(tq)))
        return {"chunks": tc.value, "quads": tq.value}

    def close(self) -> None:
        if self._h:
            self._lib.hb_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Context:
    def __init__(self, device: int = 0, seed: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.hb_create(device, C.byref(self._h)))
        self.device = device
        self.set_world(seed)

    # -- parameters ------------------------------------------------------------------------
    def set_world(self, seed: int, freq: float = 0.001, octaves: int = 6, lacunarity: float = 2.0,
                  gain: float = 0.6) -> None:
        """Defaults = server/src/generator/mod.rs:103-108."""
        check(self._lib.hb_set_world(self._h, seed, freq, octaves, lacunarity, gain))
        self.seed = seed

    def set_noise_kind(self, kind: int) -> None:
        """_lib.NOISE_FBM (the game's, default) / NOISE_RIDGE / NOISE_TURBULENCE / NOISE_GRADIENT."""
        check(self._lib.hb_set_noise_kind(self._h, int(kind)))

    def set_fma_mode(self, fused: bool) -> None:
        """True (default): neg_mul_add fused as on AVX2+FMA / NEON hosts; False: the SSE2 / SSE4.1 / scalar column
        (crates/simd-noise/src/simd/ops/f32.rs:141-162)."""
        check(self._lib.hb_set_fma_mode(self._h, 1 if fused else 0))

    def set_materials(self, colors: list) -> None:
        """colors[m] = list of rgba tuples of material id m+1 (server/src/chunk/material.rs:27-61)."""
        n = len(colors)
        ncol = np.array([len(c) for c in colors], dtype=np.int32)
        rgba = np.zeros((n, MAX_COLORS, 4), dtype=np.float32)
        for i, cs in enumerate(colors):
            if len(cs) > MAX_COLORS:
                raise HerbGpuError(_lib.HB_ERR_INVALID, f"material {i + 1}: more than {MAX_COLORS} colours")
            for j, c in enumerate(cs):
                rgba[i, j] = c
        check(self._lib.hb_set_materials(self._h, n, ptr(ncol), ptr(rgba)))

    # -- generation ------------------------------------------------------------------------
    def noise_tiles(self, cols_xz, want_heights: bool = False):
        cols = np.ascontiguousarray(np.asarray(cols_xz, dtype=np.int32).reshape(-1, 2))
        n = cols.shape[0]
        noise = np.empty((n, CHUNK_AREA), dtype=np.float32)
        heights = np.empty((n, CHUNK_AREA), dtype=np.int32) if want_heights else None
        check(self._lib.hb_noise_tiles(self._h, ptr(cols), n, ptr(noise), ptr(heights)))
        return (noise, heights) if want_heights else noise

    def noise_fbm3d(self, start, extent, freq) -> np.ndarray:
        nx, ny, nz = (int(v) for v in extent)
        f = np.asarray(freq, dtype=np.float32).reshape(3)
        out = np.empty(max(nx, 0) * max(ny, 0) * max(nz, 0), dtype=np.float32)
        check(self._lib.hb_noise_fbm3d(self._h, float(start[0]), float(start[1]), float(start[2]), nx, ny, nz, ptr(f),
                                       ptr(out)))
        return out.reshape(nz, ny, nx)

    def generate(self, positions, want_cubes: bool = False, want_noise: bool = False) -> dict:
        pts = as_chunk_pts(positions)
        n = pts.shape[0]
        ids = np.empty((n, CHUNK_VOLUME), dtype=np.uint16)
        cubes = np.empty((n, CHUNK_VOLUME), dtype=CUBE_DTYPE) if want_cubes else None
        noise = np.empty((n, CHUNK_AREA), dtype=np.float32) if want_noise else None
        check(self._lib.hb_generate(self._h, ptr(pts), n, ptr(ids), ptr(cubes), ptr(noise)))
        return {"ids": ids, "cubes": cubes, "noise": noise}

    # -- meshing ---------------------------------------------------------------------------
    def build_neighbor_index(self, positions) -> np.ndarray:
        pts = as_chunk_pts(positions)
        out = np.empty((pts.shape[0], 27), dtype=np.int32)
        check(self._lib.hb_build_neighbor_index(ptr(pts), pts.shape[0], ptr(out)))
        return out

    def mesh_cubes(self, positions, cubes: np.ndarray, neighbor_index: Optional[np.ndarray] = None) -> dict:
        """hb_mesh_cubes: literal generate_mesh, visible faces read from cube.flags (CUBE_DTYPE arrays)."""
        return self.mesh(positions, cubes, neighbor_index, _cubes=True)

    def mesh(self, positions, ids: np.ndarray, neighbor_index: Optional[np.ndarray] = None, _cubes: bool = False) -> dict:
        """Returns {"instances": INSTANCE_DTYPE[span], "offsets": u64[n], "counts": u32[n]}."""
        pts = as_chunk_pts(positions)
        n = pts.shape[0]
        fn = self._lib.hb_mesh_cubes if _cubes else self._lib.hb_mesh
        ids = np.ascontiguousarray(ids, dtype=CUBE_DTYPE if _cubes else np.uint16).reshape(n, CHUNK_VOLUME)
        nbr = self.build_neighbor_index(pts) if neighbor_index is None else \
            np.ascontiguousarray(neighbor_index, dtype=np.int32).reshape(n, 27)
        offsets = np.zeros(n, dtype=np.uint64)
        counts = np.zeros(n, dtype=np.uint32)
        total = C.c_uint64(0)
        rc = fn(self._h, ptr(pts), n, ptr(ids), ptr(nbr), None, 0, ptr(offsets), ptr(counts), C.byref(total))
        if rc not in (_lib.HB_OK, _lib.HB_ERR_CAPACITY):
            check(rc)
        inst = np.zeros(total.value, dtype=INSTANCE_DTYPE)
        if total.value:
            check(fn(self._h, ptr(pts), n, ptr(ids), ptr(nbr), ptr(inst), total.value, ptr(offsets), ptr(counts),
                     C.byref(total)))
        return {"instances": inst, "offsets": offsets, "counts": counts}

    # -- on-disk cube codec (server/src/chunk/codec.rs) ------------------------------------
    def rle_encode(self, ids: np.ndarray) -> dict:
        """encode_cubes for n chunks: {"bytes": u8[span], "offsets": u64[n], "sizes": u32[n]}."""
        ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(-1, CHUNK_VOLUME)
        n = ids.shape[0]
        offsets = np.zeros(n, dtype=np.uint64)
        sizes = np.zeros(n, dtype=np.uint32)
        total = C.c_uint64(0)
        rc = self._lib.hb_rle_encode(self._h, ptr(ids), n, None, 0, ptr(offsets), ptr(sizes), C.byref(total))
        if rc not in (_lib.HB_OK, _lib.HB_ERR_CAPACITY):
            check(rc)
        out = np.zeros(total.value, dtype=np.uint8)
        if total.value:
            check(self._lib.hb_rle_encode(self._h, ptr(ids), n, ptr(out), total.value, ptr(offsets), ptr(sizes),
                                          C.byref(total)))
        return {"bytes": out, "offsets": offsets, "sizes": sizes}

    def rle_decode(self, data: np.ndarray, offsets, sizes) -> np.ndarray:
        """CubeGrid::decode record loop for n streams -> u16[n][32768] raw ids."""
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64).reshape(-1)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32).reshape(-1)
        n = offsets.shape[0]
        out = np.zeros((n, CHUNK_VOLUME), dtype=np.uint16)
        check(self._lib.hb_rle_decode(self._h, ptr(data), data.shape[0], ptr(offsets), ptr(sizes), n, ptr(out)))
        return out

    # -- fused region ----------------------------------------------------------------------
    def generate_mesh_region(self, region: Region, want_ids: bool = False,
                             sink: Optional[Callable[..., None]] = None, ix_begin: int = 0,
                             ix_end: Optional[int] = None) -> dict:
        """sink(first_chunk, offsets u64[n], counts u32[n], instances INSTANCE_DTYPE[span], ids or None);
        the arrays are views of pinned library buffers, valid only during the call."""

        def _cb(_user, first, n, p_off, p_cnt, p_inst, span, p_ids):
            if sink is None:
                return
            off = np.frombuffer((C.c_uint64 * n).from_address(p_off), dtype=np.uint64) if n else np.zeros(0, np.uint64)
            cnt = np.frombuffer((C.c_uint32 * n).from_address(p_cnt), dtype=np.uint32) if n else np.zeros(0, np.uint32)
            inst = np.frombuffer((C.c_char * (span * 100)).from_address(p_inst), dtype=INSTANCE_DTYPE) if span else \
                np.zeros(0, INSTANCE_DTYPE)
            ids = None
            if p_ids:
                ids = np.frombuffer((C.c_uint16 * (n * CHUNK_VOLUME)).from_address(p_ids), dtype=np.uint16)
                ids = ids.reshape(n, CHUNK_VOLUME)
            sink(first, off, cnt, inst, ids)

        cb = _lib.REGION_SINK(_cb)
        tc, tq = C.c_uint64(0), C.c_uint64(0)
        ix_end = region.ncx if ix_end is None else ix_end
        check(self._lib.hb_generate_mesh_region_slab(self._h, C.byref(region), ix_begin, ix_end, int(want_ids), cb, None,
                                                     C.byref(tc), C.byref(tq)))
        return {"chunks": tc.value, "quads": tq.value}

    def set_region_transport(self, transport: int) -> None:
        """_lib.TRANSPORT_PACKED (default: 4 B/quad over the link, host threads rebuild the records) or
        TRANSPORT_INSTANCES (100 B/quad copied as they are)."""
        check(self._lib.hb_set_region_transport(self._h, int(transport)))

    def set_host_threads(self, n: int) -> None:
        check(self._lib.hb_set_host_threads(self._h, int(n)))

    def generate_mesh_region_packed(self, region: Region, want_ids: bool = False,
                                    sink: Optional[Callable[..., None]] = None, ix_begin: int = 0,
                                    ix_end: Optional[int] = None) -> dict:
        """Like generate_mesh_region, but sink(first_chunk, offsets, counts, packed u32[span], ids or None) receives
        the hb_packed_quad words; expand them with expand_quads."""

        def _cb(_user, first, n, p_off, p_cnt, p_q, span, p_ids):
            if sink is None:
                return
            off = np.frombuffer((C.c_uint64 * n).from_address(p_off), dtype=np.uint64) if n else np.zeros(0, np.uint64)
            cnt = np.frombuffer((C.c_uint32 * n).from_address(p_cnt), dtype=np.uint32) if n else np.zeros(0, np.uint32)
            q = np.frombuffer((C.c_uint32 * span).from_address(p_q), dtype=np.uint32) if span else np.zeros(0, np.uint32)
            ids = None
            if p_ids:
                ids = np.frombuffer((C.c_uint16 * (n * CHUNK_VOLUME)).from_address(p_ids), dtype=np.uint16)
                ids = ids.reshape(n, CHUNK_VOLUME)
            sink(first, off, cnt, q, ids)

        cb = _lib.REGION_SINK(_cb)
        tc, tq = C.c_uint64(0), C.c_uint64(0)
        ix_end = region.ncx if ix_end is None else ix_end
        check(self._lib.hb_generate_mesh_region_packed(self._h, C.byref(region), ix_begin, ix_end, int(want_ids), cb, None,
                                                       C.byref(tc), C.byref(tq)))
        return {"chunks": tc.value, "quads": tq.value}

    def expand_quads(self, position, packed: np.ndarray) -> np.ndarray:
        """hb_expand_quads: the Instance3d records of one chunk's packed quads."""
        x, y, z = (int(v) for v in position)
        packed = np.ascontiguousarray(packed, dtype=np.uint32).reshape(-1)
        n = packed.shape[0]
        raw = np.zeros(((n + 3) & ~3) * 100 + 16, dtype=np.uint8)
        o = (-raw.ctypes.data) & 15
        out = raw[o:o + ((n + 3) & ~3) * 100].view(INSTANCE_DTYPE)
        check(self._lib.hb_expand_quads(self._h, _lib.ChunkPtC(x, y, z), ptr(packed), n, C.c_void_p(out.ctypes.data)))
        return out[:n]

    def encode_palette(self, descs=None) -> bytes:
        """encode_palette + Material::encode (server/src/chunk/codec.rs:70-75, material.rs:73-99) for the ctx's
        materials; descs = [(group, key, cullable_faces, has_collider, toughness), ...] or None for the game's three."""
        arr = None
        if descs is not None:
            arr = (_lib.MaterialDesc * len(descs))(*[_lib.MaterialDesc(g.encode(), k.encode(), int(cf), int(bool(hc)), float(t))
                                                     for g, k, cf, hc, t in descs])
        n = C.c_size_t(0)
        rc = self._lib.hb_encode_palette(self._h, arr, None, 0, C.byref(n))
        if rc not in (_lib.HB_OK, _lib.HB_ERR_CAPACITY):
            check(rc)
        out = np.zeros(max(1, n.value), dtype=np.uint8)
        check(self._lib.hb_encode_palette(self._h, arr, ptr(out), n.value, C.byref(n)))
        return out[:n.value].tobytes()

    def probe_d2h(self, nbytes: int = 1 << 30, reps: int = 4) -> float:
        """GB/s of plain device -> pinned host copies (the link ceiling of the end-to-end path)."""
        g = C.c_double(0.0)
        check(self._lib.hb_probe_d2h(self._h, nbytes, reps, C.byref(g)))
        return g.value

    def probe_host_write(self, nbytes: int = 1 << 30, reps: int = 4) -> float:
        """GB/s of the expander's worker pool writing host memory with non-temporal stores."""
        g = C.c_double(0.0)
        check(self._lib.hb_probe_host_write(self._h, nbytes, reps, C.byref(g)))
        return g.value

    def download(self, dev_ptr: int, dtype, count: int) -> np.ndarray:
        """Blocking device->host copy of `count` items at a raw device pointer (plan results)."""
        out = np.empty(count, dtype=dtype)
        if count:
            check(self._lib.hb_dev_download(self._h, ptr(out), C.c_void_p(dev_ptr), out.nbytes))
        return out

    def pinned_empty(self, count: int, dtype) -> np.ndarray:
        """Uninitialised array in pinned host memory (hb_host_alloc); freed when the array is garbage-collected."""
        dt = np.dtype(dtype)
        nbytes = max(1, count * dt.itemsize)
        p = C.c_void_p()
        check(self._lib.hb_host_alloc(self._h, nbytes, C.byref(p)))
        buf = (C.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt, count=count)
        lib = self._lib
        weakref.finalize(buf, lib.hb_host_free, None, C.c_void_p(p.value))
        return arr

    def region_plan(self, region: Region, **kw) -> RegionPlan:
        return RegionPlan(self, region, **kw)

    def close(self) -> None:
        if self._h:
            self._lib.hb_destroy(self._h)
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
