"""ctypes wrapper of the CPU oracle (oracle/hb_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU arm -- never by herbolution_b200."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from herbolution_b200._lib import CHUNK_AREA, CHUNK_VOLUME, CUBE_DTYPE, INSTANCE_DTYPE  # noqa: E402  (dtypes only)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from build_oracle import ORACLE_LIB, build_oracle  # noqa: E402


class World(C.Structure):
    _fields_ = [("seed", C.c_int64), ("freq", C.c_float), ("octaves", C.c_int32), ("lacunarity", C.c_float),
                ("gain", C.c_float), ("fused", C.c_int32), ("kind", C.c_int32)]


NOISE_FBM, NOISE_RIDGE, NOISE_TURBULENCE, NOISE_GRADIENT = 0, 1, 2, 3


def world(seed=0, freq=0.001, octaves=6, lacunarity=2.0, gain=0.6, fused=1, kind=NOISE_FBM) -> World:
    return World(seed, freq, octaves, lacunarity, gain, fused, kind)


class Material(C.Structure):
    _fields_ = [("n_colors", C.c_int32), ("rgba", (C.c_float * 4) * 8)]


class Palette(C.Structure):
    _fields_ = [("n_materials", C.c_int32), ("mat", Material * 16)]


class CubeMeshC(C.Structure):
    _fields_ = [("pos", C.c_int32 * 3), ("data", C.c_uint8 * (8 * CHUNK_VOLUME)), ("n_updated", C.c_uint64),
                ("upd", C.c_void_p), ("upd_cap", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.hbo_simplex2.restype = C.c_float
        L.hbo_simplex2.argtypes = [C.c_float, C.c_float, C.c_int64, C.c_int]
        L.hbo_fbm2.restype = C.c_float
        L.hbo_fbm2.argtypes = [C.c_float, C.c_float, C.POINTER(World)]
        L.hbo_noise_tile2.restype = None
        L.hbo_noise_tile2.argtypes = [C.POINTER(World), C.c_int32, C.c_int32, C.c_void_p]
        L.hbo_simplex3.restype = C.c_float
        L.hbo_simplex3.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int64]
        L.hbo_noise_tile3.restype = None
        L.hbo_noise_tile3.argtypes = [C.POINTER(World)] + [C.c_float] * 3 + [C.c_int] * 3 + [C.c_float] * 3 + [C.c_void_p]
        L.hbo_height.restype = C.c_int32
        L.hbo_height.argtypes = [C.c_float]
        L.hbo_map_create.restype = C.c_void_p
        L.hbo_map_create.argtypes = [C.POINTER(World)]
        L.hbo_map_destroy.restype = None
        L.hbo_map_destroy.argtypes = [C.c_void_p]
        for name in ("hbo_map_load", "hbo_map_find"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.hbo_map_unload.restype = None
        L.hbo_map_unload.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.hbo_map_len.restype = C.c_int
        L.hbo_map_len.argtypes = [C.c_void_p]
        L.hbo_map_set_cube.restype = None
        L.hbo_map_set_cube.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint16]
        L.hbo_map_cubes.restype = C.c_void_p
        L.hbo_map_cubes.argtypes = [C.c_void_p, C.c_int]
        L.hbo_map_position.restype = None
        L.hbo_map_position.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hbo_map_pending.restype = C.c_uint64
        L.hbo_map_pending.argtypes = [C.c_void_p, C.c_int]
        L.hbo_map_drain.restype = C.c_uint64
        L.hbo_map_drain.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.hbo_apply_updates.restype = None
        L.hbo_apply_updates.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.hbo_rle_encode.restype = C.c_size_t
        L.hbo_rle_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.hbo_rle_decode.restype = C.c_int
        L.hbo_rle_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.hbo_default_palette.restype = None
        L.hbo_default_palette.argtypes = [C.POINTER(Palette)]
        L.hbo_mesh_init.restype = None
        L.hbo_mesh_init.argtypes = [C.POINTER(CubeMeshC), C.c_int32, C.c_int32, C.c_int32]
        L.hbo_mesh_set.restype = None
        L.hbo_mesh_set.argtypes = [C.POINTER(CubeMeshC), C.c_int, C.c_int, C.c_int, C.c_uint16]
        L.hbo_generate_literal.restype = None
        L.hbo_generate_literal.argtypes = [C.POINTER(World), C.POINTER(CubeMeshC)]
        L.hbo_generate_ids.restype = None
        L.hbo_generate_ids.argtypes = [C.POINTER(World), C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.hbo_cubes_from_ids.restype = None
        L.hbo_cubes_from_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.hbo_cull_shared_faces.restype = None
        L.hbo_cull_shared_faces.argtypes = [C.POINTER(CubeMeshC), C.POINTER(CubeMeshC)]
        L.hbo_siphash13_chunk.restype = C.c_uint64
        L.hbo_siphash13_chunk.argtypes = [C.c_int32] * 3
        L.hbo_wyrand_next.restype = C.c_uint64
        L.hbo_wyrand_next.argtypes = [C.POINTER(C.c_uint64)]
        L.hbo_wyrand_f32.restype = C.c_float
        L.hbo_wyrand_f32.argtypes = [C.POINTER(C.c_uint64)]
        L.hbo_face_matrix.restype = None
        L.hbo_face_matrix.argtypes = [C.c_int, C.c_void_p]
        L.hbo_ao_value.restype = C.c_float
        L.hbo_ao_value.argtypes = [C.c_int]
        L.hbo_generate_mesh.restype = C.c_size_t
        L.hbo_generate_mesh.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(Palette), C.c_void_p, C.c_size_t]
        L.hbo_mesh_ids.restype = C.c_size_t
        L.hbo_mesh_ids.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(Palette), C.c_void_p, C.c_size_t]
        L.hbo_region_pipeline.restype = C.c_uint64
        L.hbo_region_pipeline.argtypes = [C.POINTER(World), C.POINTER(Palette)] + [C.c_int32] * 6 + [C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def default_palette() -> Palette:
    p = Palette()
    lib().hbo_default_palette(C.byref(p))
    return p


def make_palette(colors) -> Palette:
    p = Palette()
    p.n_materials = len(colors)
    for i, cs in enumerate(colors):
        p.mat[i].n_colors = len(cs)
        for j, c in enumerate(cs):
            for k in range(4):
                p.mat[i].rgba[j][k] = c[k]
    return p


def noise_tile2(w: World, cx: int, cz: int) -> np.ndarray:
    out = np.empty(CHUNK_AREA, dtype=np.float32)
    lib().hbo_noise_tile2(C.byref(w), cx, cz, out.ctypes.data)
    return out


def noise_tile3(w: World, start, extent, freq) -> np.ndarray:
    nx, ny, nz = extent
    out = np.empty(nx * ny * nz, dtype=np.float32)
    lib().hbo_noise_tile3(C.byref(w), *[float(s) for s in start], nx, ny, nz, *[float(f) for f in freq], out.ctypes.data)
    return out.reshape(nz, ny, nx)


def heights_from_noise(noise: np.ndarray) -> np.ndarray:
    L = lib()
    return np.array([L.hbo_height(float(v)) for v in noise.ravel()], dtype=np.int32).reshape(noise.shape)


def generate_ids(w: World, pts) -> np.ndarray:
    pts = np.asarray(pts, dtype=np.int64).reshape(-1, 3)
    out = np.empty((pts.shape[0], CHUNK_VOLUME), dtype=np.uint16)
    for i, (x, y, z) in enumerate(pts):
        lib().hbo_generate_ids(C.byref(w), int(x), int(y), int(z), out[i].ctypes.data)
    return out


def new_mesh(pos) -> CubeMeshC:
    m = CubeMeshC()
    lib().hbo_mesh_init(C.byref(m), int(pos[0]), int(pos[1]), int(pos[2]))
    return m


def mesh_cubes(m: CubeMeshC) -> np.ndarray:
    return np.frombuffer(m.data, dtype=CUBE_DTYPE).copy()


def generate_literal(w: World, pos) -> CubeMeshC:
    m = new_mesh(pos)
    lib().hbo_generate_literal(C.byref(w), C.byref(m))
    return m


def cubes_from_ids(ids: np.ndarray) -> np.ndarray:
    ids = np.ascontiguousarray(ids, dtype=np.uint16)
    out = np.empty(CHUNK_VOLUME, dtype=CUBE_DTYPE)
    lib().hbo_cubes_from_ids(ids.ctypes.data, out.ctypes.data)
    return out


def _shell_ptrs(shell):
    arr = (C.c_void_p * 27)()
    keep = []
    for i, s in enumerate(shell):
        if s is None:
            arr[i] = None
        else:
            a = np.ascontiguousarray(s)
            keep.append(a)
            arr[i] = a.ctypes.data
    return arr, keep


def mesh_ids(shell, pos, pal: Palette = None, cap: int = 200000) -> np.ndarray:
    """shell: 27 entries of u16[32768] or None (order (dx+1)*9+(dy+1)*3+(dz+1))."""
    pal = pal or default_palette()
    arr, keep = _shell_ptrs([None if s is None else np.asarray(s, dtype=np.uint16) for s in shell])
    out = np.zeros(cap, dtype=INSTANCE_DTYPE)
    p = (C.c_int32 * 3)(*[int(v) for v in pos])
    n = lib().hbo_mesh_ids(arr, p, C.byref(pal), out.ctypes.data, cap)
    assert n <= cap
    return out[:n].copy()


def mesh_cubes_literal(shell, pos, pal: Palette = None, cap: int = 200000) -> np.ndarray:
    """shell: 27 entries of CUBE_DTYPE[32768] or None; faces are read from the flags."""
    pal = pal or default_palette()
    arr, keep = _shell_ptrs(shell)
    out = np.zeros(cap, dtype=INSTANCE_DTYPE)
    p = (C.c_int32 * 3)(*[int(v) for v in pos])
    n = lib().hbo_generate_mesh(arr, p, C.byref(pal), out.ctypes.data, cap)
    assert n <= cap
    return out[:n].copy()


def region_pipeline(w: World, region, literal: bool, threads: int = 1, want_instances: bool = True,
                    want_ids: bool = False, pal: Palette = None, cap: int = 0):
    """region = (cx0, cz0, ncx, ncz, cy0, ncy).  Returns dict(total, counts, instances, ids, secs)."""
    pal = pal or default_palette()
    cx0, cz0, ncx, ncz, cy0, ncy = region
    n = ncx * ncz * ncy
    counts = np.zeros(n, dtype=np.uint32)
    secs = (C.c_double * 3)()
    ids = np.empty((n, CHUNK_VOLUME), dtype=np.uint16) if want_ids else None
    inst = None
    if want_instances:
        if cap == 0:
            cap = int(lib().hbo_region_pipeline(C.byref(w), C.byref(pal), cx0, cz0, ncx, ncz, cy0, ncy, 0, threads,
                                                counts.ctypes.data, None, 0, None, secs))
        inst = np.zeros(max(cap, 1), dtype=INSTANCE_DTYPE)
    total = lib().hbo_region_pipeline(C.byref(w), C.byref(pal), cx0, cz0, ncx, ncz, cy0, ncy, int(literal), threads,
                                      counts.ctypes.data, inst.ctypes.data if inst is not None else None,
                                      cap if inst is not None else 0, ids.ctypes.data if ids is not None else None, secs)
    return {"total": int(total), "counts": counts, "instances": inst[:total] if inst is not None else None,
            "ids": ids, "secs": tuple(secs)}


# == hbo_chunk_cube (12 bytes)
CHUNK_CUBE_DTYPE = np.dtype([("pos", "<u2"), ("pad0", "<u2"), ("material", "<u2"), ("pad", "<u2"), ("flags", "<u4")])


class ChunkMapOracle:
    """The reference's server ChunkMap + wire format + client copy, literal (oracle/hb_oracle.c, hbo_map_*).

    server: load / unload / set_cube / drain (Chunk::sync_with_client)
    client: `client[pos]` = the client's ChunkData.cubes, updated by apply(); remesh(pos) = generate_mesh over
            the client's 27-shell."""

    def __init__(self, w: World, pal: Palette = None):
        self._h = lib().hbo_map_create(C.byref(w))
        self.pal = pal or default_palette()
        self.client = {}

    def close(self):
        if self._h:
            lib().hbo_map_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def load(self, pos) -> int:
        x, y, z = (int(v) for v in pos)
        h = lib().hbo_map_load(self._h, x, y, z)
        assert h >= 0
        # client Chunk::create: all-None cubes (client/src/world/chunk.rs:105-119)
        self.client.setdefault((x, y, z), np.zeros(CHUNK_VOLUME, dtype=CUBE_DTYPE))
        return h

    def unload(self, pos) -> None:
        x, y, z = (int(v) for v in pos)
        lib().hbo_map_unload(self._h, x, y, z)
        self.client.pop((x, y, z), None)

    def set_cube(self, wpos, material: int) -> None:
        lib().hbo_map_set_cube(self._h, int(wpos[0]), int(wpos[1]), int(wpos[2]), int(material))

    def loaded(self):
        """[(handle, (x, y, z))] of the loaded chunks."""
        out = []
        p = (C.c_int32 * 3)()
        for h in range(lib().hbo_map_len(self._h)):
            if lib().hbo_map_cubes(self._h, h):
                lib().hbo_map_position(self._h, h, p)
                out.append((h, (p[0], p[1], p[2])))
        return out

    def cubes(self, pos) -> np.ndarray:
        h = lib().hbo_map_find(self._h, int(pos[0]), int(pos[1]), int(pos[2]))
        assert h >= 0
        addr = lib().hbo_map_cubes(self._h, h)
        return np.frombuffer((C.c_uint8 * (8 * CHUNK_VOLUME)).from_address(addr), dtype=CUBE_DTYPE).copy()

    def drain(self) -> dict:
        """sync_with_client for every chunk: {pos: CHUNK_CUBE_DTYPE entries (order and duplicates kept)}."""
        out = {}
        for h, pos in self.loaded():
            n = lib().hbo_map_pending(self._h, h)
            if n == 0:
                continue
            buf = np.zeros(n, dtype=CHUNK_CUBE_DTYPE)
            got = lib().hbo_map_drain(self._h, h, buf.ctypes.data, n)
            assert got == n, "oracle update log overflowed"
            out[pos] = buf
        return out

    def apply(self, updates: dict) -> None:
        """client apply_updates_from_server for every chunk that received a CubeUpdate."""
        for pos, entries in updates.items():
            if pos in self.client:
                e = np.ascontiguousarray(entries)
                lib().hbo_apply_updates(self.client[pos].ctypes.data, e.ctypes.data, e.shape[0])

    def remesh(self, pos) -> np.ndarray:
        x, y, z = pos
        shell = [self.client.get((x + dx, y + dy, z + dz)) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]
        return mesh_cubes_literal(shell, pos, self.pal)


def rle_encode(ids: np.ndarray) -> np.ndarray:
    """encode_cubes (server/src/chunk/codec.rs:73-101) of one chunk's raw ids."""
    ids = np.ascontiguousarray(ids, dtype=np.uint16).reshape(-1)
    n = lib().hbo_rle_encode(ids.ctypes.data, ids.shape[0], None, 0)
    out = np.zeros(n, dtype=np.uint8)
    assert lib().hbo_rle_encode(ids.ctypes.data, ids.shape[0], out.ctypes.data, n) == n
    return out


def encode_palette(pal: Palette, descs) -> bytes:
    """descs = [(group, key, cullable_faces, has_collider, toughness), ...], one per palette material."""
    n = len(descs)
    groups = (C.c_char_p * n)(*[d[0].encode() for d in descs])
    keys = (C.c_char_p * n)(*[d[1].encode() for d in descs])
    cull = (C.c_uint8 * n)(*[int(d[2]) for d in descs])
    coll = (C.c_uint8 * n)(*[int(bool(d[3])) for d in descs])
    tough = (C.c_float * n)(*[float(d[4]) for d in descs])
    f = lib().hbo_encode_palette
    f.restype = C.c_size_t
    f.argtypes = [C.POINTER(Palette), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    need = f(C.byref(pal), groups, keys, cull, coll, tough, None, 0)
    out = np.zeros(max(need, 1), dtype=np.uint8)
    assert f(C.byref(pal), groups, keys, cull, coll, tough, out.ctypes.data, need) == need
    return out[:need].tobytes()


def rle_decode(data: np.ndarray):
    """CubeGrid::decode record loop -> (status, raw ids)."""
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    out = np.zeros(CHUNK_VOLUME, dtype=np.uint16)
    rc = lib().hbo_rle_decode(data.ctypes.data if data.shape[0] else None, data.shape[0], out.ctypes.data)
    return rc, out
