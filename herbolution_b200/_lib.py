"""ctypes binding of libherbgpu.so -- the same C ABI (include/herbgpu.h) a Rust host would bind.

The library is built in-tree by herbolution_b200.build (nvcc, sm_100a).  There is no CPU fallback:
if the shared object is missing this module raises, and hb_create fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libherbgpu.so")

CHUNK_LENGTH = 32
CHUNK_AREA = 1024
CHUNK_VOLUME = 32768
MAX_MATERIALS = 16
MAX_COLORS = 8

HB_OK = 0
HB_ERR_INVALID = -1
HB_ERR_CUDA = -2
HB_ERR_CAPACITY = -3
HB_ERR_RANGE = -4
HB_ERR_NOMEM = -5

STAGE_HEIGHTS, STAGE_COUNT, STAGE_SCAN, STAGE_EMIT, STAGE_FILL = 1, 2, 4, 8, 16
STAGE_EMIT_PACKED = 32
STAGE_FUSED = 64
STAGE_ALL_MESH = 15
TRANSPORT_PACKED, TRANSPORT_INSTANCES = 0, 1
NOISE_FBM, NOISE_RIDGE, NOISE_TURBULENCE, NOISE_GRADIENT = 0, 1, 2, 3

# == hb_chunk_pt / lib::point::ChunkPt
CHUNK_PT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4")])
# == hb_instance3d / client Instance3d (100 bytes, client/src/video/world/vertex.rs:41-51)
INSTANCE_DTYPE = np.dtype([("model", "<f4", (16,)), ("rgba", "<f4", (4,)), ("light", "<u4"), ("ao", "<f4", (4,))])
# == hb_palette_cube / PaletteCube (8 bytes, server/src/chunk/cube.rs:5-10)
CUBE_DTYPE = np.dtype([("material", "<u2"), ("pad", "<u2"), ("flags", "<u4")])
# == hb_chunk_cube / ChunkCube (12 bytes: vec3u5 position + PaletteCube, server/src/chunk/handle.rs:28-32)
CHUNK_CUBE_DTYPE = np.dtype([("pos", "<u2"), ("pad0", "<u2"), ("material", "<u2"), ("pad", "<u2"), ("flags", "<u4")])
assert INSTANCE_DTYPE.itemsize == 100 and CUBE_DTYPE.itemsize == 8 and CHUNK_PT_DTYPE.itemsize == 12
assert CHUNK_CUBE_DTYPE.itemsize == 12


class MaterialDesc(C.Structure):  # == hb_material_desc
    _fields_ = [("group", C.c_char_p), ("key", C.c_char_p), ("cullable_faces", C.c_uint8), ("has_collider", C.c_uint8),
                ("toughness", C.c_float)]


class ChunkPtC(C.Structure):
    """hb_chunk_pt passed by value."""

    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32)]



class Region(C.Structure):
    """hb_region: cx in [cx0, cx0+ncx), cz in [cz0, cz0+ncz), cy in [cy0, cy0+ncy)."""

    _fields_ = [("cx0", C.c_int32), ("cz0", C.c_int32), ("ncx", C.c_int32), ("ncz", C.c_int32),
                ("cy0", C.c_int32), ("ncy", C.c_int32)]

    @property
    def n_chunks(self) -> int:
        return self.ncx * self.ncz * self.ncy

    def chunk_points(self) -> np.ndarray:
        """Chunk coordinates in region index order ((ix*ncz)+iz)*ncy+iy."""
        ix, iz, iy = np.meshgrid(np.arange(self.ncx), np.arange(self.ncz), np.arange(self.ncy), indexing="ij")
        pts = np.empty(self.n_chunks, dtype=CHUNK_PT_DTYPE)
        pts["x"] = (self.cx0 + ix).ravel()
        pts["y"] = (self.cy0 + iy).ravel()
        pts["z"] = (self.cz0 + iz).ravel()
        return pts


class HerbGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"herbgpu error {code}: {message}")
        self.code = code


REGION_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                          C.c_void_p)

# every symbol include/herbgpu.h declares: (restype, argtypes)
PROTOTYPES = {
    "hb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "hb_destroy": (None, [C.c_void_p]),
    "hb_set_noise_kind": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_set_fma_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_last_error": (C.c_char_p, []),
    "hb_version": (C.c_char_p, []),
    "hb_set_world": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_float, C.c_float]),
    "hb_set_materials": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "hb_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "hb_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_noise_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "hb_noise_fbm3d": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p]),
    "hb_generate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_build_neighbor_index": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "hb_mesh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                          C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "hb_mesh_cubes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "hb_generate_mesh_region": (C.c_int, [C.c_void_p, C.POINTER(Region), C.c_int, REGION_SINK, C.c_void_p,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hb_generate_mesh_region_slab": (C.c_int, [C.c_void_p, C.POINTER(Region), C.c_int32, C.c_int32, C.c_int, REGION_SINK,
                                               C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hb_set_region_transport": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_set_host_threads": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_generate_mesh_region_packed": (C.c_int, [C.c_void_p, C.POINTER(Region), C.c_int32, C.c_int32, C.c_int, REGION_SINK,
                                                 C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hb_expand_quads": (C.c_int, [C.c_void_p, ChunkPtC, C.c_void_p, C.c_uint32, C.c_void_p]),
    "hb_multi_create": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    "hb_multi_destroy": (None, [C.c_void_p]),
    "hb_multi_device_count": (C.c_int32, [C.c_void_p]),
    "hb_multi_ctx": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "hb_multi_set_world": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_float, C.c_float]),
    "hb_multi_set_noise_kind": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_multi_set_fma_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_multi_set_materials": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "hb_multi_set_region_transport": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_multi_set_host_threads": (C.c_int, [C.c_void_p, C.c_int32]),
    "hb_multi_slab": (C.c_int, [C.c_void_p, C.POINTER(Region), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hb_multi_generate_mesh_region": (C.c_int, [C.c_void_p, C.POINTER(Region), C.c_int, REGION_SINK, C.c_void_p,
                                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hb_encode_palette": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "hb_host_expand_quads": (C.c_int, [ChunkPtC, C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_host_fill_ids": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "hb_probe_d2h": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_double)]),
    "hb_probe_host_write": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_double)]),
    "hb_region_plan_packed": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "hb_region_plan_create": (C.c_int, [C.c_void_p, C.POINTER(Region), C.c_int32, C.c_int32, C.c_uint64, C.c_int,
                                        C.POINTER(C.c_void_p)]),
    "hb_region_plan_destroy": (None, [C.c_void_p]),
    "hb_region_plan_enqueue": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hb_region_plan_result": (C.c_int, [C.c_void_p, C.c_void_p] + [C.POINTER(C.c_uint64)] * 3 + [C.POINTER(C.c_void_p)] * 5),
    "hb_region_plan_launches": (C.c_int, [C.c_void_p, C.c_int]),
    "hb_dev_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hb_dev_generate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_dev_mesh_workspace_bytes": (C.c_size_t, [C.c_size_t]),
    "hb_dev_mesh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    # device-resident world (server ChunkMap + CubeUpdate wire format + client remesh)
    "hb_world_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "hb_world_destroy": (None, [C.c_void_p]),
    "hb_world_len": (C.c_size_t, [C.c_void_p]),
    "hb_world_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "hb_world_unload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "hb_world_set_cubes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hb_world_last_edit_groups": (C.c_uint32, [C.c_void_p]),
    "hb_host_group_edits": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint32)]),
    "hb_world_sync": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]),
    "hb_world_fetch_updates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint64, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]),
    "hb_world_remesh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]),
    "hb_world_remesh_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]),
    "hb_world_read_chunk": (C.c_int, [C.c_void_p, ChunkPtC, C.c_void_p]),
    # on-disk cube codec (server/src/chunk/codec.rs)
    "hb_rle_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_uint64)]),
    "hb_rle_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hb_dev_rle_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libherbgpu.so and bind every exported symbol.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m herbolution_b200.build` "
            "(libherbgpu is CUDA-only; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != HB_OK:
        raise HerbGpuError(rc, load().hb_last_error().decode("utf-8", "replace"))


def ptr(a) -> C.c_void_p:
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return C.c_void_p(None)
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return C.c_void_p(a.ctypes.data)
