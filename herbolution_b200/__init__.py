"""herbolution_b200 -- B200-native chunk pipeline (terrain generation + chunk meshing) of herbolution.

The product is libherbgpu.so (hand-written sm_100a CUDA behind the C ABI in include/herbgpu.h).
This package is the thin host-side mirror of the reference's generator / mesher interface over that
ABI.  Importing it never touches the CPU oracle under oracle/; using it without the built CUDA
library, or without a GPU, raises.
"""
from ._lib import (CHUNK_AREA, CHUNK_CUBE_DTYPE, CHUNK_LENGTH, CHUNK_PT_DTYPE, CHUNK_VOLUME, CUBE_DTYPE, INSTANCE_DTYPE, HerbGpuError,
                   Region, STAGE_ALL_MESH, STAGE_COUNT, STAGE_EMIT, STAGE_EMIT_PACKED, STAGE_FILL, STAGE_FUSED, STAGE_HEIGHTS, STAGE_SCAN,
                   TRANSPORT_INSTANCES, TRANSPORT_PACKED)
from .context import Context, MultiContext, RegionPlan
from .generator import ChunkGenerator, ChunkPt, CubeMesh, GenerationParams
from .mesher import ChunkData, ChunkMap, create_chunk_shell, generate_mesh
from .world import ServerChunkMap, World

__all__ = [
    "CHUNK_AREA", "CHUNK_CUBE_DTYPE", "CHUNK_LENGTH", "CHUNK_PT_DTYPE", "CHUNK_VOLUME", "CUBE_DTYPE", "INSTANCE_DTYPE", "HerbGpuError",
    "Region", "STAGE_ALL_MESH", "STAGE_COUNT", "STAGE_EMIT", "STAGE_EMIT_PACKED", "STAGE_FILL", "STAGE_FUSED", "STAGE_HEIGHTS", "STAGE_SCAN",
    "TRANSPORT_INSTANCES", "TRANSPORT_PACKED",
    "Context", "MultiContext", "RegionPlan", "ChunkGenerator", "ChunkPt", "CubeMesh", "GenerationParams",
    "ChunkData", "ChunkMap", "create_chunk_shell", "generate_mesh", "ServerChunkMap", "World",
]
