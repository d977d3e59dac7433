// common.cuh -- shared types for libherbgpu's sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/herbgpu.h"

namespace hb {

// World parameters as the kernels see them (server/src/generator/mod.rs:103-108).
struct WorldDev {
    int32_t seed_lo;  // `seed as i32`: the only part of the seed the 2-D/3-D paths use
                      // (functions/gradient_32.rs:15, hash3d_32.rs:23)
    float freq;
    int32_t octaves;
    float lacunarity;
    float gain;
    int32_t kind;     // HB_NOISE_*: which of the crate's Sample32 implementations the tile walk calls
    int32_t fma;      // 1: neg_mul_add is one fused operation (AVX2+FMA / NEON hosts); 0: product and subtraction
                      // rounded separately (SSE2 / SSE4.1 / scalar hosts) -- simd/ops/f32.rs:141-162
};

// Constant tables built on the host with the reference's formulas (api.cu) and passed by value.
struct MeshTables {
    float face_mat[6][12];  // Instance3d model_0..2 per face (vertex.rs:53-62)
    float ao_lut[4];        // max(1 - k/3, 0.15) (client/src/world/chunk.rs:271-275)
};

struct MaterialsDev {
    int32_t n_materials;
    int32_t n_colors[HB_MAX_MATERIALS];
    alignas(16) float rgba[HB_MAX_MATERIALS][HB_MAX_COLORS][4];  // read as float4
};

struct RegionDev {
    int32_t cx0, cz0, ncx, ncz, cy0, ncy;
    int32_t ix_begin, ix_end;  // slab of the region this launch handles
    int32_t hx_begin, hx_end;  // x-range of columns present in the heights buffer (slab + apron)
};

constexpr int kHeightAbsent = INT32_MIN;  // heights-tile sentinel: column belongs to an absent chunk

static_assert(sizeof(hb_instance3d) == 100, "Instance3d must be 100 bytes");
static_assert(sizeof(hb_palette_cube) == 8, "PaletteCube must be 8 bytes");

// ---- launch wrappers implemented in gen.cu / mesh.cu (all enqueue-only) ----
cudaError_t launch_heights(cudaStream_t s, const WorldDev& w, const int32_t* d_cols_xz, size_t ncol,
                           int32_t* d_heights, float* d_noise);
// columns [col_begin, col_end) of the heights buffer (column = (ix - r.hx_begin) * r.ncz + iz)
cudaError_t launch_heights_region(cudaStream_t s, const WorldDev& w, const RegionDev& r, int32_t* d_heights,
                                  size_t col_begin, size_t col_end);
cudaError_t launch_fill(cudaStream_t s, const hb_chunk_pt* d_pts, const int32_t* d_col_of_chunk, size_t n,
                        const int32_t* d_heights, uint16_t* d_ids, hb_palette_cube* d_cubes);
// PaletteCube images of n chunks written to slots d_slots[i] of a chunk store + their updated_positions sets
cudaError_t launch_fill_world(cudaStream_t s, const hb_chunk_pt* d_pts, const int32_t* d_col_of_chunk, size_t n,
                              const int32_t* d_heights, hb_palette_cube* d_store, const int32_t* d_slots,
                              uint32_t* d_touched);
cudaError_t launch_fill_region(cudaStream_t s, const RegionDev& r, const int32_t* d_heights, uint16_t* d_ids);
cudaError_t launch_noise3d(cudaStream_t s, const WorldDev& w, float sx, float sy, float sz, int nx, int ny, int nz,
                           float fx, float fy, float fz, float* d_out);

// d_borders: [n][6][32] border records, d_counts: the chunk-local face counts (launch_mesh_ids(emit = false) adds the hull)
cudaError_t launch_pack(cudaStream_t s, const uint16_t* d_ids, size_t n, int n_materials, uint32_t* d_bitmaps,
                        uint32_t* d_summary, uint32_t* d_borders, uint32_t* d_counts, uint64_t* d_total);
cudaError_t launch_pack_cubes(cudaStream_t s, const hb_palette_cube* d_cubes, size_t n, int n_materials,
                              uint32_t* d_bitmaps, uint32_t* d_summary, uint32_t* d_planes, uint64_t* d_total,
                              const int32_t* d_list = nullptr);
// d_list (optional, both calls): work item j handles chunk d_list[j]; counts/offsets stay indexed by j while
// cubes, bitmaps, summary, planes, pts and nbr are indexed by chunk (device-resident world slots).
// d_blocks: u16 ids (d_planes == nullptr: faces recomputed) or hb_palette_cube (faces from d_planes [n][6][1024])
cudaError_t launch_mesh_ids(cudaStream_t s, bool emit, const MeshTables& tb, const MaterialsDev* d_mats,
                            const hb_chunk_pt* d_pts, size_t n, const void* d_blocks, const int32_t* d_nbr,
                            const uint32_t* d_bitmaps, const uint32_t* d_summary, const uint32_t* d_planes,
                            uint32_t* d_counts, const uint64_t* d_offsets, hb_instance3d* d_out, uint64_t capacity,
                            uint64_t* d_total, const int32_t* d_list = nullptr, const uint32_t* d_borders = nullptr);
cudaError_t launch_mesh_region(cudaStream_t s, bool emit, const MeshTables& tb, const MaterialsDev* d_mats,
                               const RegionDev& r, const int32_t* d_heights, uint32_t* d_counts,
                               const uint64_t* d_offsets, void* d_out, uint64_t capacity,
                               uint64_t* d_total, int ctas_per_sm = 0, bool packed = false);
// Single-launch fused pipeline (noise + count + offsets + emit); see k_region_fused.  d_status: one u64 per column.
// d_heights (nullable) receives the centre tiles like launch_heights_region (apron columns of the slab are not
// written).  `fast` = noise_fast_index_ok(w).
bool region_fused_supported(const WorldDev& w, const RegionDev& r);
size_t region_fused_status_bytes(const RegionDev& r);
cudaError_t launch_region_fused(cudaStream_t s, const MeshTables& tb, const MaterialsDev* d_mats, const WorldDev& w,
                                const RegionDev& r, int32_t* d_heights, uint32_t* d_counts, uint64_t* d_offsets,
                                hb_instance3d* d_out, uint64_t capacity, uint64_t* d_total, uint64_t* d_status, bool fast,
                                int ctas_per_sm = 0);
bool noise_fast_index_ok(const WorldDev& w);
// counts -> offsets (each chunk's start rounded up to 4 instances); d_total[0]=span, d_total[1]=quads.
// accumulate: offsets start at the current d_total[0] and the totals are advanced (slab chaining).
size_t scan_workspace_bytes(size_t n);
int scan_launch_count(size_t n);  // kernels launch_scan enqueues for n chunks (1 up to 16 384 chunks, else 3)
cudaError_t launch_scan(cudaStream_t s, const uint32_t* d_counts, size_t n, uint64_t* d_offsets, void* d_ws,
                        uint64_t* d_total, bool accumulate);
int num_sms();

}  // namespace hb
