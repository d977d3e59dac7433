// gen.cu -- terrain generation kernels (sm_100a).
//
//   k_heights : batched 2-D simplex FBM, one CTA per chunk column, permutation + gradient tables in shared
//               memory, 4 samples per thread evaluated two at a time with Blackwell's packed f32x2
//               arithmetic (FFMA2 / FADD2 / FMUL2).  FP32/ALU-issue bound.
//               Replaces get_2d_noise_helper_f32 + fbm_2d + simplex_2d (crates/simd-noise) and the
//               height computation of GenerationParams::generate (server/src/generator/mod.rs:74-79).
//   k_fill    : heights -> u16 block ids, one 128-bit streaming store per thread per 8 voxels,
//               warp-contiguous (512 B per store instruction).  HBM-write bound.
//               Replaces the fill loop of GenerationParams::generate (generator/mod.rs:77-94).
//   k_fill_cubes : heights -> 8-byte PaletteCube image incl. the in-chunk face flags CubeMesh::set
//               leaves behind (server/src/chunk/mesh.rs:125-232).
//   k_noise3d : 3-D simplex FBM tile (crates/simd-noise/src/noise/f32.rs:131-198).
#include <math.h>

#include "noise.cuh"

namespace hb {

namespace {

constexpr int kThreads = 256;

struct ColList {
    const int32_t* cols_xz;
    __device__ __forceinline__ int first() const { return 0; }
    __device__ __forceinline__ void get(int col, int& cx, int& cz) const {
        cx = cols_xz[2 * col];
        cz = cols_xz[2 * col + 1];
    }
};
struct ColRegion {
    RegionDev r;
    int col0;  // first heights-buffer column of this launch
    __device__ __forceinline__ int first() const { return col0; }
    __device__ __forceinline__ void get(int col, int& cx, int& cz) const {
        const int ix = r.hx_begin + col / r.ncz;
        const int iz = col % r.ncz;
        cx = r.cx0 + ix;
        cz = r.cz0 + iz;
    }
};

// One work item = one chunk column = 32x32 samples; persistent CTAs stride through the columns so the tables are
// staged once per CTA and there is no barrier in the column loop.  Thread mapping: sample i has x = i >> 5,
// z = i & 31, i.e. heights are produced directly in voxel-column order x*32 + z (coalesced stores, no transpose);
// the optional raw noise tile keeps the crate's order x + z*32 (noise/f32.rs:95-113) with a strided store.
template <class Cols, int KIND, bool FAST, bool FMA>
__global__ void __launch_bounds__(kThreads, 8) k_heights(WorldDev w, Cols cols, int ncol, int32_t* __restrict__ heights,
                                                       float* __restrict__ noise) {
    __shared__ NoiseTables s_tab;
    load_noise_tables(s_tab, w.seed_lo);
    __syncthreads();
    for (int c = blockIdx.x; c < ncol; c += gridDim.x) {
        const int col = cols.first() + c;
        int cx, cz;
        cols.get(col, cx, cz);
        // chunk.position.xz().cast() * CHUNK_LENGTH as f32 (generator/mod.rs:74); lane coordinate =
        // start + i, exact in f32 for |c| <= HB_MAX_CHUNK_COORD
        const float start_x = __fmul_rn(__int2float_rn(cx), 32.0f);
        const float start_z = __fmul_rn(__int2float_rn(cz), 32.0f);
#pragma unroll
        for (int k = 0; k < 4; k += 2) {  // two samples per packed evaluation (FFMA2 / FADD2 / FMUL2)
            const int i0 = threadIdx.x + k * kThreads, i1 = i0 + kThreads;
            const float2 x = pk(__fadd_rn(start_x, __int2float_rn(i0 >> 5)), __fadd_rn(start_x, __int2float_rn(i1 >> 5)));
            const float2 z = pk(__fadd_rn(start_z, __int2float_rn(i0 & 31)), __fadd_rn(start_z, __int2float_rn(i1 & 31)));
            const float2 r = noise2_pair<KIND, FAST, FMA>(s_tab, smul2(x, bc(w.freq)), smul2(z, bc(w.freq)), w);  // scalar muls: they feed adds
            if (noise) {
                noise[(size_t)col * HB_CHUNK_AREA + (i0 >> 5) + (i0 & 31) * 32] = r.x;
                noise[(size_t)col * HB_CHUNK_AREA + (i1 >> 5) + (i1 & 31) * 32] = r.y;
            }
            if (heights) {
                heights[(size_t)col * HB_CHUNK_AREA + i0] = height_from_noise(r.x);
                heights[(size_t)col * HB_CHUNK_AREA + i1] = height_from_noise(r.y);
            }
        }
    }
}

struct ChunkList {
    const hb_chunk_pt* pts;
    const int32_t* col_of_chunk;
    __device__ __forceinline__ void get(size_t chunk, int& col, int& cy) const {
        col = col_of_chunk[chunk];
        cy = pts[chunk].y;
    }
};
struct ChunkRegion {
    RegionDev r;
    // chunk index is relative to the slab: ((ix-ix_begin)*ncz + iz)*ncy + iy
    __device__ __forceinline__ void get(size_t chunk, int& col, int& cy) const {
        const int iy = (int)(chunk % r.ncy);
        const size_t c2 = chunk / r.ncy;
        const int iz = (int)(c2 % r.ncz);
        const int ix = r.ix_begin + (int)(c2 / r.ncz);
        col = (ix - r.hx_begin) * r.ncz + iz;
        cy = r.cy0 + iy;
    }
};

// generator/mod.rs:85-91 with d = h - y: stone if y < h-6, dirt if y < h-1, grass if y < h
__device__ __forceinline__ uint32_t block_id(int d) { return d > 6 ? 1u : (d > 1 ? 2u : (d > 0 ? 3u : 0u)); }

// 8 consecutive voxels of a column starting at local y0 depend only on d = h - (wy0 + y0), and only through
// clamp(d, 0, 14): d <= 0 is all air, d >= 14 is all stone.  The 15 possible 128-bit id vectors live in shared
// memory, so a store costs one table load instead of ~40 select instructions.
template <class Chunks>
__global__ void __launch_bounds__(kThreads) k_fill(Chunks chunks, const int32_t* __restrict__ heights,
                                                    uint16_t* __restrict__ ids) {
    __shared__ int s_h[HB_CHUNK_AREA];
    __shared__ uint4 s_run[16];
    const size_t chunk = blockIdx.x;
    int col, cy;
    chunks.get(chunk, col, cy);
    const int4* src = reinterpret_cast<const int4*>(heights + (size_t)col * HB_CHUNK_AREA);
    reinterpret_cast<int4*>(s_h)[threadIdx.x] = __ldg(src + threadIdx.x);
    if (threadIdx.x < 15) {
        const int d = threadIdx.x;
        uint4 o;
        o.x = block_id(d) | (block_id(d - 1) << 16);
        o.y = block_id(d - 2) | (block_id(d - 3) << 16);
        o.z = block_id(d - 4) | (block_id(d - 5) << 16);
        o.w = block_id(d - 6) | (block_id(d - 7) << 16);
        s_run[d] = o;
    }
    __syncthreads();
    const long long wy0 = (long long)cy * 32;
    uint4* dst = reinterpret_cast<uint4*>(ids + chunk * HB_CHUNK_VOLUME);
#pragma unroll 8
    for (int it = 0; it < 16; ++it) {
        const int u = it * kThreads + threadIdx.x;  // 8-voxel unit: column u>>2, y0 = (u&3)*8
        const long long dl = (long long)s_h[u >> 2] - (wy0 + (u & 3) * 8);
        const int d = (int)(dl < 0 ? 0 : (dl > 14 ? 14 : dl));
        __stcs(dst + u, s_run[d]);
    }
}

// PaletteCube image after generate(): for a solid voxel, face f is kept iff the in-chunk neighbour
// in direction f is air or f leaves the chunk; an air voxel has flags = 1 iff some in-chunk neighbour
// is solid (CubeMesh::set touches it through remove_neighboring_faces, mesh.rs:209-232), else 0.
//
// Phase 1: one thread per column turns heights into eight 32-bit masks over y (six "face kept" masks, the
// solid mask, "some in-chunk neighbour is solid").  Phase 2: one thread per voxel pair reads its column's masks
// (two 128-bit shared loads, broadcast across the 16 lanes of a column) and writes 16 bytes; a warp instruction
// stores 512 contiguous bytes.
//
// World mode (slots != nullptr): chunk i is written to slot slots[i] of the device-resident store, and the set of
// positions CubeMesh::set pushed to updated_positions -- every solid voxel and every in-chunk neighbour of one --
// goes to touched[slot][column] (bit y).
template <class Chunks>
__global__ void __launch_bounds__(kThreads) k_fill_cubes(Chunks chunks, const int32_t* __restrict__ heights,
                                                          hb_palette_cube* __restrict__ cubes,
                                                          const int32_t* __restrict__ slots, uint32_t* __restrict__ touched) {
    __shared__ int s_d[HB_CHUNK_AREA];                   // clamp(h - wy0, 0, 40): solid voxels in the column, + margin
    __shared__ __align__(16) uint32_t s_m[HB_CHUNK_AREA][8];  // E W U D N S kept-face masks, solid, neighbour-solid
    const size_t chunk = blockIdx.x;
    const size_t slot = slots ? (size_t)slots[chunk] : chunk;
    int col, cy;
    chunks.get(chunk, col, cy);
    const long long wy0 = (long long)cy * 32;
    {
        const int4 h4 = __ldg(reinterpret_cast<const int4*>(heights + (size_t)col * HB_CHUNK_AREA) + threadIdx.x);
        auto rel = [&](int h) -> int {
            const long long d = (long long)h - wy0;
            return (int)(d < 0 ? 0 : (d > 40 ? 40 : d));
        };
        reinterpret_cast<int4*>(s_d)[threadIdx.x] = make_int4(rel(h4.x), rel(h4.y), rel(h4.z), rel(h4.w));
    }
    __syncthreads();
    auto solid_mask = [&](int c) -> uint32_t {
        const int d = s_d[c];
        return d >= 32 ? 0xFFFFFFFFu : ((1u << d) - 1u);
    };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = k * kThreads + threadIdx.x, x = c >> 5, z = c & 31;
        const uint32_t so = solid_mask(c);
        const uint32_t nE = x < 31 ? solid_mask(c + 32) : 0u, nW = x > 0 ? solid_mask(c - 32) : 0u;
        const uint32_t nN = z < 31 ? solid_mask(c + 1) : 0u, nS = z > 0 ? solid_mask(c - 1) : 0u;
        const uint32_t nU = so >> 1, nD = so << 1;  // in-chunk voxel above / below
        uint4 lo, hi;
        lo.x = so & ~nE; lo.y = so & ~nW; lo.z = so & ~nU; lo.w = so & ~nD;
        hi.x = so & ~nN; hi.y = so & ~nS; hi.z = so; hi.w = nE | nW | nU | nD | nN | nS;
        *reinterpret_cast<uint4*>(&s_m[c][0]) = lo;
        *reinterpret_cast<uint4*>(&s_m[c][4]) = hi;
        if (touched) touched[slot * HB_CHUNK_AREA + c] = hi.z | hi.w;
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(cubes + slot * HB_CHUNK_VOLUME);
#pragma unroll 4
    for (int it = 0; it < 64; ++it) {
        const int p = it * kThreads + threadIdx.x;  // voxel pair: L = 2p, 2p+1 (same column, y even / odd)
        const int c = p >> 4, y = (p & 15) * 2;
        const uint4 lo = *reinterpret_cast<const uint4*>(&s_m[c][0]);
        const uint4 hi = *reinterpret_cast<const uint4*>(&s_m[c][4]);
        const int d = s_d[c] - y;  // h - wy of the even voxel
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int yy = y + e;
            const uint32_t faces = ((lo.x >> yy) & 1u) | (((lo.y >> yy) & 1u) << 1) | (((lo.z >> yy) & 1u) << 2) |
                                   (((lo.w >> yy) & 1u) << 3) | (((hi.x >> yy) & 1u) << 4) | (((hi.y >> yy) & 1u) << 5);
            const bool solid = (hi.z >> yy) & 1u;
            w[2 * e] = block_id(d - e);                                       // material u16 + zero pad
            w[2 * e + 1] = solid ? ((faces << 1) | 1u) : ((hi.w >> yy) & 1u);  // CubeFlags (cube.rs:69-72)
        }
        __stcs(dst + p, make_uint4(w[0], w[1], w[2], w[3]));
    }
}

__global__ void __launch_bounds__(kThreads) k_noise3d(WorldDev w, float sx, float sy, float sz, int nx, int ny,
                                                       int nz, float fx, float fy, float fz,
                                                       float* __restrict__ out) {
    const size_t total = (size_t)nx * ny * nz;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
        const int ix = (int)(i % nx);
        const int iy = (int)((i / nx) % ny);
        const int iz = (int)(i / ((size_t)nx * ny));
        // coordinates advance by repeated +1.0 / +WIDTH in the reference; integers, exact in f32
        const float x = __fadd_rn(sx, __int2float_rn(ix));
        const float y = __fadd_rn(sy, __int2float_rn(iy));
        const float z = __fadd_rn(sz, __int2float_rn(iz));
        out[i] = noise3(__fmul_rn(x, fx), __fmul_rn(y, fy), __fmul_rn(z, fz), w);
    }
}

// FAST needs every lattice coordinate floor(x + F2*(x + y)) below 2^20 in magnitude (noise.cuh).  World block
// coordinates are bounded by the chunk-coordinate limit, each octave multiplies by the lacunarity, and the skew adds
// at most 2*F2 < 0.74 of the larger coordinate; 1.75 leaves a margin for rounding.
bool fast_index_ok_impl(const WorldDev& w) {
    const double c = ((double)HB_MAX_CHUNK_COORD + 1.0) * 32.0 * fabs((double)w.freq);
    const double l = fabs((double)w.lacunarity) > 1.0 ? fabs((double)w.lacunarity) : 1.0;
    const int oct = w.kind == HB_NOISE_GRADIENT ? 1 : w.octaves;
    const double top = c * pow(l, oct - 1) * 1.75;
    return top == top && top < 1048576.0;
}

template <class Cols>
void launch_heights_kind(cudaStream_t s, const WorldDev& w, Cols cols, unsigned grid, int32_t* d_heights, float* d_noise) {
    const int ncol = (int)grid;
    const unsigned cap = (unsigned)num_sms() * 16u;  // two waves of the 8 resident CTAs per SM: measured best (8: +2 %, one CTA per column: +2 %)
    if (grid > cap) grid = cap;
    // the unfused arithmetic column (w.fma == 0) exists with the exact-conversion index path only
#define HB_LAUNCH_KIND(K)                                                                                           \
    do {                                                                                                            \
        if (!w.fma) k_heights<Cols, K, false, false><<<grid, kThreads, 0, s>>>(w, cols, ncol, d_heights, d_noise);     \
        else if (fast) k_heights<Cols, K, true, true><<<grid, kThreads, 0, s>>>(w, cols, ncol, d_heights, d_noise);    \
        else k_heights<Cols, K, false, true><<<grid, kThreads, 0, s>>>(w, cols, ncol, d_heights, d_noise);            \
    } while (0)
    const bool fast = fast_index_ok_impl(w);
    switch (w.kind) {
        case HB_NOISE_RIDGE: HB_LAUNCH_KIND(HB_NOISE_RIDGE); break;
        case HB_NOISE_TURBULENCE: HB_LAUNCH_KIND(HB_NOISE_TURBULENCE); break;
        case HB_NOISE_GRADIENT: HB_LAUNCH_KIND(HB_NOISE_GRADIENT); break;
        default: HB_LAUNCH_KIND(HB_NOISE_FBM); break;
    }
#undef HB_LAUNCH_KIND
}

}  // namespace

bool noise_fast_index_ok(const WorldDev& w) { return fast_index_ok_impl(w); }

cudaError_t launch_heights(cudaStream_t s, const WorldDev& w, const int32_t* d_cols_xz, size_t ncol,
                           int32_t* d_heights, float* d_noise) {
    if (ncol == 0) return cudaSuccess;
    launch_heights_kind(s, w, ColList{d_cols_xz}, (unsigned)ncol, d_heights, d_noise);
    return cudaGetLastError();
}

cudaError_t launch_heights_region(cudaStream_t s, const WorldDev& w, const RegionDev& r, int32_t* d_heights,
                                  size_t col_begin, size_t col_end) {
    if (col_end <= col_begin) return cudaSuccess;
    launch_heights_kind(s, w, ColRegion{r, (int)col_begin}, (unsigned)(col_end - col_begin), d_heights, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_fill(cudaStream_t s, const hb_chunk_pt* d_pts, const int32_t* d_col_of_chunk, size_t n,
                        const int32_t* d_heights, uint16_t* d_ids, hb_palette_cube* d_cubes) {
    if (n == 0) return cudaSuccess;
    ChunkList cl{d_pts, d_col_of_chunk};
    if (d_ids) k_fill<ChunkList><<<(unsigned)n, kThreads, 0, s>>>(cl, d_heights, d_ids);
    if (d_cubes) k_fill_cubes<ChunkList><<<(unsigned)n, kThreads, 0, s>>>(cl, d_heights, d_cubes, nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_fill_world(cudaStream_t s, const hb_chunk_pt* d_pts, const int32_t* d_col_of_chunk, size_t n,
                              const int32_t* d_heights, hb_palette_cube* d_store, const int32_t* d_slots,
                              uint32_t* d_touched) {
    if (n == 0) return cudaSuccess;
    k_fill_cubes<ChunkList><<<(unsigned)n, kThreads, 0, s>>>(ChunkList{d_pts, d_col_of_chunk}, d_heights, d_store, d_slots,
                                                             d_touched);
    return cudaGetLastError();
}

cudaError_t launch_fill_region(cudaStream_t s, const RegionDev& r, const int32_t* d_heights, uint16_t* d_ids) {
    const size_t n = (size_t)(r.ix_end - r.ix_begin) * r.ncz * r.ncy;
    if (n == 0) return cudaSuccess;
    k_fill<ChunkRegion><<<(unsigned)n, kThreads, 0, s>>>(ChunkRegion{r}, d_heights, d_ids);
    return cudaGetLastError();
}

cudaError_t launch_noise3d(cudaStream_t s, const WorldDev& w, float sx, float sy, float sz, int nx, int ny, int nz,
                           float fx, float fy, float fz, float* d_out) {
    const size_t total = (size_t)nx * ny * nz;
    if (total == 0) return cudaSuccess;
    size_t blocks = (total + kThreads - 1) / kThreads;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_noise3d<<<(unsigned)blocks, kThreads, 0, s>>>(w, sx, sy, sz, nx, ny, nz, fx, fy, fz, d_out);
    return cudaGetLastError();
}

}  // namespace hb
