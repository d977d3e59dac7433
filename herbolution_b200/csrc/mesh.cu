// mesh.cu -- chunk meshing kernels (sm_100a): visible-face culling incl. cross-chunk borders,
// ambient occlusion, per-face colour, atomics-free compaction into packed Instance3d buffers.
//
// Replaces, for a whole batch per launch:
//   - the in-chunk face flags CubeMesh::set maintains        server/src/chunk/mesh.rs:125-232
//   - CubeMesh::cull_shared_faces across chunk borders         server/src/chunk/mesh.rs:76-123
//   - client generate_mesh / facial_ao / Instance3d::new        client/src/world/chunk.rs:176-276,
//                                                               client/src/video/world/vertex.rs:53-68
//
// Data layout.  A chunk (plus its one-voxel halo from the 26 neighbours) is staged in shared memory
// as an OCCUPANCY TILE: 34x34 columns of 64-bit words, bit (y+1) of occ[x+1][z+1] = voxel (x,y,z)
// solid (y is the fastest voxel axis, so one word is one vertical run).  Absent neighbour chunks
// contribute zero bits, which yields the reference's semantics for unloaded chunks (border faces
// emitted, AO probes "empty").  Face visibility is then six 32-bit mask expressions per column,
// counting is popc, and the per-chunk compaction is a warp-shuffle scan + one block-level combine
// in the reference's emission order (ascending L = x*1024+z*32+y, then face order) -- no atomics.
//
// Emission: visible faces -> 32-bit descriptors in shared memory (L | face<<15 | ao<<18), then one
// lane per quad expands a descriptor to the 100-byte Instance3d in a warp-private shared staging
// buffer (stride 25 words: bank-conflict free), and one lane hands the 3200-byte window to the TMA
// engine as a 1-D bulk shared->global copy (cp.async.bulk, SASS UBLKCP).  Each chunk's output starts on a 4-instance (400 B) boundary so every
// vector store is 16-byte aligned.
#include <algorithm>

#include "noise.cuh"

namespace hb {

namespace {

constexpr int kT = 256;         // threads per CTA
constexpr int kCtasPerSm = 4;   // register/smem budget the mesh kernels are compiled for (<= 64 regs, 48.5 KB)

constexpr int kTile = 34;
constexpr int kTileN = kTile * kTile;
constexpr int kDescCap = 2048;  // descriptors staged per round
constexpr int kStageCap = kT;   // quads expanded per window (one per thread)

typedef unsigned long long u64;

constexpr int kSeedLayers = 32;
// DESC = descriptors staged per round (the fused region kernel halves it to fit 5 CTAs per SM)
template <int DESC>
struct MeshSmemT {
    static constexpr int kDesc = DESC;
    u64 occ[kTileN];                            // 9248 B
    __align__(16) float stage[kStageCap * 25];  // 25600 B
    uint32_t desc[DESC];                        // 8192 B
    int hts[kTileN];                            // 4624 B (region kernels: heights halo tile)
    __align__(16) float face_mat[6][12];
    float ao_lut[4];
    __align__(16) int ao_tab[6][8];             // per face: 8 AO probes as (tile index offset)*4 + (bit offset + 1)
    uint32_t warp_tot[32];
    u64 rng_seed;
    int nbr[27];
    int red[3][8];
    u64 layer_seed[kSeedLayers];                // region kernels: the column's per-layer chunk hashes
};
using MeshSmem = MeshSmemT<kDescCap>;

// ---- SipHash-1-3 with zero keys over the 12 bytes of ChunkPt: std DefaultHasher as used at
// client/src/world/chunk.rs:182-184 ----
__device__ __forceinline__ u64 rotl64(u64 x, int b) { return (x << b) | (x >> (64 - b)); }
#define HB_SIPROUND                                                   \
    do {                                                              \
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32); \
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                      \
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                      \
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32); \
    } while (0)
__device__ u64 siphash13_chunk(int x, int y, int z) {
    u64 v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    u64 v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    const u64 m0 = (u64)(uint32_t)x | ((u64)(uint32_t)y << 32);
    v3 ^= m0; HB_SIPROUND; v0 ^= m0;
    const u64 b = ((u64)12 << 56) | (u64)(uint32_t)z;
    v3 ^= b; HB_SIPROUND; v0 ^= b;
    v2 ^= 0xff;
    HB_SIPROUND; HB_SIPROUND; HB_SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

// fastrand 2.3.0 wyrand in counter form: the reference draws 6 f32 per voxel in index order
// (client/src/world/chunk.rs:195, lib/src/spatial/face.rs:430-442), so draw n = 6*L + face uses
// state seed + (n+1)*0x2d358dccaa6c78a5.  Returns Rng::f32().
__device__ __forceinline__ float wyrand_f32_at(u64 seed, uint32_t n) {
    const u64 s = seed + (u64)(n + 1u) * 0x2d358dccaa6c78a5ULL;
    const u64 b = s ^ 0x8bb84b93962eacc9ULL;
    const u64 r = (s * b) ^ __umul64hi(s, b);
    return __fsub_rn(__uint_as_float(0x3F800000u | ((uint32_t)r >> 9)), 1.0f);
}

// generator/mod.rs:85-91 with d = h - y
__device__ __forceinline__ uint32_t block_id(int d) { return d > 6 ? 1u : (d > 1 ? 2u : (d > 0 ? 3u : 0u)); }

// vertex_ao, client/src/world/chunk.rs:251-253
__device__ __forceinline__ uint32_t vao(uint32_t s1, uint32_t s2, uint32_t c) { return (s1 & s2) ? 3u : s1 + s2 + c; }

// facial_ao (chunk.rs:255-276), table driven so that one thread per quad can evaluate it without
// divergence.  orthonormal_basis (lib/src/spatial/face.rs:84-93) gives (u, v, n) per face; the 8
// probes are position + n + a*u + b*v for (a, b) in the order tl, tc, tr, ml, mr, bl, bc, br.
__device__ const signed char kBasis[6][9] = {
    // ux uy uz  vx vy vz  nx ny nz
    {0, 0, -1, 0, 1, 0, 1, 0, 0},   // East  (-Z,  Y,  X)
    {0, 0, 1, 0, 1, 0, -1, 0, 0},   // West  ( Z,  Y, -X)
    {1, 0, 0, 0, 0, -1, 0, 1, 0},   // Up    ( X, -Z,  Y)
    {1, 0, 0, 0, 0, 1, 0, -1, 0},   // Down  ( X,  Z, -Y)
    {1, 0, 0, 0, 1, 0, 0, 0, 1},    // North ( X,  Y,  Z)
    {-1, 0, 0, 0, 1, 0, 0, 0, -1},  // South (-X,  Y, -Z)
};
__device__ const signed char kProbeAB[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};

// entry = (dx*kTile + dz)*4 + (dy + 1); decode: index offset = e >> 2 (arithmetic), bit offset = (e & 3) - 1
__device__ __forceinline__ int ao_tab_entry(int f, int p) {
    const int a = kProbeAB[p][0], b = kProbeAB[p][1];
    const int dx = kBasis[f][6] + a * kBasis[f][0] + b * kBasis[f][3];
    const int dy = kBasis[f][7] + a * kBasis[f][1] + b * kBasis[f][4];
    const int dz = kBasis[f][8] + a * kBasis[f][2] + b * kBasis[f][5];
    return (dx * kTile + dz) * 4 + (dy + 1);
}

// Returns the four vertex_ao values (2 bits each: bl, tl, br, tr order of chunk.rs:266-269).
__device__ __forceinline__ uint32_t ao_bits(const u64* __restrict__ occ, const int* __restrict__ tab, int x, int y, int z) {
    const int base = (x + 1) * kTile + (z + 1);
    const int4 t0 = *reinterpret_cast<const int4*>(tab);
    const int4 t1 = *reinterpret_cast<const int4*>(tab + 4);
#define HB_P(e) ((uint32_t)(occ[base + ((e) >> 2)] >> (y + ((e) & 3))) & 1u)
    const uint32_t tl = HB_P(t0.x), tc = HB_P(t0.y), tr = HB_P(t0.z), ml = HB_P(t0.w);
    const uint32_t mr = HB_P(t1.x), bl = HB_P(t1.y), bc = HB_P(t1.z), br = HB_P(t1.w);
#undef HB_P
    return vao(ml, tc, tl) | (vao(ml, bc, bl) << 2) | (vao(mr, tc, tr) << 4) | (vao(mr, bc, br) << 6);
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ---- TMA 1-D bulk copy shared -> global (cp.async.bulk; SASS: UBLKCP).  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc_shared, uint32_t bytes) {
    // The instances are written once and never read back by this pipeline: an evict_first L2 policy lets them
    // leave the L2 in stream order instead of displacing the heights tiles (measured: -2 % on the emit kernel).
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n\tcp.async.bulk.commit_group;"
                 :: "l"(gdst), "r"(ssrc_shared), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// The six visible-face masks (bit y) of column (x, z) of the occupancy tile.
__device__ __forceinline__ void column_masks(const u64* __restrict__ occ, int x, int z, uint32_t m[6]) {
    const u64* o = occ + (x + 1) * kTile + (z + 1);
    const u64 C = o[0];
    const uint32_t cen = (uint32_t)(C >> 1);
    if (cen) {
        m[0] = cen & ~(uint32_t)(o[kTile] >> 1);   // East : (x+1, y, z) empty
        m[1] = cen & ~(uint32_t)(o[-kTile] >> 1);  // West
        m[2] = cen & ~(uint32_t)(C >> 2);          // Up   : (x, y+1, z)
        m[3] = cen & ~(uint32_t)C;                 // Down : (x, y-1, z)
        m[4] = cen & ~(uint32_t)(o[1] >> 1);       // North: (x, y, z+1)
        m[5] = cen & ~(uint32_t)(o[-1] >> 1);      // South
    } else {
        m[0] = m[1] = m[2] = m[3] = m[4] = m[5] = 0u;
    }
}

// Where a column's six visible-face masks come from: recomputed from the occupancy tile (the state the
// reference's flags converge to), or read from face planes extracted from cube.flags (literal generate_mesh).
struct FacesFromOcc {
    const u64* occ;
    __device__ __forceinline__ void operator()(int x, int z, uint32_t m[6]) const { column_masks(occ, x, z, m); }
};
struct FacesFromPlanes {
    const uint32_t* planes;  // this chunk: [6][1024], word = column x*32+z, bit y
    __device__ __forceinline__ void operator()(int x, int z, uint32_t m[6]) const {
        const uint32_t* p = planes + x * 32 + z;
#pragma unroll
        for (int f = 0; f < 6; ++f) m[f] = __ldg(p + f * HB_CHUNK_AREA);
    }
};

// One window of up to 32 quads, one lane per quad (L = voxel index, f = face, ao = four 2-bit vertex_ao values):
// material, colour draw, 25 words into the warp-private staging buffer `ws` (stride 25: bank-conflict free), then
// one TMA bulk shared->global copy (UBLKCP) of the whole window by lane 0 instead of 200 LDS.128 + STG.128 through
// the LSU.  The window is zero-padded to a multiple of 4 quads.  The wait for the engine to release the buffer from
// this warp's previous window sits after the register work so the two overlap.
// MATS_SMEM: `mats` points into shared memory (k_emit_ids keeps a copy there: its L1 is thrashed by the block-array reads)
template <bool MATS_SMEM = false, class SM>
__device__ __forceinline__ void emit_window(SM& sm, const MaterialsDev* __restrict__ mats, float* __restrict__ ws,
                                            uint32_t ws_shared, char* __restrict__ gdst, uint32_t nw, uint32_t L, uint32_t f,
                                            uint32_t ao, uint32_t id, int bx, int by, int bz) {
    const int lane = threadIdx.x & 31;
    const uint32_t nw4 = (nw + 3u) & ~3u;
    float* sq = ws + lane * 25;
    const bool active = (uint32_t)lane < nw;
    const int x = L >> 10, z = (L >> 5) & 31, y = L & 31;
    float4 rgba = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        // Material::get_color(perms[face]) (server/src/chunk/material.rs:67-71)
        const float p = wyrand_f32_at(sm.rng_seed, 6u * L + f);
        // ids above the palette are reported through total[2] & 2 by the pack stage; the device-pointer entry points
        // emit regardless, so keep the table reads in bounds (such a quad gets material 1's colours)
        const uint32_t mi = id - 1u < (uint32_t)HB_MAX_MATERIALS ? id - 1u : 0u;
        const int nc = MATS_SMEM ? mats->n_colors[mi] : __ldg(&mats->n_colors[mi]);
        const int ci = __float2int_rz(__fmul_rn(__int2float_rn(nc > 0 ? nc - 1 : 0), p));
        const float4* const col = reinterpret_cast<const float4*>(mats->rgba[mi][ci]);
        rgba = MATS_SMEM ? *col : __ldg(col);
    }
    if (lane == 0) bulk_wait_read();  // previous window of this warp has left the staging buffer
    __syncwarp();
    if (active) {
        const float4* fm = reinterpret_cast<const float4*>(sm.face_mat[f]);
        const float4 m0 = fm[0], m1 = fm[1], m2 = fm[2];
        sq[0] = m0.x; sq[1] = m0.y; sq[2] = m0.z; sq[3] = m0.w;
        sq[4] = m1.x; sq[5] = m1.y; sq[6] = m1.z; sq[7] = m1.w;
        sq[8] = m2.x; sq[9] = m2.y; sq[10] = m2.z; sq[11] = m2.w;
        sq[12] = __int2float_rn(bx + x);  // (chunk_position + position).cast::<f32>()
        sq[13] = __int2float_rn(by + y);
        sq[14] = __int2float_rn(bz + z);
        sq[15] = 1.0f;
        sq[16] = rgba.x; sq[17] = rgba.y; sq[18] = rgba.z; sq[19] = rgba.w;
        sq[20] = __uint_as_float(0u);  // light = 0 (chunk.rs:207)
        sq[21] = sm.ao_lut[ao & 3u];
        sq[22] = sm.ao_lut[(ao >> 2) & 3u];
        sq[23] = sm.ao_lut[(ao >> 4) & 3u];
        sq[24] = sm.ao_lut[(ao >> 6) & 3u];
    } else if ((uint32_t)lane < nw4) {
#pragma unroll
        for (int i = 0; i < 25; ++i) sq[i] = 0.0f;
    }
    fence_proxy_async_smem();  // generic-proxy writes must be fenced for the async proxy
    __syncwarp();
    if (lane == 0) bulk_store(gdst, ws_shared, nw4 * 100u);
}

// Meshes the occupancy tile in sm.occ (already built and synchronised).  Column mapping: thread t
// owns columns (x = k*8 + t/32, z = t%32), k = 0..3, so a warp reads 32 consecutive tile words.
// Returns the chunk's quad count (all threads).  EMIT additionally writes the instances at `out`
// (16-byte aligned), zero-filling up to the next multiple of 4 instances.
//
// PACKED (with EMIT): instead of the 100-byte instance, `out` receives one 32-bit word per quad (hb_packed_quad:
// L | face << 15 | ao << 18 | (material - 1) << 26), the gap zero-filled the same way.  The colour is not part of
// the word: it is a pure function of (chunk, L, face) and the expander recomputes it (expand.cpp).
template <bool EMIT, bool PACKED = false, bool MATS_SMEM = false, class SM, class MatFn, class FacesFn>
__device__ __forceinline__ uint32_t mesh_tile(SM& sm, const MatFn& mat, const FacesFn& faces, int bx, int by, int bz,
                                              const MaterialsDev* __restrict__ mats, float* __restrict__ out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t cnt[4], incl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t m[6];
        faces(k * 8 + warp, lane, m);
        const uint32_t c = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]) + __popc(m[4]) + __popc(m[5]);
        cnt[k] = c;
        incl[k] = warp_incl_scan(c, lane);
        if (lane == 31) sm.warp_tot[k * 8 + warp] = incl[k];
    }
    __syncthreads();
    // 32 (k, warp) totals in column order; every warp scans them redundantly (no second barrier)
    const uint32_t wt = sm.warp_tot[lane];
    const uint32_t ws = warp_incl_scan(wt, lane);
    const uint32_t total = __shfl_sync(0xffffffffu, ws, 31);
    if constexpr (EMIT) {
    uint32_t base[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = k * 8 + warp;
        base[k] = __shfl_sync(0xffffffffu, ws, idx) - __shfl_sync(0xffffffffu, wt, idx) + incl[k] - cnt[k];
    }

    float* const ws = sm.stage + warp * (32 * 25);  // warp-private staging: 32 quads = 3200 B
    const uint32_t ws_shared = (uint32_t)__cvta_generic_to_shared(ws);
    char* const out_bytes = reinterpret_cast<char*>(out);
    for (uint32_t d0 = 0; d0 < total; d0 += SM::kDesc) {
        // (a) visible faces of my columns -> descriptors (L | face << 15), in emission order
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t o = base[k];
            if (cnt[k] == 0 || o >= d0 + SM::kDesc || o + cnt[k] <= d0) continue;
            const uint32_t Lc = (uint32_t)(((k * 8 + warp) << 10) | (lane << 5));
            uint32_t m[6];
            faces(k * 8 + warp, lane, m);  // cheap to re-evaluate; keeps registers free
            uint32_t any = m[0] | m[1] | m[2] | m[3] | m[4] | m[5];
            while (any) {
                const int y = __ffs(any) - 1;
                any &= any - 1;
                uint32_t b = ((m[0] >> y) & 1u) | (((m[1] >> y) & 1u) << 1) | (((m[2] >> y) & 1u) << 2) |
                             (((m[3] >> y) & 1u) << 3) | (((m[4] >> y) & 1u) << 4) | (((m[5] >> y) & 1u) << 5);
                const uint32_t Ly = Lc | (uint32_t)y;
                while (b) {  // faces of this voxel in bit order E, W, U, D, N, S
                    const uint32_t f = (uint32_t)__ffs(b) - 1u;
                    b &= b - 1u;
                    if (o - d0 < (uint32_t)SM::kDesc) sm.desc[o - d0] = Ly | (f << 15);  // unsigned: also rejects o < d0
                    ++o;
                }
            }
        }
        __syncthreads();
        const uint32_t nd = min((uint32_t)SM::kDesc, total - d0);
        if constexpr (PACKED) {
            // (b') one thread per quad: AO probes + material -> one coalesced 32-bit store
            uint32_t* const o32 = reinterpret_cast<uint32_t*>(out) + d0;
            const uint32_t nd4 = (d0 + nd == total) ? ((nd + 3u) & ~3u) : nd;  // zero-fill the gap after the last quad
            for (uint32_t q = tid; q < nd4; q += kT) {
                uint32_t w = 0u;
                if (q < nd) {
                    const uint32_t d = sm.desc[q];
                    const uint32_t L = d & 0x7FFFu, f = d >> 15;
                    const uint32_t ao = ao_bits(sm.occ, sm.ao_tab[f], L >> 10, L & 31, (L >> 5) & 31);
                    const uint32_t id = mat(L, L >> 10, L & 31, (L >> 5) & 31);
                    w = d | (ao << 18) | (((id - 1u) & 15u) << 26);
                }
                __stcs(o32 + q, w);
            }
        } else {
        // (b) each warp expands windows of 32 descriptors (one quad per lane) and streams them out; no block-level
        //     barrier inside this loop
        // The descriptor and the material of the NEXT window are fetched before the current one is expanded, so a
        // block-array material read (a scattered 2-byte global load in the general mesher) overlaps a whole window.
        uint32_t d_next = 0, id_next = 0;
        if (warp * 32u + lane < nd) {
            d_next = sm.desc[warp * 32u + lane];
            const uint32_t Ln = d_next & 0x7FFFu;
            id_next = mat(Ln, Ln >> 10, Ln & 31, (Ln >> 5) & 31);
        }
        for (uint32_t q0 = warp * 32u; q0 < nd; q0 += (kT / 32) * 32u) {
            const uint32_t nw = min(32u, nd - q0);
            const uint32_t d = d_next, id = id_next;
            const uint32_t qn = q0 + (kT / 32) * 32u + lane;
            if (qn < nd) {
                d_next = sm.desc[qn];
                const uint32_t Ln = d_next & 0x7FFFu;
                id_next = mat(Ln, Ln >> 10, Ln & 31, (Ln >> 5) & 31);
            }
            const uint32_t L = d & 0x7FFFu, f = d >> 15;
            uint32_t ao = 0;
            if ((uint32_t)lane < nw) ao = ao_bits(sm.occ, sm.ao_tab[f], L >> 10, L & 31, (L >> 5) & 31);
            emit_window<MATS_SMEM>(sm, mats, ws, ws_shared, out_bytes + (d0 + q0) * 100u, nw, L, f, ao, id, bx, by, bz);  // < 2^24 B / chunk
        }
        }  // !PACKED
        if (d0 + SM::kDesc < total) __syncthreads();  // descriptors are rewritten in the next round
    }
    }  // if constexpr (EMIT)
    return total;
}

template <class SM>
__device__ __forceinline__ void load_tables(SM& sm, const MeshTables& tb) {
    for (int i = threadIdx.x; i < 72; i += kT) (&sm.face_mat[0][0])[i] = (&tb.face_mat[0][0])[i];
    if (threadIdx.x < 4) sm.ao_lut[threadIdx.x] = tb.ao_lut[threadIdx.x];
    if (threadIdx.x >= 64 && threadIdx.x < 112) {
        const int e = threadIdx.x - 64;
        sm.ao_tab[e >> 3][e & 7] = ao_tab_entry(e >> 3, e & 7);
    }
}

// ------------------------------------------------------------------------------------------------
// General path: arbitrary block arrays.
// ------------------------------------------------------------------------------------------------

// K3a: u16 ids -> 1 bit per voxel (bitmap[c] for column c = x*32+z, bit y) + a per-chunk summary
// word: bit0 any solid, bit1 all solid, bit(2+f) boundary slab at face f entirely solid.
// One 128-bit streaming load per thread per 8 voxels; the chunk's 64 KiB are read exactly once.
//
// While the words are still on chip the kernel also does the chunk-local half of the COUNT stage: the faces between
// two voxels of this chunk (`counts[chunk]`), and a 768-byte border record (six 32-word planes, face order
// E W U D N S: the x = 31 / x = 0 rows by z, the y = 31 / y = 0 bits as bit z of word x, the z = 31 / z = 0 words
// by x) from which k_count_border adds the faces on the chunk's hull without touching the bitmaps again.
constexpr int kBorderWords = 6 * 32;
__global__ void __launch_bounds__(kT, 3) k_pack(const uint16_t* __restrict__ ids, size_t n, int n_materials,
                                             uint32_t* __restrict__ bitmaps, uint32_t* __restrict__ summary,
                                             uint32_t* __restrict__ borders, uint32_t* __restrict__ counts,
                                             u64* __restrict__ total) {
    __shared__ uint32_t s_red[8][8];
    __shared__ uint32_t s_bm[HB_CHUNK_AREA];
    __shared__ uint32_t s_cnt[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (size_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(ids + chunk * HB_CHUNK_VOLUME);
        uint32_t any = 0, all = ~0u, fE = ~0u, fW = ~0u, fN = ~0u, fS = ~0u, mx = 0;
        // Two ids per 32-bit word, tested with the packed 16-bit min/max of the DPX unit (VIMNMX3.U16x2):
        //   min(id, 1) is the "non-zero" bit of each halfword, the running max is compared with the palette size once
        //   per chunk.
        const uint32_t one = 0x00010001u;
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int u = it * kT + tid;
            const uint4 v = __ldg(src + u);
            const uint32_t y0 = __vimin3_u16x2(v.x, one, one), y1 = __vimin3_u16x2(v.y, one, one);
            const uint32_t y2 = __vimin3_u16x2(v.z, one, one), y3 = __vimin3_u16x2(v.w, one, one);
            const uint32_t t = y0 + y1 * 4u + y2 * 16u + y3 * 64u;  // low id of word k -> bit 2k, high id -> bit 16 + 2k
            mx = __vimax3_u16x2(mx, v.x, v.y);
            mx = __vimax3_u16x2(mx, v.z, v.w);
            const uint32_t m8 = (t & 0x55u) | ((t >> 15) & 0xAAu);  // bit j = id j of this 8-id unit is non-zero
            uint32_t w = m8 << ((u & 3) * 8);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            const int c = u >> 2, x = c >> 5, z = c & 31;
            if ((lane & 3) == 0) {
                bitmaps[chunk * HB_CHUNK_AREA + c] = w;
                s_bm[c] = w;
            }
            any |= w;
            all &= w;
            if (x == 31) fE &= w;
            if (x == 0) fW &= w;
            if (z == 31) fN &= w;
            if (z == 0) fS &= w;
        }
        any = __reduce_or_sync(0xffffffffu, any);
        all = __reduce_and_sync(0xffffffffu, all);
        fE = __reduce_and_sync(0xffffffffu, fE);
        fW = __reduce_and_sync(0xffffffffu, fW);
        fN = __reduce_and_sync(0xffffffffu, fN);
        fS = __reduce_and_sync(0xffffffffu, fS);
        const uint32_t bad =
            __reduce_or_sync(0xffffffffu, (uint32_t)((mx & 0xFFFFu) > (uint32_t)n_materials || (mx >> 16) > (uint32_t)n_materials));
        if (lane == 0) {
            s_red[warp][0] = any; s_red[warp][1] = all; s_red[warp][2] = fE; s_red[warp][3] = fW;
            s_red[warp][4] = fN; s_red[warp][5] = fS; s_red[warp][6] = bad;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = 0, l = ~0u, e = ~0u, w_ = ~0u, nn = ~0u, s = ~0u, b = 0;
            for (int i = 0; i < 8; ++i) {
                a |= s_red[i][0]; l &= s_red[i][1]; e &= s_red[i][2]; w_ &= s_red[i][3];
                nn &= s_red[i][4]; s &= s_red[i][5]; b |= s_red[i][6];
            }
            uint32_t sw = (a ? 1u : 0u) | (l == ~0u ? 2u : 0u);
            sw |= (e == ~0u ? 1u : 0u) << 2;          // East slab (x = 31)
            sw |= (w_ == ~0u ? 1u : 0u) << 3;         // West slab (x = 0)
            sw |= ((l >> 31) & 1u) << 4;              // Up slab (y = 31 of every column)
            sw |= (l & 1u) << 5;                      // Down slab (y = 0)
            sw |= (nn == ~0u ? 1u : 0u) << 6;         // North slab (z = 31)
            sw |= (s == ~0u ? 1u : 0u) << 7;          // South slab (z = 0)
            summary[chunk] = sw;
            if (b) atomicOr(total + 2, 2ULL);         // error path only: unknown material id
        }
        // chunk-local faces: a voxel's face is visible when the voxel behind it is empty; on the hull that voxel
        // belongs to another chunk, so it is treated as solid here and k_count_border settles it
        uint32_t* rec = borders + chunk * kBorderWords;
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = k * 8 + warp, z = lane, col = x * 32 + z;
            const uint32_t cen = s_bm[col];
            const uint32_t e = x < 31 ? s_bm[col + 32] : ~0u;
            const uint32_t w = x > 0 ? s_bm[col - 32] : ~0u;
            const uint32_t nn = z < 31 ? s_bm[col + 1] : ~0u;
            const uint32_t ss = z > 0 ? s_bm[col - 1] : ~0u;
            const uint32_t up = (cen >> 1) | 0x80000000u, dn = (cen << 1) | 1u;
            cnt += __popc(cen & ~e) + __popc(cen & ~w) + __popc(cen & ~up) + __popc(cen & ~dn) + __popc(cen & ~nn) +
                   __popc(cen & ~ss);
            const uint32_t top = __ballot_sync(0xffffffffu, cen >> 31), bot = __ballot_sync(0xffffffffu, cen & 1u);
            if (lane == 0) {
                rec[2 * 32 + x] = top;
                rec[3 * 32 + x] = bot;
            }
        }
        if (tid < 32) rec[tid] = s_bm[31 * 32 + tid];                         // E: x = 31, by z
        else if (tid < 64) rec[tid] = s_bm[tid - 32];                         // W: x = 0, by z
        else if (tid >= 128 && tid < 160) rec[tid] = s_bm[(tid - 128) * 32 + 31];  // N: z = 31, by x
        else if (tid >= 160 && tid < 192) rec[tid] = s_bm[(tid - 160) * 32];       // S: z = 0, by x
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) s_cnt[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += s_cnt[i];
            counts[chunk] = t;
        }
    }
}

// COUNT, second half (ids path): the faces on the hull of every chunk, from the border records of the chunk and of its six
// face neighbours (shell index (dx+1)*9 + (dy+1)*3 + (dz+1), client/src/world/chunk.rs:158-174); a missing
// neighbour is empty space.  One warp per chunk, 12 coalesced 128-byte loads.  Also appends the chunks that have
// quads to nz[1..] (nz[0] = how many; zeroed by the launch wrapper): the work list of k_emit_ids.  The list's order
// depends on CTA arrival; it only decides which CTA emits a chunk, never where the quads go.
__global__ void __launch_bounds__(kT) k_count_border(size_t n, const int32_t* __restrict__ nbr, const uint32_t* __restrict__ borders,
                                                      uint32_t* __restrict__ counts, uint32_t* __restrict__ nz) {
    __shared__ uint32_t s_tot[kT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t chunk = (size_t)blockIdx.x * (kT / 32) + warp;
    uint32_t tot = 0;
    if (chunk < n) {
        const int32_t* nb = nbr + chunk * 27;
        const int shell[6] = {22, 4, 16, 10, 14, 12};  // E W U D N S
        const uint32_t* rec = borders + chunk * kBorderWords;
        uint32_t c = 0;
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            const int o = __ldg(nb + shell[f]);
            const uint32_t own = __ldg(rec + f * 32 + lane);
            const uint32_t other = o >= 0 ? __ldg(borders + (size_t)o * kBorderWords + (f ^ 1) * 32 + lane) : 0u;
            c += __popc(own & ~other);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        tot = counts[chunk] + c;  // k_pack's chunk-local faces + the hull
        if (lane == 0 && c) counts[chunk] = tot;
    }
    if (lane == 0) s_tot[warp] = tot;
    __syncthreads();
    if (warp == 0) {
        const bool has = lane < kT / 32 && s_tot[lane] != 0;
        const uint32_t m = __ballot_sync(0xffffffffu, has);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(nz, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (has) nz[1 + base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(blockIdx.x * (kT / 32) + lane);
    }
}

// K3a for PaletteCube input (literal generate_mesh): occupancy bitmap (material.is_some()) + summary as k_pack,
// plus six face planes from cube.flags: CubeFlags::faces() = (value & 1) ? (value >> 1) & 63 : none
// (server/src/chunk/cube.rs:35-41), kept only where the voxel has a material (chunk.rs:187 gates on it).
// `list` (optional): work item j processes chunk list[j] instead of chunk j (device-resident world: only the
// slots whose cubes changed are re-packed).
__global__ void __launch_bounds__(kT) k_pack_cubes(const hb_palette_cube* __restrict__ cubes, size_t n, int n_materials,
                                                   uint32_t* __restrict__ bitmaps, uint32_t* __restrict__ summary,
                                                   uint32_t* __restrict__ planes, u64* __restrict__ total,
                                                   const int32_t* __restrict__ list) {
    __shared__ uint32_t s_red[8][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (size_t item = blockIdx.x; item < n; item += gridDim.x) {
        const size_t chunk = list ? (size_t)list[item] : item;
        const uint4* src = reinterpret_cast<const uint4*>(cubes + chunk * HB_CHUNK_VOLUME);
        uint32_t any = 0, bad = 0;
        for (int it = 0; it < 16; ++it) {
            const int u = it * kT + tid;  // 8-voxel unit (64 bytes): column u>>2, y0 = (u&3)*8
            uint32_t m8 = 0, f8[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 v = __ldg(src + u * 4 + j);  // two cubes: {material|pad, flags, material|pad, flags}
                const uint32_t mat0 = v.x & 0xFFFFu, mat1 = v.z & 0xFFFFu;
                const uint32_t fa0 = (mat0 && (v.y & 1u)) ? (v.y >> 1) & 63u : 0u;
                const uint32_t fa1 = (mat1 && (v.w & 1u)) ? (v.w >> 1) & 63u : 0u;
                bad |= (mat0 > (uint32_t)n_materials) | (mat1 > (uint32_t)n_materials);
                m8 |= ((mat0 ? 1u : 0u) | (mat1 ? 2u : 0u)) << (2 * j);
#pragma unroll
                for (int f = 0; f < 6; ++f) f8[f] |= (((fa0 >> f) & 1u) | (((fa1 >> f) & 1u) << 1)) << (2 * j);
            }
            const int sh = (u & 3) * 8, c = u >> 2;
            uint32_t w = m8 << sh;
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            if ((lane & 3) == 0) bitmaps[chunk * HB_CHUNK_AREA + c] = w;
            any |= w;
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                uint32_t pw = f8[f] << sh;
                pw |= __shfl_xor_sync(0xffffffffu, pw, 1);
                pw |= __shfl_xor_sync(0xffffffffu, pw, 2);
                if ((lane & 3) == 0) planes[(chunk * 6 + f) * HB_CHUNK_AREA + c] = pw;
            }
        }
        any = __reduce_or_sync(0xffffffffu, any);
        bad = __reduce_or_sync(0xffffffffu, bad);
        if (lane == 0) { s_red[warp][0] = any; s_red[warp][1] = bad; }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = 0, b = 0;
            for (int i = 0; i < 8; ++i) { a |= s_red[i][0]; b |= s_red[i][1]; }
            summary[chunk] = a ? 1u : 0u;  // only "any solid" is used in flags mode
            if (b) atomicOr(total + 2, 2ULL);  // error path only: unknown material id
        }
        __syncthreads();
    }
}

struct MatIds {
    const uint16_t* ids;  // centre chunk
    __device__ __forceinline__ uint32_t operator()(uint32_t L, int, int, int) const { return __ldg(ids + L); }
};
struct MatCubes {
    const hb_palette_cube* cubes;  // centre chunk, reference PaletteCube layout
    __device__ __forceinline__ uint32_t operator()(uint32_t L, int, int, int) const { return __ldg(&cubes[L].material); }
};

// K3b (COUNT) for PaletteCube input: the visible faces are the bits of the six planes k_pack_cubes wrote.
__global__ void __launch_bounds__(kT) k_count_ids(size_t n, const uint32_t* __restrict__ summary, const uint32_t* __restrict__ planes,
                                                   uint32_t* __restrict__ counts, const int32_t* __restrict__ list) {
    __shared__ uint32_t s_part[kT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (size_t item = blockIdx.x; item < n; item += gridDim.x) {
        const size_t chunk = list ? (size_t)list[item] : item;  // counts[] is indexed by work item
        if (!(summary[chunk] & 1u)) {  // no solid voxel
            if (tid == 0) counts[item] = 0;
            continue;
        }
        uint32_t c = 0;
        const uint32_t* p = planes + chunk * 6 * HB_CHUNK_AREA;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int f = 0; f < 6; ++f) c += __popc(__ldg(p + f * HB_CHUNK_AREA + (k * 8 + warp) * 32 + lane));
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) s_part[warp] = c;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int i = 0; i < kT / 32; ++i) t += s_part[i];
            counts[item] = t;
        }
        __syncthreads();
    }
}

// FLAGS: `blocks` are hb_palette_cube arrays and the visible faces are taken from `planes` (cube.flags) instead of
// being recomputed from occupancy.
template <bool EMIT, bool FLAGS>  // EMIT is always true since COUNT moved to k_count_ids; kept for mesh_tile's signature
__global__ void __launch_bounds__(kT, kCtasPerSm)
k_mesh_ids(const MeshTables tb, const MaterialsDev* __restrict__ mats, const hb_chunk_pt* __restrict__ pts, size_t n,
           const void* __restrict__ blocks, const int32_t* __restrict__ nbr, const uint32_t* __restrict__ bitmaps,
           const uint32_t* __restrict__ summary, const uint32_t* __restrict__ planes, uint32_t* __restrict__ counts,
           const u64* __restrict__ offsets, float* __restrict__ out, u64 capacity, u64* __restrict__ total,
           const int32_t* __restrict__ list, uint32_t batch) {
    __shared__ MeshSmem sm;
    __shared__ unsigned long long s_ticket;
    const int tid = threadIdx.x;
    if (EMIT) load_tables(sm, tb);
    // Work distribution: chunks differ wildly in cost (three quarters of a terrain batch emit nothing), so the CTAs
    // draw batches of `batch` consecutive work items from a ticket counter (total[3], zeroed by the launch wrapper)
    // instead of striding statically.  This only decides WHICH CTA meshes a chunk; where its quads go is fixed by the
    // scanned offsets, so the output stays deterministic and the face compaction itself uses no atomics.
    for (;;) {
        __syncthreads();  // previous batch's readers of s_ticket / sm are done
        if (tid == 0) s_ticket = atomicAdd(total + 3, (u64)batch);
        __syncthreads();
        const size_t first = (size_t)s_ticket;
        if (first >= n) break;
        const size_t last = first + batch < n ? first + batch : n;
    for (size_t item = first; item < last; ++item) {
        // block-uniform early out: the COUNT stage already knows which chunks have no visible face at all
        // (all air, or buried: k_count_ids), so EMIT needs one load per chunk to skip them
        const uint32_t my_count = counts[item];
        if (my_count == 0) continue;
        const u64 off = offsets[item];
        const size_t chunk = list ? (size_t)list[item] : item;  // counts[] / offsets[] are indexed by work item
        if (off + (((u64)my_count + 3ULL) & ~3ULL) > capacity) {
            if (tid == 0) atomicOr(total + 2, 1ULL);  // error path only
            continue;
        }
        __syncthreads();  // previous chunk's readers of sm.nbr / sm.occ are done
        if (tid < 27) sm.nbr[tid] = nbr[chunk * 27 + tid];
        const hb_chunk_pt pt = pts[chunk];
        if (EMIT && tid == 32) sm.rng_seed = siphash13_chunk(pt.x, pt.y, pt.z);
        __syncthreads();
        // all (up to 15) bitmap words of this thread's tile elements are requested before any is used, so the
        // L2/HBM latency is paid once per chunk instead of once per element
        constexpr int kIter = (kTileN + kT - 1) / kT;
        uint32_t w0[kIter], wd[kIter], wu[kIter];
#pragma unroll
        for (int it = 0; it < kIter; ++it) {
            const int i = tid + it * kT;
            w0[it] = wd[it] = wu[it] = 0u;
            if (i < kTileN) {
                const int xi = i / kTile, zi = i - xi * kTile;
                const int dx = (xi == 0) ? -1 : (xi == kTile - 1 ? 1 : 0);
                const int dz = (zi == 0) ? -1 : (zi == kTile - 1 ? 1 : 0);
                const int c = ((xi - 1) & 31) * 32 + ((zi - 1) & 31);
                const int sb = (dx + 1) * 9 + (dz + 1);
                const int n0 = sm.nbr[sb + 3];  // dy = 0
                const int nd = sm.nbr[sb];      // dy = -1: its y = 31
                const int nu = sm.nbr[sb + 6];  // dy = +1: its y = 0
                if (n0 >= 0) w0[it] = __ldg(bitmaps + (size_t)n0 * HB_CHUNK_AREA + c);
                if (nd >= 0) wd[it] = __ldg(bitmaps + (size_t)nd * HB_CHUNK_AREA + c);
                if (nu >= 0) wu[it] = __ldg(bitmaps + (size_t)nu * HB_CHUNK_AREA + c);
            }
        }
#pragma unroll
        for (int it = 0; it < kIter; ++it) {
            const int i = tid + it * kT;
            if (i < kTileN) sm.occ[i] = ((u64)w0[it] << 1) | (u64)(wd[it] >> 31) | ((u64)(wu[it] & 1u) << 33);
        }
        __syncthreads();
        uint32_t q;
        if (FLAGS) {
            const MatCubes mat{static_cast<const hb_palette_cube*>(blocks) + chunk * HB_CHUNK_VOLUME};
            q = mesh_tile<EMIT>(sm, mat, FacesFromPlanes{planes + chunk * 6 * HB_CHUNK_AREA}, pt.x * 32, pt.y * 32, pt.z * 32,
                                mats, EMIT ? out + off * 25ULL : nullptr);
        } else {
            const MatIds mat{static_cast<const uint16_t*>(blocks) + chunk * HB_CHUNK_VOLUME};
            q = mesh_tile<EMIT>(sm, mat, FacesFromOcc{sm.occ}, pt.x * 32, pt.y * 32, pt.z * 32, mats,
                                EMIT ? out + off * 25ULL : nullptr);
        }
        if (!EMIT && tid == 0) counts[item] = q;
    }
    }  // ticket loop
    if (EMIT) bulk_wait_all();  // every bulk store issued by this thread has landed before the CTA retires
}

// ---- EMIT for the ids path, software-pipelined across chunks --------------------------------------------------------
// A chunk's EMIT starts with a chain of dependent loads (which chunk -> its neighbour row -> 1156 + 576 occupancy
// words) before the first quad can be written; in k_mesh_ids that chain is paid in full by every chunk (23 % of the
// kernel's stall samples sit on the first use of the tile words, profiles/r2_ncu_c3emit_summary.csv).  Here the
// chain of chunk i+1 runs underneath the emission of chunk i:
//   * COUNT leaves a compacted list of the chunks that have quads (k_count_border), so one ticket = one chunk;
//   * warp 0 keeps the control pipeline in registers: the ticket of chunk i+2 is drawn and the neighbour row, count,
//     offset and position of chunk i+1 are loaded while the CTA still waits for chunk i's tile;
//   * the tile of chunk i+1 is fetched with cp.async (no registers, no scoreboard) into a raw buffer as soon as the
//     occupancy tile of chunk i has been built from it, and lands during chunk i's emission.
// The words above and below the tile come from the 128-byte U/D planes of the border records instead of one full
// bitmap word per column.  Which CTA meshes a chunk is decided by tickets; where its quads go is fixed by the scanned
// offsets, so the output is deterministic.
struct IdsItem {
    u64 off;
    uint32_t count;
    int32_t chunk;  // < 0: no such item
    hb_chunk_pt pt;
    int32_t nbr[27];
};
struct IdsPipe {
    uint32_t pl[2][9][32];  // [0]: U planes (y = 31) of the 9 chunks below, [1]: D planes (y = 0) of the 9 chunks above
    IdsItem item[2];
};

struct EmitIdsSmem {
    MeshSmem sm;
    IdsPipe pp;
    MaterialsDev mats;
};
static_assert((sizeof(EmitIdsSmem) + 1024) * kCtasPerSm <= 228 * 1024, "k_emit_ids: shared memory per SM");

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// tile cell i of a 34x34 halo tile -> (shell index of the lateral neighbour at dy = -1, 3x3 index, column in that chunk)
__device__ __forceinline__ void tile_cell(int i, int& sb, int& nb9, int& x, int& z) {
    const int xi = i / kTile, zi = i - xi * kTile;
    const int dx = (xi == 0) ? -1 : (xi == kTile - 1 ? 1 : 0);
    const int dz = (zi == 0) ? -1 : (zi == kTile - 1 ? 1 : 0);
    x = (xi - 1) & 31;
    z = (zi - 1) & 31;
    sb = (dx + 1) * 9 + (dz + 1);
    nb9 = (dx + 1) * 3 + (dz + 1);
}

__global__ void __launch_bounds__(kT, kCtasPerSm)
k_emit_ids(const MeshTables tb, const MaterialsDev* __restrict__ mats, const hb_chunk_pt* __restrict__ pts,
           const uint16_t* __restrict__ ids, const int32_t* __restrict__ nbr, const uint32_t* __restrict__ bitmaps,
           const uint32_t* __restrict__ borders, const uint32_t* __restrict__ nz, const uint32_t* __restrict__ counts,
           const u64* __restrict__ offsets, float* __restrict__ out, u64 capacity, u64* __restrict__ total) {
    extern __shared__ __align__(16) unsigned char emit_ids_smem_raw[];
    EmitIdsSmem& es = *reinterpret_cast<EmitIdsSmem*>(emit_ids_smem_raw);
    MeshSmem& sm = es.sm;
    IdsPipe& pp = es.pp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* const raw = reinterpret_cast<uint32_t*>(sm.hts);  // dy = 0 words of the tile being fetched (region kernels keep heights here)
    load_tables(sm, tb);
    for (int i = tid; i < (int)(sizeof(MaterialsDev) / sizeof(uint32_t)); i += kT)
        reinterpret_cast<uint32_t*>(&es.mats)[i] = __ldg(reinterpret_cast<const uint32_t*>(mats) + i);
    const uint32_t n_nz = __ldg(nz);
    const uint32_t* const list = nz + 1;

    // warp 0: this lane's part of item `chunk` (27 neighbour entries, count, offset, position)
    struct Piece { u64 off; int32_t v; };
    auto load_piece = [&](int32_t chunk) -> Piece {
        Piece p{0, 0};
        if (chunk < 0) return p;
        if (lane < 27) p.v = __ldg(nbr + (size_t)chunk * 27 + lane);
        else if (lane == 27) p.v = (int32_t)__ldg(counts + chunk);
        else if (lane == 28) p.off = __ldg(offsets + chunk);
        else p.v = __ldg(reinterpret_cast<const int32_t*>(pts + chunk) + (lane - 29));
        return p;
    };
    auto store_piece = [&](IdsItem& it, int32_t chunk, const Piece& p) {
        if (lane == 0) it.chunk = chunk;
        if (chunk < 0) return;
        if (lane < 27) it.nbr[lane] = p.v;
        else if (lane == 27) it.count = (uint32_t)p.v;
        else if (lane == 28) it.off = p.off;
        else (&it.pt.x)[lane - 29] = p.v;
    };
    // all threads: start fetching the tile of item `it` (block-uniform; it.nbr is visible)
    auto fetch_tile = [&](const IdsItem& it) {
        constexpr int kIter = (kTileN + kT - 1) / kT;
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
            const int i = tid + k * kT;
            if (i < kTileN) {
                int sb, nb9, x, z;
                tile_cell(i, sb, nb9, x, z);
                const int n0 = it.nbr[sb + 3];  // dy = 0
                if (n0 >= 0) cp_async4(raw + i, bitmaps + (size_t)n0 * HB_CHUNK_AREA + x * 32 + z);
                else raw[i] = 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = tid + k * kT;
            if (e < 2 * 9 * 32) {
                const int v = e / 288, r = e - v * 288, nb9 = r >> 5, w = r & 31;
                const int nbc = it.nbr[(nb9 / 3) * 9 + (nb9 % 3) + (v ? 6 : 0)];
                // the chunk above shows its y = 0 bits (D plane, 3), the chunk below its y = 31 bits (U plane, 2)
                if (nbc >= 0) cp_async4(&pp.pl[v][nb9][w], borders + (size_t)nbc * kBorderWords + (v ? 3 : 2) * 32 + w);
                else pp.pl[v][nb9][w] = 0u;
            }
        }
    };

    // prologue: the first two tickets; the first item's record and tile
    int32_t next_chunk = -1;  // warp 0 (uniform): the item after the current one, not yet loaded
    if (warp == 0) {
        u64 t = 0;
        if (lane == 0) t = atomicAdd(total + 3, 2ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        const int32_t first = t < n_nz ? (int32_t)__ldg(list + t) : -1;
        next_chunk = t + 1 < n_nz ? (int32_t)__ldg(list + t + 1) : -1;
        store_piece(pp.item[0], first, load_piece(first));
    }
    __syncthreads();
    int cur = 0;
    if (pp.item[0].chunk >= 0) fetch_tile(pp.item[0]);
    while (pp.item[cur].chunk >= 0) {
        // (1) warp 0: the next item's record and the ticket after it go out while this item's tile is still in flight
        Piece piece{0, 0};
        u64 ticket = 0;
        if (warp == 0) {
            piece = load_piece(next_chunk);
            if (lane == 0) ticket = atomicAdd(total + 3, 1ULL);
        }
        // (2) this item's tile has landed
        cp_async_wait_all();
        __syncthreads();
        const IdsItem& it = pp.item[cur];
        constexpr int kIter = (kTileN + kT - 1) / kT;
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
            const int i = tid + k * kT;
            if (i < kTileN) {
                int sb, nb9, x, z;
                tile_cell(i, sb, nb9, x, z);
                const u64 below = (pp.pl[0][nb9][x] >> z) & 1u, above = (pp.pl[1][nb9][x] >> z) & 1u;
                sm.occ[i] = ((u64)raw[i] << 1) | below | (above << 33);
            }
        }
        if (warp == 0) store_piece(pp.item[cur ^ 1], next_chunk, piece);
        if (tid == 32) sm.rng_seed = siphash13_chunk(it.pt.x, it.pt.y, it.pt.z);
        __syncthreads();  // occupancy tile complete, raw buffers free, next record visible
        // (3) the next item's tile is fetched underneath this item's emission
        if (pp.item[cur ^ 1].chunk >= 0) fetch_tile(pp.item[cur ^ 1]);
        if (warp == 0) {  // identity of the item after the next (consumed at the top of the next iteration)
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            next_chunk = ticket < n_nz ? (int32_t)__ldg(list + ticket) : -1;
        }
        // (4) mesh
        const u64 off = it.off;
        if (off + (((u64)it.count + 3ULL) & ~3ULL) > capacity) {
            if (tid == 0) atomicOr(total + 2, 1ULL);  // error path only
        } else {
            const hb_chunk_pt pt = it.pt;
            mesh_tile<true, false, true>(sm, MatIds{ids + (size_t)it.chunk * HB_CHUNK_VOLUME}, FacesFromOcc{sm.occ}, pt.x * 32,
                                         pt.y * 32, pt.z * 32, &es.mats, out + off * 25ULL);
        }
        cur ^= 1;
        // the next iteration's barrier (2) also fences this item's readers of sm.occ / item[cur ^ 1] against (1)..(3)
    }
    cp_async_wait_all();
    bulk_wait_all();  // every bulk store issued by this thread has landed before the CTA retires
}

// ------------------------------------------------------------------------------------------------
// Fused path: heightmap terrain straight from the heights tiles (no block-array round trip).
// One work item = one chunk column; the 34x34 heights halo is staged once and reused for all cy.
// ------------------------------------------------------------------------------------------------

struct MatHeights {
    const int* hts;  // relative heights (RelHeight)
    int a;           // the chunk's first voxel layer, relative: iy * 32
    __device__ __forceinline__ uint32_t operator()(uint32_t, int x, int y, int z) const {
        return block_id(hts[(x + 1) * kTile + (z + 1)] - (a + y));
    }
};

// Stages the 34x34 heights halo tile of region column (ix, iz): the column's own 32x32 tile (one
// 128-bit load per thread) plus one row/column from each of the 4 edge neighbours and one value from
// each of the 4 corner neighbours; columns outside the region are ABSENT.  Also returns this
// thread's partial max over the centre, min over the whole tile (hmin) and min over centre + edges
// (emin).  Needs kT == 256.
// Heights relative to the region's lowest voxel (world y = cy0*32), clamped to [-1, 32*ncy + 8] so that all
// per-layer arithmetic stays in 32 bits (the +8 keeps `h - y > 6`, the stone test, exact up to the region's
// top voxel); absent columns (INT_MIN) become -1: never solid.
struct RelHeight {
    long long a0;
    int top;
    __device__ __forceinline__ int operator()(int h) const {
        const long long d = (long long)h - a0;
        return (int)(d < -1 ? -1 : (d > top ? top : d));
    }
};
struct TileRegs {
    int4 centre;  // 4 consecutive z of row tid / 8
    int edge;     // this thread's edge or corner value (threads 0..131)
};

// global -> registers half of the tile staging (raw heights; ABSENT for columns outside the region)
__device__ __forceinline__ TileRegs fetch_heights_tile(const RegionDev& r, const int32_t* __restrict__ heights, int ix, int iz) {
    const int tid = threadIdx.x;
    // the eight neighbour tiles lie at fixed strides from the column's own tile: one 64-bit address per column, the
    // rest is 32-bit offsets (ncz * 1024 < 2^31 within the coordinate limit)
    const int32_t* const c0 = heights + ((size_t)(ix - r.hx_begin) * r.ncz + iz) * HB_CHUNK_AREA;
    const int dxs = r.ncz * HB_CHUNK_AREA;  // elements between the tiles of ix and ix + 1 (iz + 1: HB_CHUNK_AREA)
    TileRegs t;
    t.centre = __ldg(reinterpret_cast<const int4*>(c0) + tid);  // element e = x*32 + z
    t.edge = kHeightAbsent;
    if (tid < 132) {
        int dx, dz, src;
        if (tid < 128) {  // edges: west (x = -1), east (x = 32), south (z = -1), north (z = 32); 32 values each
            const int side = tid >> 5, j = tid & 31;
            dx = side == 0 ? -1 : (side == 1 ? 1 : 0);
            dz = side == 2 ? -1 : (side == 3 ? 1 : 0);
            src = side == 0 ? 31 * 32 + j : side == 1 ? j : side == 2 ? j * 32 + 31 : j * 32;
        } else {  // corners
            const int c = tid - 128;
            dx = (c & 1) ? 1 : -1;
            dz = (c & 2) ? 1 : -1;
            src = (dx < 0 ? 31 : 0) * 32 + (dz < 0 ? 31 : 0);
        }
        // columns outside the region are ABSENT
        if ((unsigned)(ix + dx) < (unsigned)r.ncx && (unsigned)(iz + dz) < (unsigned)r.ncz)
            t.edge = __ldg(c0 + (dx * dxs + dz * HB_CHUNK_AREA + src));
    }
    return t;
}

// registers -> shared half; also this thread's partial max over the centre, min over the whole tile (hmin) and
// min over centre + edges (emin: everything the six face masks look at)
template <class Conv>
__device__ __forceinline__ void store_heights_tile(const TileRegs& t, int* __restrict__ hts, int& cmax, int& hmin, int& emin,
                                                   const Conv& conv) {
    const int tid = threadIdx.x;
    {
        const int4 v = make_int4(conv(t.centre.x), conv(t.centre.y), conv(t.centre.z), conv(t.centre.w));
        const int x = tid >> 3, z = (tid & 7) * 4;
        int* o = hts + (x + 1) * kTile + (z + 1);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        const int mx = max(max(v.x, v.y), max(v.z, v.w)), mn = min(min(v.x, v.y), min(v.z, v.w));
        cmax = max(cmax, mx);
        hmin = min(hmin, mn);
        emin = min(emin, mn);
    }
    if (tid < 128) {
        const int side = tid >> 5, j = tid & 31;
        const int dst = side == 0 ? (j + 1) : side == 1 ? 33 * kTile + (j + 1) : side == 2 ? (j + 1) * kTile : (j + 1) * kTile + 33;
        const int h = conv(t.edge);
        hts[dst] = h;
        hmin = min(hmin, h);
        emin = min(emin, h);
    } else if (tid < 132) {
        const int c = tid - 128;
        const int h = conv(t.edge);
        hts[((c & 1) ? 33 : 0) * kTile + ((c & 2) ? 33 : 0)] = h;
        hmin = min(hmin, h);
    }
}

template <class Conv>
__device__ __forceinline__ void load_heights_tile(const RegionDev& r, const int32_t* __restrict__ heights, int ix, int iz,
                                                  int* __restrict__ hts, int& cmax, int& hmin, int& emin, const Conv& conv) {
    store_heights_tile(fetch_heights_tile(r, heights, ix, iz), hts, cmax, hmin, emin, conv);
}

// Heightmap terrain, closed-form face count (the fused path's COUNT stage).  With solid(y) = y < h the
// six mask popcounts of column_masks() collapse to interval arithmetic.  For a chunk layer starting at
// world y = a, let r = clamp(h - a, 0, 32) be the column's solid voxels inside the chunk and rf the same
// for the lateral neighbour f (absent neighbour -> 0).  Then
//   side f : max(0, r - rf)          (y in [a, a+32) with y < h and y >= hf)
//   up     : r > 0 and (h <= a + 32 or the chunk above is absent)
//   down   : r > 0 and the chunk below is absent            (inside the terrain y-1 is always solid)
// which is exactly what k_mesh_region<EMIT> emits from the occupancy tile (parity tests cover both).
__global__ void __launch_bounds__(kT) k_count_region(const RegionDev r, const int32_t* __restrict__ heights,
                                                      uint32_t* __restrict__ counts, u64* __restrict__ ticket) {
    __shared__ uint32_t s_acc[kT / 32][32];  // per warp, per layer of the current group of 32 layers
    __shared__ int s_h[kTileN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const RelHeight rel{(long long)r.cy0 * 32, 32 * r.ncy + 8};  // converted once while the tile is staged
    const int last = r.ncy - 1;
    const int ncols = (r.ix_end - r.ix_begin) * r.ncz;
    // the next column's tile is fetched into registers while the current one is counted
    __shared__ int s_next;
    TileRegs next{};
    if ((int)blockIdx.x < ncols) next = fetch_heights_tile(r, heights, r.ix_begin + (int)blockIdx.x / r.ncz, (int)blockIdx.x % r.ncz);
    for (int colw = blockIdx.x; colw < ncols;) {
    int cmax = INT32_MIN, hmin = INT32_MAX, emin = INT32_MAX;
    __syncthreads();  // previous column's readers of s_h / s_next are done
    // columns are drawn from a ticket counter (`ticket`, zeroed by the launch wrapper): the per-column cost follows
    // the terrain (how many layers a column's surface spans), and a static stride left a tail of late CTAs
    // (drawing the ticket two columns ahead, to hide the atomic's round trip, changes nothing: 0.353 vs 0.350 ms)
    if (tid == 0) s_next = (int)gridDim.x + (int)atomicAdd(ticket, 1ULL);
    store_heights_tile(next, s_h, cmax, hmin, emin, rel);
    __syncthreads();
    const int coln = s_next;
    if (coln < ncols) next = fetch_heights_tile(r, heights, r.ix_begin + coln / r.ncz, coln % r.ncz);
    for (int iy0 = 0; iy0 < r.ncy; iy0 += 32) {
        const int nl = min(32, r.ncy - iy0);
        s_acc[warp][lane] = 0;
        __syncwarp();
        // Each thread holds four cells ADJACENT in x (rows 4*warp + k, column z = lane): their east/west neighbours
        // are each other, so per layer every clamped height is computed once and the three inner x-pairs cost one
        // |difference| each (max(0, a - b) + max(0, b - a) = |a - b|: the pair's two facing side counts together).
        // Layers below `lo` see a cell and its four neighbours completely solid; layers above `hi` see it empty.  Only
        // the union [wlo, whi] over the warp's 128 cells needs the formula (which is valid for every layer); outside
        // it the only faces are the region's bottom (layer 0, Down) and top (last layer, Up) boundary faces.  One pair
        // of range reductions and one sum reduction per layer serve all four rows.
        int hc[4], hN[4], hS[4], hWest, hEast;
        int lo = INT32_MAX, hi = -1, solid = 0;
        {
            const int* o = s_h + (4 * warp + 1) * kTile + (lane + 1);
            hWest = o[-kTile];
            hEast = o[4 * kTile];
#pragma unroll
            for (int k = 0; k < 4; ++k) { hc[k] = o[k * kTile]; hN[k] = o[k * kTile + 1]; hS[k] = o[k * kTile - 1]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (hc[k] > 0) {
                    const int hW = k ? hc[k - 1] : hWest, hE = k < 3 ? hc[k + 1] : hEast;
                    const int l = min(min(hc[k] - 1, hE), min(min(hW, hN[k]), hS[k]));
                    lo = min(lo, l < 0 ? 0 : l >> 5);
                    hi = max(hi, (hc[k] - 1) >> 5);
                    ++solid;
                }
            }
        }
        const int wlo = __reduce_min_sync(0xffffffffu, lo);
        const int whi = __reduce_max_sync(0xffffffffu, hi);
        const int nsolid = __reduce_add_sync(0xffffffffu, solid);
        const int j1 = min(min(whi, last), iy0 + nl - 1);
        for (int j = max(wlo, iy0); j <= j1; ++j) {  // warp-uniform bounds
            const int a = j * 32;
            int rc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rc[k] = min(max(hc[k] - a, 0), 32);
            // x direction: west of cell 0, the three inner pairs, east of cell 3
            uint32_t v = max(0, rc[0] - min(max(hWest - a, 0), 32)) + abs(rc[0] - rc[1]) + abs(rc[1] - rc[2]) + abs(rc[2] - rc[3]) +
                         max(0, rc[3] - min(max(hEast - a, 0), 32));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (rc[k] > 0) {
                    v += max(0, rc[k] - min(max(hN[k] - a, 0), 32)) + max(0, rc[k] - min(max(hS[k] - a, 0), 32));
                    v += (hc[k] - a <= 32 || j == last) ? 1u : 0u;  // up: real top, or the chunk above is absent
                    v += (j == 0) ? 1u : 0u;                        // down: only where the chunk below is absent
                }
            }
            v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) s_acc[warp][j - iy0] = v;  // the only writer of this slot (zeroed above)
        }
        if (lane == 0 && nsolid) {
            if (wlo > 0 && iy0 == 0) s_acc[warp][0] += nsolid * (last == 0 ? 2 : 1);             // bottom faces (+ top if 1 layer)
            if (last > 0 && last < wlo && last >= iy0 && last < iy0 + nl) s_acc[warp][last - iy0] += nsolid;  // top faces
        }
        __syncthreads();
        if (tid < nl) {
            uint32_t t = 0;
#pragma unroll
            for (int wq = 0; wq < kT / 32; ++wq) t += s_acc[wq][tid];
            counts[(size_t)colw * r.ncy + iy0 + tid] = t;
        }
        __syncthreads();
    }
    colw = coln;
    }  // persistent column loop
}

template <bool EMIT, bool PACKED = false>
__global__ void __launch_bounds__(kT, kCtasPerSm)
k_mesh_region(const MeshTables tb, const MaterialsDev* __restrict__ mats, const RegionDev r,
              const int32_t* __restrict__ heights, uint32_t* __restrict__ counts, const u64* __restrict__ offsets,
              float* __restrict__ out, u64 capacity, u64* __restrict__ total) {
    __shared__ MeshSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (EMIT) load_tables(sm, tb);
    const RelHeight rel{(long long)r.cy0 * 32, 32 * r.ncy + 8};
    const int ncols = (r.ix_end - r.ix_begin) * r.ncz;
    __shared__ int s_col;
    for (int colw = blockIdx.x;;) {
        if (colw >= ncols) break;
        const int ix = r.ix_begin + colw / r.ncz, iz = colw % r.ncz;
        __syncthreads();  // previous column's readers of sm.hts / sm.red are done
        // next column: drawn from a ticket counter (total[3], zeroed by the launch wrapper) so that CTAs that meet
        // cheap columns take more of them; tickets are handed out in column order, which keeps all CTAs writing inside
        // one moving window of the output
        if (tid == 0) s_col = (int)gridDim.x + (int)atomicAdd(total + 3, 1ULL);
        int cmax = INT32_MIN, hmin = INT32_MAX, emin = INT32_MAX;
        const TileRegs tile_regs = fetch_heights_tile(r, heights, ix, iz);
        // The chunk hashes (SipHash-1-3, ~200 dependent 64-bit operations) of all layers of this column, one lane per
        // layer, while the heights tile is in flight -- instead of one thread hashing per chunk in front of a barrier.
        if (EMIT && !PACKED && warp == 2 && lane < min(r.ncy, kSeedLayers)) sm.layer_seed[lane] = siphash13_chunk(r.cx0 + ix, r.cy0 + lane, r.cz0 + iz);
        store_heights_tile(tile_regs, sm.hts, cmax, hmin, emin, rel);
        cmax = __reduce_max_sync(0xffffffffu, cmax);
        hmin = __reduce_min_sync(0xffffffffu, hmin);
        emin = __reduce_min_sync(0xffffffffu, emin);
        if (lane == 0) { sm.red[0][warp] = cmax; sm.red[1][warp] = hmin; sm.red[2][warp] = emin; }
        __syncthreads();
        // every warp folds the 8 per-warp partials with one more warp reduction (lanes 0..7 hold them)
        cmax = __reduce_max_sync(0xffffffffu, lane < 8 ? sm.red[0][lane] : INT32_MIN);
        hmin = __reduce_min_sync(0xffffffffu, lane < 8 ? sm.red[1][lane] : INT32_MAX);
        emin = __reduce_min_sync(0xffffffffu, lane < 8 ? sm.red[2][lane] : INT32_MAX);

        for (int iy = 0; iy < r.ncy; ++iy) {
            const size_t chunk = (size_t)colw * r.ncy + iy;  // slab-relative chunk index
            const int cy = r.cy0 + iy;
            const int a = iy * 32;  // relative y of the chunk's first layer
            // block-uniform early outs: centre all air, or the whole 34^3 neighbourhood solid
            const bool empty = cmax <= a;
            const bool buried = hmin >= a + 33 && iy > 0 && iy < r.ncy - 1;
            if (empty || buried) {
                if (!EMIT && tid == 0) counts[chunk] = 0;
                continue;
            }
            u64 off = 0;
            if (EMIT) {
                off = offsets[chunk];
                const u64 span = ((u64)counts[chunk] + 3ULL) & ~3ULL;
                if (off + span > capacity) {
                    if (tid == 0) atomicOr(total + 2, 1ULL);  // error path only
                    continue;
                }
            }
            if (EMIT && iy == 0 && r.ncy > 1 && emin >= a + 33) {
                // Region-bottom slab: the chunk, its four lateral neighbours and the layer above are solid, the chunk
                // below is absent -> exactly one Down face per column, in column order, and every AO probe (all at
                // y-1) reads "empty".  No tile, masks, scan or descriptors are needed (k_count_region counted 1024).
                const MatHeights mat{sm.hts, a};
                if constexpr (PACKED) {
                    uint32_t* const o32 = reinterpret_cast<uint32_t*>(out) + off;
                    for (uint32_t q = tid; q < (uint32_t)HB_CHUNK_AREA; q += kT) {
                        const uint32_t L = q << 5;
                        __stcs(o32 + q, L | (3u << 15) | (((mat(L, L >> 10, 0, (L >> 5) & 31) - 1u) & 15u) << 26));
                    }
                    continue;
                }
                if (tid == 32) sm.rng_seed = iy < kSeedLayers ? sm.layer_seed[iy] : siphash13_chunk(r.cx0 + ix, cy, r.cz0 + iz);
                __syncthreads();
                float* const ws = sm.stage + warp * (32 * 25);
                const uint32_t ws_shared = (uint32_t)__cvta_generic_to_shared(ws);
                char* const out_bytes = reinterpret_cast<char*>(out + off * 25ULL);
                for (uint32_t q0 = warp * 32u; q0 < (uint32_t)HB_CHUNK_AREA; q0 += (kT / 32) * 32u) {
                    const uint32_t L = (q0 + lane) << 5;  // voxel (x, 0, z) of column q0 + lane
                    emit_window(sm, mats, ws, ws_shared, out_bytes + q0 * 100u, 32u, L, 3u, 0u, mat(L, L >> 10, 0, (L >> 5) & 31),
                                (r.cx0 + ix) * 32, cy * 32, (r.cz0 + iz) * 32);
                }
                continue;
            }
            __syncthreads();  // previous chunk's readers of sm.occ are done
            const u64 ymask = 0x3FFFFFFFFULL & ~(iy == 0 ? 1ULL : 0ULL)         // chunk below outside the region: absent
                                             & ~(iy == r.ncy - 1 ? (1ULL << 33) : 0ULL);  // chunk above absent
            for (int i = tid; i < kTileN; i += kT) {
                const int d = min(max(sm.hts[i] - (a - 1), 0), 34);  // solid voxels at tile bits 0..33
                sm.occ[i] = ((1ULL << d) - 1ULL) & ymask;
            }
            if (EMIT && !PACKED && tid == 32) sm.rng_seed = iy < kSeedLayers ? sm.layer_seed[iy] : siphash13_chunk(r.cx0 + ix, cy, r.cz0 + iz);
            __syncthreads();
            const MatHeights mat{sm.hts, a};
            const uint32_t q = mesh_tile<EMIT, PACKED>(sm, mat, FacesFromOcc{sm.occ}, (r.cx0 + ix) * 32, cy * 32, (r.cz0 + iz) * 32,
                                                       mats, EMIT ? (PACKED ? out + off : out + off * 25ULL) : nullptr);
            if (!EMIT && tid == 0) counts[chunk] = q;
        }
        colw = s_col;  // written before this column's first barrier, read after several more
    }
    if (EMIT && !PACKED) bulk_wait_all();  // every bulk store issued by this thread has landed before the CTA retires
}


// ------------------------------------------------------------------------------------------------
// Single-pass fused region kernel: noise -> heights -> count -> offsets -> emit in ONE launch.
//
// The staged pipeline runs k_heights (FP32/ALU-issue bound), k_count_region (ALU/latency) and the emit kernel (HBM-write
// bound) one after the other, so the ALU-bound third of the step leaves HBM idle and the emit kernel leaves the ALUs
// idle.  Here every persistent CTA takes one chunk column at a time and does all of it: the 34x34 heights halo is
// evaluated from the noise function straight into shared memory (the one-column apron is recomputed: +13 % samples,
// no heights round trip through HBM), the per-layer face counts come from the closed form of k_count_region, and the
// column's place in the output comes from a chained scan over the columns in ticket order ("decoupled look-back":
// each column publishes its span, then its inclusive prefix, in one 64-bit status word).  With four CTAs per SM in
// different phases the SM's issue slots serve the noise of one column while the TMA engine streams another's quads.
//
// Offsets are the same exclusive scan over chunk index as launch_scan produces (tickets are handed out in column
// order and a column's chunks are consecutive), so the output is bit-identical to the staged pipeline's.
// Requires ncy <= 32 and the FBM noise kind; the launch wrapper falls back to the staged kernels otherwise.
// ------------------------------------------------------------------------------------------------
// 4 CTAs/SM like the staged emit kernel.  5 CTAs/SM (48 registers; descriptor buffer halved, the noise tables and
// count accumulators placed inside emit-only buffers to fit shared memory) measured slower: 4.32 vs 4.19 ms per
// config-4 step -- the kernel is bound by instruction issue (profiles/r2_ncu_region_fused_5cta_summary.csv: 3.58 G
// warp instructions, 73 % of the issue slots, top stall = barrier), not by residency.
constexpr int kFusedCtasPerSm = 4;
using FusedMesh = MeshSmemT<kDescCap>;
struct FusedSmem {
    FusedMesh sm;
    NoiseTables nt;
    uint32_t acc[kT / 32][32];  // per warp, per layer
    uint32_t cnt[32];           // the column's quad count per layer
    u64 off[32];                // and where each layer's quads go
    int next_col;
};
static_assert((sizeof(FusedSmem) + 1024) * kFusedCtasPerSm <= 228 * 1024, "fused kernel: shared memory per SM");

constexpr u64 kStatusAggregate = 1ULL << 62, kStatusPrefix = 2ULL << 62, kStatusValue = (1ULL << 62) - 1ULL;

__device__ __forceinline__ u64 ld_status(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

template <bool FAST>
__global__ void __launch_bounds__(kT, kFusedCtasPerSm)
k_region_fused(const MeshTables tb, const MaterialsDev* __restrict__ mats, const WorldDev w, const RegionDev r,
               int32_t* __restrict__ heights_out, uint32_t* __restrict__ counts, u64* __restrict__ offsets,
               float* __restrict__ out, u64 capacity, u64* __restrict__ total, u64* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char fused_smem_raw[];
    FusedSmem& fs = *reinterpret_cast<FusedSmem*>(fused_smem_raw);
    FusedMesh& sm = fs.sm;
    NoiseTables& nt = fs.nt;
    uint32_t (*const acc)[32] = fs.acc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_tables(sm, tb);
    load_noise_tables(nt, w.seed_lo);
    const RelHeight rel{(long long)r.cy0 * 32, 32 * r.ncy + 8};
    const int ncols = (r.ix_end - r.ix_begin) * r.ncz;
    const int last = r.ncy - 1;
    // every column, the first included, is drawn from the ticket counter (total[3]) at the moment its CTA starts on
    // it: a ticket is only ever held by a CTA that is already evaluating that column's noise, so the look-back below
    // waits at most one noise phase for a predecessor (drawing the next ticket early, while the current column is
    // still being emitted, makes every successor wait for that emit: measured 6x slower)
    for (;;) {
        __syncthreads();  // previous column's readers of fs.* are done
        if (tid == 0) fs.next_col = (int)atomicAdd(total + 3, 1ULL);
        __syncthreads();
        const int colw = fs.next_col;
        if (colw >= ncols) break;
        const int ix = r.ix_begin + colw / r.ncz, iz = colw % r.ncz;
        const int cx = r.cx0 + ix, cz = r.cz0 + iz;
        if (warp == 2 && lane < r.ncy) sm.layer_seed[lane] = siphash13_chunk(cx, r.cy0 + lane, cz);

        // ---- 1. noise: the 34 x 34 halo of heights, two samples per packed evaluation (element e = xi*34 + zi;
        //         pair = elements e and e + 578, i.e. tile rows xi and xi + 17)
        int cmax = INT32_MIN, hmin = INT32_MAX, emin = INT32_MAX;
        for (int p = tid; p < kTileN / 2; p += kT) {
            const int xa = p / kTile, zi = p - xa * kTile, xb = xa + kTile / 2;
            // world block coordinates are exact in f32 (|c| <= HB_MAX_CHUNK_COORD): the same values as the
            // reference's start + i of whichever chunk the sample belongs to (noise/f32.rs:87-112)
            const float2 x = pk(__int2float_rn(cx * 32 + xa - 1), __int2float_rn(cx * 32 + xb - 1));
            const float2 z = bc(__int2float_rn(cz * 32 + zi - 1));
            const float2 v = noise2_pair<HB_NOISE_FBM, FAST>(nt, smul2(x, bc(w.freq)), smul2(z, bc(w.freq)), w);
            const int raw[2] = {height_from_noise(v.x), height_from_noise(v.y)};
            const int dz = zi == 0 ? -1 : (zi == kTile - 1 ? 1 : 0);
            const bool z_in = (unsigned)(iz + dz) < (unsigned)r.ncz;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int xi = k ? xb : xa;
                const int dx = xi == 0 ? -1 : (xi == kTile - 1 ? 1 : 0);
                const bool present = z_in && (unsigned)(ix + dx) < (unsigned)r.ncx;  // columns outside the region are ABSENT
                const int h = rel(present ? raw[k] : kHeightAbsent);
                sm.hts[xi * kTile + zi] = h;
                const bool edge_x = dx != 0, edge_z = dz != 0;
                hmin = min(hmin, h);
                if (!(edge_x && edge_z)) emin = min(emin, h);
                if (!edge_x && !edge_z) {
                    cmax = max(cmax, h);
                    if (heights_out)
                        heights_out[((size_t)(ix - r.hx_begin) * r.ncz + iz) * HB_CHUNK_AREA + (xi - 1) * 32 + (zi - 1)] = raw[k];
                }
            }
        }
        cmax = __reduce_max_sync(0xffffffffu, cmax);
        hmin = __reduce_min_sync(0xffffffffu, hmin);
        emin = __reduce_min_sync(0xffffffffu, emin);
        if (lane == 0) { sm.red[0][warp] = cmax; sm.red[1][warp] = hmin; sm.red[2][warp] = emin; }
        acc[warp][lane] = 0;
        __syncthreads();
        cmax = __reduce_max_sync(0xffffffffu, lane < 8 ? sm.red[0][lane] : INT32_MIN);
        hmin = __reduce_min_sync(0xffffffffu, lane < 8 ? sm.red[1][lane] : INT32_MAX);
        emin = __reduce_min_sync(0xffffffffu, lane < 8 ? sm.red[2][lane] : INT32_MAX);

        // ---- 2. count: closed form per layer, exactly k_count_region's (same masks as the emit stage)
        for (int k = 0; k < 4; ++k) {
            const int* o = sm.hts + (k * 8 + warp + 1) * kTile + (lane + 1);
            const int hc = o[0], hE = o[kTile], hW = o[-kTile], hN = o[1], hS = o[-1];
            int lo = min(min(hc - 1, hE), min(min(hW, hN), hS));
            lo = hc > 0 ? (lo < 0 ? 0 : lo >> 5) : INT32_MAX;
            const int hi = hc > 0 ? (hc - 1) >> 5 : -1;
            const int wlo = __reduce_min_sync(0xffffffffu, lo);
            const int whi = __reduce_max_sync(0xffffffffu, hi);
            const int nsolid = __popc(__ballot_sync(0xffffffffu, hc > 0));
            const int j1 = min(whi, last);
            for (int j = max(wlo, 0); j <= j1; ++j) {  // warp-uniform bounds
                const int a = j * 32;
                const int rc = min(max(hc - a, 0), 32);
                uint32_t v = 0;
                if (rc > 0) {
                    v = max(0, rc - min(max(hE - a, 0), 32)) + max(0, rc - min(max(hW - a, 0), 32)) +
                        max(0, rc - min(max(hN - a, 0), 32)) + max(0, rc - min(max(hS - a, 0), 32));
                    v += (hc - a <= 32 || j == last) ? 1u : 0u;
                    v += (j == 0) ? 1u : 0u;
                }
                v = __reduce_add_sync(0xffffffffu, v);
                if (lane == 0) acc[warp][j] += v;
            }
            if (lane == 0 && nsolid) {
                if (wlo > 0) acc[warp][0] += nsolid * (last == 0 ? 2 : 1);
                if (last > 0 && last < wlo) acc[warp][last] += nsolid;
            }
        }
        __syncthreads();

        // ---- 3. offsets: warp 0 scans the column's layers and chains the column into the global order
        if (warp == 0) {
            uint32_t c = 0;
            if (lane < r.ncy) {
#pragma unroll
                for (int wq = 0; wq < kT / 32; ++wq) c += acc[wq][lane];
            }
            const u64 span = ((u64)c + 3ULL) & ~3ULL;
            u64 incl = span;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const u64 col_span = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t col_quads = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0) st_status(status + colw, kStatusAggregate | col_span);
            // look back over the preceding columns, 32 at a time: sum aggregates down to the nearest inclusive prefix
            u64 excl = 0;
            for (int j = colw - 1; j >= 0;) {
                const int idx = j - lane;
                u64 v;
                do {
                    v = idx >= 0 ? ld_status(status + idx) : kStatusPrefix;  // before column 0: prefix 0
                } while (__any_sync(0xffffffffu, (v >> 62) == 0ULL));
                const uint32_t is_prefix = __ballot_sync(0xffffffffu, (v >> 62) == 2ULL);
                const int stop = is_prefix ? __ffs(is_prefix) - 1 : 32;  // nearest prefix (lowest lane = closest column)
                u64 part = lane <= stop ? (v & kStatusValue) : 0ULL;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                excl += part;
                if (is_prefix) break;
                j -= 32;
            }
            if (lane == 0) {
                st_status(status + colw, kStatusPrefix | (excl + col_span));
                if (col_quads) atomicAdd(total + 1, (u64)col_quads);
                if (colw == ncols - 1) total[0] = excl + col_span;
            }
            if (lane < r.ncy) {
                const u64 o = excl + incl - span;
                fs.cnt[lane] = c;
                fs.off[lane] = o;
                counts[(size_t)colw * r.ncy + lane] = c;
                offsets[(size_t)colw * r.ncy + lane] = o;
            }
        }
        __syncthreads();

        // ---- 4. emit, layer by layer (the body of k_mesh_region<true>)
        for (int iy = 0; iy < r.ncy; ++iy) {
            const uint32_t my_count = fs.cnt[iy];
            if (my_count == 0) continue;  // block-uniform: all air, buried, or no visible face
            const u64 off = fs.off[iy];
            if (off + (((u64)my_count + 3ULL) & ~3ULL) > capacity) {
                if (tid == 0) atomicOr(total + 2, 1ULL);  // error path only
                continue;
            }
            const int cy = r.cy0 + iy;
            const int a = iy * 32;
            const MatHeights mat{sm.hts, a};
            if (iy == 0 && r.ncy > 1 && emin >= a + 33) {  // region-bottom slab: one Down face per column
                if (tid == 32) sm.rng_seed = sm.layer_seed[iy];
                __syncthreads();
                float* const ws = sm.stage + warp * (32 * 25);
                const uint32_t ws_shared = (uint32_t)__cvta_generic_to_shared(ws);
                char* const out_bytes = reinterpret_cast<char*>(out + off * 25ULL);
                for (uint32_t q0 = warp * 32u; q0 < (uint32_t)HB_CHUNK_AREA; q0 += (kT / 32) * 32u) {
                    const uint32_t L = (q0 + lane) << 5;
                    emit_window(sm, mats, ws, ws_shared, out_bytes + q0 * 100u, 32u, L, 3u, 0u, mat(L, L >> 10, 0, (L >> 5) & 31),
                                cx * 32, cy * 32, cz * 32);
                }
                continue;
            }
            __syncthreads();  // previous layer's readers of sm.occ / sm.rng_seed are done
            const u64 ymask = 0x3FFFFFFFFULL & ~(iy == 0 ? 1ULL : 0ULL) & ~(iy == r.ncy - 1 ? (1ULL << 33) : 0ULL);
            for (int i = tid; i < kTileN; i += kT) {
                const int d = min(max(sm.hts[i] - (a - 1), 0), 34);
                sm.occ[i] = ((1ULL << d) - 1ULL) & ymask;
            }
            if (tid == 32) sm.rng_seed = sm.layer_seed[iy];
            __syncthreads();
            mesh_tile<true>(sm, mat, FacesFromOcc{sm.occ}, cx * 32, cy * 32, cz * 32, mats, out + off * 25ULL);
        }
    }
    bulk_wait_all();  // every bulk store issued by this thread has landed before the CTA retires
}

// ------------------------------------------------------------------------------------------------
// counts -> offsets.  Each chunk's span is its count rounded up to 4 instances.  Three small
// kernels (tile sums, scan of tile sums, offsets); no atomics, deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kScanT = 256;
constexpr int kScanPer = 8;
constexpr int kScanTile = kScanT * kScanPer;

__device__ __forceinline__ u64 block_sum_u64(u64 v, u64* s_tmp) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    u64 t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_tmp[i];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kScanT) k_tile_sums(const uint32_t* __restrict__ counts, size_t n,
                                                       u64* __restrict__ tile_span, u64* __restrict__ tile_quads) {
    __shared__ u64 s_tmp[8];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPer;
    u64 span = 0, quads = 0;
#pragma unroll
    for (int e = 0; e < kScanPer; ++e) {
        if (base + e < n) {
            const uint32_t c = counts[base + e];
            quads += c;
            span += ((u64)c + 3ULL) & ~3ULL;
        }
    }
    span = block_sum_u64(span, s_tmp);
    quads = block_sum_u64(quads, s_tmp);
    if (threadIdx.x == 0) { tile_span[blockIdx.x] = span; tile_quads[blockIdx.x] = quads; }
}

// single CTA: exclusive scan of tile spans in place, starting at total[0]; total[0] (span) and total[1]
// (quads) are advanced, so consecutive slabs of one plan chain their offsets
__global__ void __launch_bounds__(1024) k_scan_sums(u64* __restrict__ tile_span, const u64* __restrict__ tile_quads,
                                                     size_t ntiles, u64* __restrict__ total) {
    __shared__ u64 s_w[32];
    __shared__ u64 s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = total[0];  // running span of the slabs scanned before this one (0 after a reset)
    u64 quads = 0;
    __syncthreads();
    for (size_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const size_t i = b0 + tid;
        const u64 v = i < ntiles ? tile_span[i] : 0;
        if (i < ntiles) quads += tile_quads[i];
        u64 s = v;
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) s_w[warp] = s;
        __syncthreads();
        u64 wbase = 0;
        for (int k = 0; k < warp; ++k) wbase += s_w[k];
        const u64 carry = s_carry;
        if (i < ntiles) tile_span[i] = carry + wbase + s - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + wbase + s;
        __syncthreads();
    }
    quads = block_sum_u64(quads, s_w);
    if (tid == 0) { total[0] = s_carry; total[1] += quads; }
}

__global__ void __launch_bounds__(kScanT) k_offsets(const uint32_t* __restrict__ counts, size_t n,
                                                     const u64* __restrict__ tile_base, u64* __restrict__ offsets) {
    __shared__ u64 s_w[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)tid * kScanPer;
    u64 sp[kScanPer];
    u64 mine = 0;
#pragma unroll
    for (int e = 0; e < kScanPer; ++e) {
        sp[e] = (base + e < n) ? (((u64)counts[base + e] + 3ULL) & ~3ULL) : 0ULL;
        mine += sp[e];
    }
    u64 s = mine;
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
    }
    if (lane == 31) s_w[warp] = s;
    __syncthreads();
    u64 run = tile_base[blockIdx.x] + s - mine;
    for (int k = 0; k < warp; ++k) run += s_w[k];
#pragma unroll
    for (int e = 0; e < kScanPer; ++e) {
        if (base + e < n) offsets[base + e] = run;
        run += sp[e];
    }
}

// Small batches (<= 16 384 chunks: a streamed slab, a world remesh, BASELINE's 4 096-chunk configs): the whole scan in
// one CTA -- per-thread runs of up to 16 consecutive spans, one block scan, offsets, totals -- instead of three
// launches of 5-6 us each.  `reset`: start from zero totals (what the separate memset did); else chain onto total[0].
constexpr int kSmallScanT = 1024, kSmallScanPer = 16;
__global__ void __launch_bounds__(kSmallScanT) k_scan_small(const uint32_t* __restrict__ counts, uint32_t n, u64* __restrict__ offsets,
                                                             u64* __restrict__ total, int reset) {
    __shared__ u64 s_w[32];
    __shared__ u64 s_q[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n + kSmallScanT - 1) / kSmallScanT;  // <= kSmallScanPer
    const uint32_t base = (uint32_t)tid * per;
    u64 sp[kSmallScanPer];
    u64 mine = 0, quads = 0;
#pragma unroll
    for (int e = 0; e < kSmallScanPer; ++e) {
        sp[e] = 0;
        if ((uint32_t)e < per && base + e < n) {
            const uint32_t c = counts[base + e];
            quads += c;
            sp[e] = ((u64)c + 3ULL) & ~3ULL;
            mine += sp[e];
        }
    }
    u64 s = mine;
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
    }
    for (int d = 16; d > 0; d >>= 1) quads += __shfl_xor_sync(0xffffffffu, quads, d);
    if (lane == 31) s_w[warp] = s;
    if (lane == 0) s_q[warp] = quads;
    __syncthreads();
    const u64 start = reset ? 0ULL : total[0];
    const u64 q0 = reset ? 0ULL : total[1];
    u64 run = start + s - mine, all = 0, allq = 0;
    for (int k = 0; k < 32; ++k) {
        if (k < warp) run += s_w[k];
        all += s_w[k];
        allq += s_q[k];
    }
#pragma unroll
    for (int e = 0; e < kSmallScanPer; ++e) {
        if ((uint32_t)e < per && base + e < n) offsets[base + e] = run;
        run += sp[e];
    }
    __syncthreads();  // every thread has read total[] before it is rewritten
    if (tid == 0) { total[0] = start + all; total[1] = q0 + allq; }
}

int g_num_sms = 0;

}  // namespace

int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        g_num_sms = v;
    }
    return g_num_sms;
}

cudaError_t launch_pack(cudaStream_t s, const uint16_t* d_ids, size_t n, int n_materials, uint32_t* d_bitmaps,
                        uint32_t* d_summary, uint32_t* d_borders, uint32_t* d_counts, uint64_t* d_total) {
    if (n == 0) return cudaSuccess;
    size_t grid = (size_t)num_sms() * 32;
    if (grid > n) grid = n;
    k_pack<<<(unsigned)grid, kT, 0, s>>>(d_ids, n, n_materials, d_bitmaps, d_summary, d_borders, d_counts, (u64*)d_total);
    return cudaGetLastError();
}

cudaError_t launch_pack_cubes(cudaStream_t s, const hb_palette_cube* d_cubes, size_t n, int n_materials,
                              uint32_t* d_bitmaps, uint32_t* d_summary, uint32_t* d_planes, uint64_t* d_total,
                              const int32_t* d_list) {
    if (n == 0) return cudaSuccess;
    size_t grid = (size_t)num_sms() * 8;
    if (grid > n) grid = n;
    k_pack_cubes<<<(unsigned)grid, kT, 0, s>>>(d_cubes, n, n_materials, d_bitmaps, d_summary, d_planes, (u64*)d_total, d_list);
    return cudaGetLastError();
}

cudaError_t launch_mesh_ids(cudaStream_t s, bool emit, const MeshTables& tb, const MaterialsDev* d_mats,
                            const hb_chunk_pt* d_pts, size_t n, const void* d_blocks, const int32_t* d_nbr,
                            const uint32_t* d_bitmaps, const uint32_t* d_summary, const uint32_t* d_planes,
                            uint32_t* d_counts, const uint64_t* d_offsets, hb_instance3d* d_out, uint64_t capacity,
                            uint64_t* d_total, const int32_t* d_list, const uint32_t* d_borders) {
    if (n == 0) return cudaSuccess;
    // Persistent CTAs.  EMIT draws work from a ticket counter (see k_mesh_ids); COUNT strides statically: batches are
    // usually ordered (ix, iz, iy) with a small number of layers, so an even stride would hand every CTA the same
    // layer (all surface or all buried) for the whole launch; an odd stride rotates through the layers.
    size_t grid = ((size_t)num_sms() * kCtasPerSm) - 1;
    if (grid > n) grid = n;
    // ticket batches: one chunk for small launches (best balance), up to 8 when every CTA still gets >= 16 batches
    const uint32_t batch = (uint32_t)std::min<size_t>(8, std::max<size_t>(1, n / (grid * 16)));
#define HB_LAUNCH(E, F)                                                                                              \
    k_mesh_ids<E, F><<<(unsigned)grid, kT, 0, s>>>(tb, d_mats, d_pts, n, d_blocks, d_nbr, d_bitmaps, d_summary, d_planes, \
                                                   d_counts, (const u64*)d_offsets, (float*)d_out, capacity, (u64*)d_total, d_list, batch)
    if (!emit) {
        size_t cgrid = (size_t)num_sms() * 8 - 1;  // odd, see above
        if (cgrid > n) cgrid = n;
        if (d_planes) {
            k_count_ids<<<(unsigned)cgrid, kT, 0, s>>>(n, d_summary, d_planes, d_counts, d_list);
        } else {  // ids path: k_pack left the chunk-local faces in d_counts, the hull is added from the border records
            if (!d_borders || d_list) return cudaErrorInvalidValue;
            uint32_t* d_nz = const_cast<uint32_t*>(d_borders) + n * kBorderWords;  // [0] = length, [1..] = chunks with quads
            const cudaError_t e = cudaMemsetAsync(d_nz, 0, sizeof(uint32_t), s);
            if (e != cudaSuccess) return e;
            k_count_border<<<(unsigned)((n + kT / 32 - 1) / (kT / 32)), kT, 0, s>>>(n, d_nbr, d_borders, d_counts, d_nz);
        }
    } else {
        const cudaError_t e = cudaMemsetAsync(d_total + 3, 0, sizeof(uint64_t), s);  // the ticket counter
        if (e != cudaSuccess) return e;
        if (d_planes) {
            HB_LAUNCH(true, true);
        } else {
            if (!d_borders || d_list) return cudaErrorInvalidValue;
            const size_t egrid = (size_t)num_sms() * kCtasPerSm;  // k_emit_ids exits at once when its first ticket is past the list
            static bool attr_set[16] = {};
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev >= 0 && dev < 16 && !attr_set[dev]) {
                const cudaError_t ea = cudaFuncSetAttribute(k_emit_ids, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitIdsSmem));
                if (ea != cudaSuccess) return ea;
                attr_set[dev] = true;
            }
            k_emit_ids<<<(unsigned)egrid, kT, sizeof(EmitIdsSmem), s>>>(tb, d_mats, d_pts, static_cast<const uint16_t*>(d_blocks), d_nbr, d_bitmaps, d_borders,
                                                      d_borders + n * kBorderWords, d_counts, (const u64*)d_offsets, (float*)d_out,
                                                      capacity, (u64*)d_total);
        }
    }
#undef HB_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_mesh_region(cudaStream_t s, bool emit, const MeshTables& tb, const MaterialsDev* d_mats,
                               const RegionDev& r, const int32_t* d_heights, uint32_t* d_counts,
                               const uint64_t* d_offsets, void* d_out, uint64_t capacity,
                               uint64_t* d_total, int ctas_per_sm, bool packed) {
    const size_t ncols = (size_t)(r.ix_end - r.ix_begin) * r.ncz;
    if (ncols == 0 || r.ncy == 0) return cudaSuccess;
    if (ctas_per_sm < 1 || ctas_per_sm > kCtasPerSm) ctas_per_sm = kCtasPerSm;
    size_t grid = (size_t)num_sms() * ctas_per_sm;  // persistent CTAs
    if (grid > ncols) grid = ncols;
    if (emit) {
        const cudaError_t e = cudaMemsetAsync(d_total + 3, 0, sizeof(uint64_t), s);  // the column ticket counter
        if (e != cudaSuccess) return e;
        if (packed)
            k_mesh_region<true, true><<<(unsigned)grid, kT, 0, s>>>(tb, d_mats, r, d_heights, d_counts, (const u64*)d_offsets,
                                                                    (float*)d_out, capacity, (u64*)d_total);
        else
            k_mesh_region<true><<<(unsigned)grid, kT, 0, s>>>(tb, d_mats, r, d_heights, d_counts, (const u64*)d_offsets,
                                                              (float*)d_out, capacity, (u64*)d_total);
    }
    else {
        const cudaError_t e = cudaMemsetAsync(d_total + 3, 0, sizeof(uint64_t), s);  // the column ticket counter
        if (e != cudaSuccess) return e;
        k_count_region<<<(unsigned)std::min<size_t>(ncols, (size_t)num_sms() * 8), kT, 0, s>>>(r, d_heights, d_counts, (u64*)d_total + 3);
    }
    return cudaGetLastError();
}

bool region_fused_supported(const WorldDev& w, const RegionDev& r) { return w.kind == HB_NOISE_FBM && w.fma && r.ncy <= 32; }

size_t region_fused_status_bytes(const RegionDev& r) { return (size_t)(r.ix_end - r.ix_begin) * r.ncz * sizeof(u64); }

cudaError_t launch_region_fused(cudaStream_t s, const MeshTables& tb, const MaterialsDev* d_mats, const WorldDev& w,
                                const RegionDev& r, int32_t* d_heights, uint32_t* d_counts, uint64_t* d_offsets,
                                hb_instance3d* d_out, uint64_t capacity, uint64_t* d_total, uint64_t* d_status, bool fast,
                                int ctas_per_sm) {
    const size_t ncols = (size_t)(r.ix_end - r.ix_begin) * r.ncz;
    if (ncols == 0 || r.ncy == 0) return cudaSuccess;
    static bool attr_set[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_region_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_region_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    cudaError_t e = cudaMemsetAsync(d_total, 0, 4 * sizeof(uint64_t), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_status, 0, ncols * sizeof(u64), s);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm < 1 || ctas_per_sm > kFusedCtasPerSm) ctas_per_sm = kFusedCtasPerSm;
    size_t grid = (size_t)num_sms() * ctas_per_sm;
    if (grid > ncols) grid = ncols;
    if (fast)
        k_region_fused<true><<<(unsigned)grid, kT, sizeof(FusedSmem), s>>>(tb, d_mats, w, r, d_heights, d_counts, (u64*)d_offsets,
                                                                           (float*)d_out, capacity, (u64*)d_total, (u64*)d_status);
    else
        k_region_fused<false><<<(unsigned)grid, kT, sizeof(FusedSmem), s>>>(tb, d_mats, w, r, d_heights, d_counts, (u64*)d_offsets,
                                                                            (float*)d_out, capacity, (u64*)d_total, (u64*)d_status);
    return cudaGetLastError();
}

int scan_launch_count(size_t n) { return n == 0 ? 0 : (n <= (size_t)kSmallScanT * kSmallScanPer ? 1 : 3); }

size_t scan_workspace_bytes(size_t n) {
    const size_t ntiles = (n + kScanTile - 1) / kScanTile;
    return 2 * (ntiles + 1) * sizeof(u64);
}

cudaError_t launch_scan(cudaStream_t s, const uint32_t* d_counts, size_t n, uint64_t* d_offsets, void* d_ws,
                        uint64_t* d_total, bool accumulate) {
    if (n != 0 && n <= (size_t)kSmallScanT * kSmallScanPer) {
        k_scan_small<<<1, kSmallScanT, 0, s>>>(d_counts, (uint32_t)n, (u64*)d_offsets, (u64*)d_total, accumulate ? 0 : 1);
        return cudaGetLastError();
    }
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(d_total, 0, 2 * sizeof(u64), s);
        if (e != cudaSuccess) return e;
    }
    if (n == 0) return cudaSuccess;
    const size_t ntiles = (n + kScanTile - 1) / kScanTile;
    u64* tile_span = (u64*)d_ws;
    u64* tile_quads = tile_span + ntiles + 1;
    k_tile_sums<<<(unsigned)ntiles, kScanT, 0, s>>>(d_counts, n, tile_span, tile_quads);
    k_scan_sums<<<1, 1024, 0, s>>>(tile_span, tile_quads, ntiles, (u64*)d_total);
    k_offsets<<<(unsigned)ntiles, kScanT, 0, s>>>(d_counts, n, tile_span, (u64*)d_offsets);
    return cudaGetLastError();
}

}  // namespace hb
