// host_common.cuh -- host-side definitions shared by api.cu and world.cu: error plumbing, scratch buffers,
// chunk-coordinate hash keys and the context object behind the opaque hb_ctx handle.
#pragma once
#include <stdarg.h>
#include <stdint.h>

#include <atomic>
#include <functional>
#include <mutex>

#include "common.cuh"

namespace hb {

// records the thread-local message hb_last_error() returns, and hands back `code` (api.cu)
int fail(int code, const char* fmt, ...);

#define HB_CUDA(expr)                                                                                 \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(HB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// Growable device / pinned-host scratch buffer.
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = pinned ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; cap = 0; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
        p = nullptr; cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};


struct KeyXZ {
    int32_t x, z;
    bool operator==(const KeyXZ& o) const { return x == o.x && z == o.z; }
};
struct KeyXZHash {
    size_t operator()(const KeyXZ& k) const {
        return (size_t)(((uint64_t)(uint32_t)k.x << 32) | (uint32_t)k.z) * 0x9E3779B97F4A7C15ULL >> 7;
    }
};
struct KeyXYZ {
    int32_t x, y, z;
    bool operator==(const KeyXYZ& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct KeyXYZHash {
    size_t operator()(const KeyXYZ& k) const {
        uint64_t h = (uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ULL;
        h ^= ((uint64_t)(uint32_t)k.y + 0x7F4A7C15ULL) * 0xC2B2AE3D27D4EB4FULL;
        h = (h << 31) | (h >> 33);
        h ^= (uint64_t)(uint32_t)k.z * 0x165667B19E3779F9ULL;
        return (size_t)(h ^ (h >> 29));
    }
};

inline bool coord_ok(int32_t c) { return c >= -HB_MAX_CHUNK_COORD && c <= HB_MAX_CHUNK_COORD; }

// ---- packed-quad transport, host half (expand.cu) ----
// Everything an Instance3d holds besides the packed word and the chunk position.
struct ExpandTables {
    alignas(16) float face_mat[6][12];
    alignas(16) float ao4[256][4];                              // ao byte -> Instance3d::ao
    alignas(16) float rgba[HB_MAX_MATERIALS][HB_MAX_COLORS][4];
    float color_scale[HB_MAX_MATERIALS];                        // (len - 1) as f32, material.rs:69
    void build(const MeshTables& tb, const MaterialsDev& m);
};
uint64_t chunk_color_seed(int32_t x, int32_t y, int32_t z);
void expand_chunk(const ExpandTables& T, hb_chunk_pt pt, const uint32_t* packed, uint32_t n, hb_instance3d* out);
// heights tile of the chunk's column (int32[1024], index x*32 + z) -> the chunk's 32 768 block ids (out 16-byte aligned)
void fill_chunk_ids(const int32_t* heights, int32_t cy, uint16_t* out);

// Persistent host worker pool (the caller participates): run() returns when fn(i) ran for every i in [0, n_items).
class HostPool {
   public:
    HostPool();
    ~HostPool();
    void resize(int n_threads);  // total threads incl. the caller; <= 1 = run inline
    int threads() const;
    void run(size_t n_items, size_t grain, const std::function<void(size_t)>& fn);
    // The same on the workers alone: start() returns at once (with a single thread it runs the items inline), wait()
    // returns when they are done.  Nothing else may use the pool between the two.
    void start(size_t n_items, size_t grain, const std::function<void(size_t)>& fn);
    void wait();

   private:
    struct Impl;
    Impl* impl_;
};
int default_host_threads();  // min(hardware threads, 32), or $HB_HOST_THREADS

}  // namespace hb

struct hb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;  // compute
    cudaStream_t copy = nullptr;    // D2H of finished slabs
    std::mutex mu;
    hb::WorldDev world{};
    int64_t seed = 0;
    hb::MeshTables tables{};
    hb::MaterialsDev mats{};
    hb::MaterialsDev* d_mats = nullptr;
    // scratch for the host-pointer entry points
    hb::Buf d_pts, d_cols, d_colidx, d_heights, d_noise, d_ids, d_cubes, d_nbr, d_bitmaps, d_summary, d_counts,
        d_offsets, d_scan, d_total, d_out;
    hb::Buf h_total;
    // packed-quad transport: expander tables (rebuilt when the materials change), worker pool, the expanded slab
    hb::ExpandTables xt;
    std::atomic<bool> xt_valid{false};
    hb::HostPool pool;
    int host_threads = 0;           // 0 = not started yet (default_host_threads() on first use)
    int transport = 0;              // HB_TRANSPORT_*
    void* h_expanded = nullptr;     // plain (pageable, huge-page advised) host memory: the sink's Instance3d view
    size_t h_expanded_cap = 0;
    void* h_ids_plain = nullptr;    // same kind of memory for the block ids rebuilt from heights (want_ids, packed transports)
    size_t h_ids_plain_cap = 0;
    hb::Buf h_heights[2];           // pinned landing buffers of the slab's heights tiles
    hb::Buf h_packed[2];            // pinned landing buffers of the packed words
    hb::Buf h_bounce[2];            // pinned bounce buffers of copy_to_host (pageable destinations)
    cudaEvent_t ev_bounce[2] = {nullptr, nullptr};
    // region streaming: pinned staging, and the two double-buffered plans + events kept across calls so that a
    // steady-state hb_generate_mesh_region call allocates and frees nothing
    hb::Buf h_inst[2], h_off[2], h_cnt[2], h_ids[2];
    struct hb_region_plan* stream_plan[2] = {nullptr, nullptr};
    cudaEvent_t ev_kernels[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    hb_ctx() {
        h_total.pinned = true;
        h_bounce[0].pinned = h_bounce[1].pinned = true;
        h_packed[0].pinned = h_packed[1].pinned = true;
        h_heights[0].pinned = h_heights[1].pinned = true;
        for (int i = 0; i < 2; ++i) h_inst[i].pinned = h_off[i].pinned = h_cnt[i].pinned = h_ids[i].pinned = true;
    }
};

namespace hb {
// Blocking device->host copy on stream `s` (c->mu held by the caller).  A pinned destination (hb_host_alloc,
// cudaHostAlloc, cudaHostRegister) is one DMA at link speed; a pageable one is streamed through two pinned bounce
// buffers so that the DMA of piece k+1 overlaps the memcpy of piece k (instead of the driver's staged path).
int copy_to_host(hb_ctx* c, cudaStream_t s, void* dst, const void* d_src, size_t bytes);
}  // namespace hb
