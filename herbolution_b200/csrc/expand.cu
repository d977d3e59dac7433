// expand.cu -- host half of the packed-quad transport (include/herbgpu.h: hb_packed_quad, hb_expand_quads).
//
// The mesher's result is 100 bytes per visible face (Instance3d, client/src/video/world/vertex.rs:41-68) but only
// 30 bits of it are information: voxel index (15), face (3), the four vertex_ao values (8) and the material (4).
// The rest is a pure function of those bits, the chunk position and constant tables:
//   model_0..2 = the face's rotation matrix        (lib/src/spatial/face.rs:60-69 -> 6 constant matrices)
//   model_3    = (chunk * 32 + voxel, 1) as f32    (client/src/world/chunk.rs:201-204)
//   rgba       = material.colors[(len-1) * p],  p = draw 6*L+face of fastrand::Rng::with_seed(hash(ChunkPt))
//                                                  (chunk.rs:182-199, server/src/chunk/material.rs:67-71)
//   light      = 0, ao = lut[vertex_ao]            (chunk.rs:207,255-276)
// So the device->host link carries 4 bytes per quad and the host threads below write the 100-byte records with
// non-temporal stores -- the link moves 25x less and the records are produced at host memory-write speed.
// This is product code (the drop-in returns the reference's Instance3d); it shares no code with oracle/.
#include <emmintrin.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>
#include <vector>

#include "host_common.cuh"

namespace hb {

namespace {

inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

}  // namespace

// std::hash::DefaultHasher (SipHash-1-3, keys 0,0) over the 12 bytes of ChunkPt -- the seed of the chunk's colour
// RNG (client/src/world/chunk.rs:182-184).  Same function as siphash13_chunk in mesh.cu.
uint64_t chunk_color_seed(int32_t x, int32_t y, int32_t z) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL, v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    auto round = [&]() {
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
    };
    const uint64_t m0 = (uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32);
    v3 ^= m0; round(); v0 ^= m0;
    const uint64_t b = ((uint64_t)12 << 56) | (uint64_t)(uint32_t)z;
    v3 ^= b; round(); v0 ^= b;
    v2 ^= 0xff;
    round(); round(); round();
    return v0 ^ v1 ^ v2 ^ v3;
}

void ExpandTables::build(const MeshTables& tb, const MaterialsDev& m) {
    memcpy(face_mat, tb.face_mat, sizeof(face_mat));
    for (int a = 0; a < 256; ++a)
        for (int v = 0; v < 4; ++v) ao4[a][v] = tb.ao_lut[(a >> (2 * v)) & 3];
    for (int i = 0; i < HB_MAX_MATERIALS; ++i) {
        const int nc = m.n_colors[i];
        color_scale[i] = (float)(nc > 0 ? nc - 1 : 0);
        memcpy(rgba[i], m.rgba[i], sizeof(rgba[i]));
    }
}

// One chunk: n packed quads -> n Instance3d at `out` (16-byte aligned), zero-filled up to the next multiple of 4.
// Four records (400 B = 25 x 16 B) are assembled in a stack buffer and leave as 25 non-temporal 16-byte stores.
void expand_chunk(const ExpandTables& T, hb_chunk_pt pt, const uint32_t* __restrict__ packed, uint32_t n, hb_instance3d* out) {
    const uint64_t seed = chunk_color_seed(pt.x, pt.y, pt.z);
    const __m128i base = _mm_set_epi32(1, pt.z * 32, pt.y * 32, pt.x * 32);
    alignas(64) float tmp[100];
    char* dst = reinterpret_cast<char*>(out);
    for (uint32_t q0 = 0; q0 < n; q0 += 4) {
        const uint32_t m = n - q0 < 4 ? n - q0 : 4;
        if (m < 4) memset(tmp, 0, sizeof(tmp));
        for (uint32_t k = 0; k < m; ++k) {
            const uint32_t w = packed[q0 + k];
            const uint32_t L = w & 0x7FFFu, f = (w >> 15) & 7u, ao = (w >> 18) & 255u, mi = (w >> 26) & 15u;
            float* r = tmp + k * 25;
            const float* fm = T.face_mat[f < 6 ? f : 0];
            _mm_storeu_ps(r, _mm_loadu_ps(fm));
            _mm_storeu_ps(r + 4, _mm_loadu_ps(fm + 4));
            _mm_storeu_ps(r + 8, _mm_loadu_ps(fm + 8));
            // (chunk_position + position).cast::<f32>(), w = 1
            const __m128i loc = _mm_set_epi32(0, (int)((L >> 5) & 31u), (int)(L & 31u), (int)(L >> 10));
            _mm_storeu_ps(r + 12, _mm_cvtepi32_ps(_mm_add_epi32(base, loc)));
            // fastrand 2.3.0 wyrand, draw number 6*L + f (counter form; same arithmetic as wyrand_f32_at in mesh.cu)
            const uint64_t s = seed + (uint64_t)(6u * L + f + 1u) * 0x2d358dccaa6c78a5ULL;
            const unsigned __int128 t = (unsigned __int128)s * (s ^ 0x8bb84b93962eacc9ULL);
            const uint32_t bits = 0x3F800000u | ((uint32_t)((uint64_t)t ^ (uint64_t)(t >> 64)) >> 9);
            float p;
            memcpy(&p, &bits, 4);
            p -= 1.0f;
            const int ci = (int)(T.color_scale[mi] * p);  // Material::get_color: colors[((len - 1) as f32 * p) as usize]
            _mm_storeu_ps(r + 16, _mm_load_ps(T.rgba[mi][ci]));
            r[20] = 0.0f;  // light = 0 (bit pattern 0)
            _mm_storeu_ps(r + 21, _mm_load_ps(T.ao4[ao]));
        }
        for (int i = 0; i < 25; ++i) _mm_stream_ps(reinterpret_cast<float*>(dst) + i * 4, _mm_load_ps(tmp + i * 4));
        dst += 400;
    }
    _mm_sfence();  // non-temporal stores are weakly ordered: make them visible before the job is reported done
}

// Host half of the block-id transport (want_ids): the generator's block ids are a function of the column's heights
// alone (GenerationParams::generate, server/src/generator/mod.rs:85-91: with d = h - y, stone for d > 6, dirt for d > 1,
// grass for d > 0, else air), and a chunk column shares one 4 KiB heights tile between all its layers.  So the heights
// cross the link (4 KiB per column instead of 64 KiB per chunk) and the ids are written here, one 64-byte voxel column
// per table row: row d = clamp(h - 32*cy, 0, 38) holds the 32 ids of a column whose surface is d voxels above the chunk's
// floor (d >= 38: all stone).  Same ids as k_fill writes on the device (tests compare both transports).
void fill_chunk_ids(const int32_t* __restrict__ heights, int32_t cy, uint16_t* __restrict__ out) {
    struct Lut {
        alignas(64) uint16_t row[39][32];
        Lut() {
            for (int d = 0; d < 39; ++d)
                for (int y = 0; y < 32; ++y) {
                    const int k = d - y;
                    row[d][y] = (uint16_t)(k > 6 ? 1 : (k > 1 ? 2 : (k > 0 ? 3 : 0)));
                }
        }
    };
    static const Lut lut;
    const long long wy0 = (long long)cy * 32;
    for (int c = 0; c < HB_CHUNK_AREA; ++c) {
        const long long dl = (long long)heights[c] - wy0;
        const int d = (int)(dl < 0 ? 0 : (dl > 38 ? 38 : dl));
        const __m128i* src = reinterpret_cast<const __m128i*>(lut.row[d]);
        __m128i* dst = reinterpret_cast<__m128i*>(out + c * 32);
        _mm_stream_si128(dst, _mm_load_si128(src));
        _mm_stream_si128(dst + 1, _mm_load_si128(src + 1));
        _mm_stream_si128(dst + 2, _mm_load_si128(src + 2));
        _mm_stream_si128(dst + 3, _mm_load_si128(src + 3));
    }
    _mm_sfence();
}

// ---- a small persistent worker pool: run(n_items, fn) calls fn(item) for every item on all workers + the caller ----
struct HostPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::function<void(size_t)> fn;
    size_t n_items = 0, grain = 1;
    std::atomic<size_t> next{0};
    uint64_t generation = 0;
    int active = 0;
    bool stop = false;

    void drain() {
        for (;;) {
            const size_t i0 = next.fetch_add(grain, std::memory_order_relaxed);
            if (i0 >= n_items) return;
            const size_t i1 = i0 + grain < n_items ? i0 + grain : n_items;
            for (size_t i = i0; i < i1; ++i) fn(i);
        }
    }
    void worker() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            lk.unlock();
            drain();
            lk.lock();
            if (--active == 0) cv_done.notify_all();
        }
    }
};

HostPool::HostPool() : impl_(nullptr) {}
HostPool::~HostPool() { resize(0); }

int HostPool::threads() const { return impl_ ? (int)impl_->workers.size() + 1 : 1; }

void HostPool::resize(int n_threads) {
    if (impl_) {
        {
            std::lock_guard<std::mutex> lk(impl_->mu);
            impl_->stop = true;
        }
        impl_->cv_work.notify_all();
        for (auto& t : impl_->workers) t.join();
        delete impl_;
        impl_ = nullptr;
    }
    if (n_threads <= 1) return;
    impl_ = new Impl();
    for (int i = 0; i < n_threads - 1; ++i) impl_->workers.emplace_back([this] { impl_->worker(); });
}

void HostPool::run(size_t n_items, size_t grain, const std::function<void(size_t)>& fn) {
    if (n_items == 0) return;
    if (!impl_) {
        for (size_t i = 0; i < n_items; ++i) fn(i);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->fn = fn;
        impl_->n_items = n_items;
        impl_->grain = grain ? grain : 1;
        impl_->next.store(0, std::memory_order_relaxed);
        impl_->active = (int)impl_->workers.size();
        ++impl_->generation;
    }
    impl_->cv_work.notify_all();
    impl_->drain();  // the caller works too
    std::unique_lock<std::mutex> lk(impl_->mu);
    impl_->cv_done.wait(lk, [&] { return impl_->active == 0; });
}

void HostPool::start(size_t n_items, size_t grain, const std::function<void(size_t)>& fn) {
    if (n_items == 0) return;
    if (!impl_) {
        for (size_t i = 0; i < n_items; ++i) fn(i);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->fn = fn;
        impl_->n_items = n_items;
        impl_->grain = grain ? grain : 1;
        impl_->next.store(0, std::memory_order_relaxed);
        impl_->active = (int)impl_->workers.size();
        ++impl_->generation;
    }
    impl_->cv_work.notify_all();
}

void HostPool::wait() {
    if (!impl_) return;
    std::unique_lock<std::mutex> lk(impl_->mu);
    impl_->cv_done.wait(lk, [&] { return impl_->active == 0; });
}

int default_host_threads() {
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = getenv("HB_HOST_THREADS")) {
        const int v = atoi(e);
        if (v >= 1) return v > 256 ? 256 : v;
    }
    if (hw == 0) hw = 1;
    return (int)(hw > 32 ? 32 : hw);
}

}  // namespace hb
