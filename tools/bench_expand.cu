// Host expander in isolation (no GPU, no CUDA call): records per second of csrc/expand.cu for a synthetic slab, beside a
// cached-store variant and a pure non-temporal fill of the same bytes (the host memory ceiling for this pattern).
//   nvcc -O3 -std=c++17 -Xcompiler -O2 -Iinclude -o tools/_bench_expand_bin tools/bench_expand.cu && tools/_bench_expand_bin <threads>
#include "../herbolution_b200/csrc/expand.cu"
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
using namespace hb;
namespace hb {
// variant: regular (cached) stores instead of non-temporal ones
void expand_cached(const ExpandTables& T, hb_chunk_pt pt, const uint32_t* __restrict__ packed, uint32_t n, hb_instance3d* out) {
    const uint64_t seed = chunk_color_seed(pt.x, pt.y, pt.z);
    const __m128i base = _mm_set_epi32(1, pt.z * 32, pt.y * 32, pt.x * 32);
    float* dst = reinterpret_cast<float*>(out);
    for (uint32_t q = 0; q < n; ++q) {
        const uint32_t w = packed[q];
        const uint32_t L = w & 0x7FFFu, f = (w >> 15) & 7u, ao = (w >> 18) & 255u, mi = (w >> 26) & 15u;
        float* r = dst + (size_t)q * 25;
        const float* fm = T.face_mat[f < 6 ? f : 0];
        _mm_storeu_ps(r, _mm_loadu_ps(fm)); _mm_storeu_ps(r + 4, _mm_loadu_ps(fm + 4)); _mm_storeu_ps(r + 8, _mm_loadu_ps(fm + 8));
        const __m128i loc = _mm_set_epi32(0, (int)((L >> 5) & 31u), (int)(L & 31u), (int)(L >> 10));
        _mm_storeu_ps(r + 12, _mm_cvtepi32_ps(_mm_add_epi32(base, loc)));
        const uint64_t s = seed + (uint64_t)(6u * L + f + 1u) * 0x2d358dccaa6c78a5ULL;
        const unsigned __int128 t = (unsigned __int128)s * (s ^ 0x8bb84b93962eacc9ULL);
        const uint32_t bits = 0x3F800000u | ((uint32_t)((uint64_t)t ^ (uint64_t)(t >> 64)) >> 9);
        float p; memcpy(&p, &bits, 4); p -= 1.0f;
        const int ci = (int)(T.color_scale[mi] * p);
        _mm_storeu_ps(r + 16, _mm_load_ps(T.rgba[mi][ci]));
        r[20] = 0.0f;
        _mm_storeu_ps(r + 21, _mm_load_ps(T.ao4[ao]));
    }
}
// variant: the same bytes per chunk, pure non-temporal fill (no per-quad work)
void fill_only(const ExpandTables&, hb_chunk_pt, const uint32_t* __restrict__, uint32_t n, hb_instance3d* out) {
    const __m128 v = _mm_set1_ps(1.0f);
    float* dst = reinterpret_cast<float*>(out);
    const size_t n16 = (size_t)((n + 3) & ~3u) * 100 / 16;
    for (size_t i = 0; i < n16; ++i) _mm_stream_ps(dst + i * 4, v);
    _mm_sfence();
}
}
typedef void (*fn_t)(const ExpandTables&, hb_chunk_pt, const uint32_t*, uint32_t, hb_instance3d*);
int main(int argc, char** argv) {
    const int nthreads = argc > 1 ? atoi(argv[1]) : 16;
    MeshTables tb{}; for (int f = 0; f < 6; ++f) for (int i = 0; i < 12; ++i) tb.face_mat[f][i] = (float)(f * 12 + i) * 0.01f;
    for (int k = 0; k < 4; ++k) tb.ao_lut[k] = 1.0f - k * 0.3f;
    MaterialsDev m{}; m.n_materials = 3;
    for (int i = 0; i < 3; ++i) { m.n_colors[i] = 3; for (int c = 0; c < 3; ++c) for (int k = 0; k < 4; ++k) m.rgba[i][c][k] = 0.1f * (i + c + k); }
    ExpandTables T; T.build(tb, m);
    const uint32_t nq = 1080, nchunks = 16384;   // ~1.77 GB of records, like a few slabs
    std::vector<uint32_t> packed((size_t)nq * 64);
    for (auto& q : packed) q = (rand() & 0x7FFF) | ((rand() % 6) << 15) | ((rand() & 255) << 18) | ((rand() % 3) << 26);
    const size_t bytes = (size_t)nchunks * nq * 100;
    char* A = (char*)aligned_alloc(2 << 20, (bytes + (2 << 20)) & ~(size_t)((2 << 20) - 1));
    madvise(A, bytes, MADV_HUGEPAGE);
    memset(A, 0, bytes);
    HostPool pool; pool.resize(nthreads);
    struct { const char* name; fn_t f; } v[] = {{"expand (nt stores)", expand_chunk}, {"expand (cached stores)", expand_cached}, {"nt fill only", fill_only}};
    for (auto& e : v) for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        pool.run(nchunks, 8, [&](size_t i) { e.f(T, hb_chunk_pt{(int)i, -1, 3}, packed.data() + (i & 63) * nq, nq, (hb_instance3d*)A + i * nq); });
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rep) printf("%-24s threads %2d: %6.1f GB/s  (%.2f ns/quad/thread)\n", e.name, nthreads, bytes / dt / 1e9, dt * 1e9 * nthreads / ((double)nchunks * nq));
    }
    // cached stores into a reused window of F MiB (does a cache-resident hand-over buffer beat the DRAM write path?)
    for (int fmib : {4, 8, 16, 32, 64, 128}) for (int rep = 0; rep < 2; ++rep) {
        const size_t ring = ((size_t)fmib << 20) / ((size_t)nq * 100);
        auto t0 = std::chrono::steady_clock::now();
        pool.run(nchunks, 8, [&](size_t i) { expand_cached(T, hb_chunk_pt{(int)i, -1, 3}, packed.data() + (i & 63) * nq, nq, (hb_instance3d*)A + (i % ring) * nq); });
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rep) printf("cached, %3d MiB window    threads %2d: %6.1f GB/s\n", fmib, nthreads, bytes / dt / 1e9);
    }
    return 0;
}
