// probe_hostbw.cpp -- host stream-write ceiling for the packed-quad expander (tools, not product).
//   g++ -O2 -pthread -o tools/_probe_hostbw tools/probe_hostbw.cpp && tools/_probe_hostbw [GiB per thread-set]
// For T in {1,2,4,8,16,...,all}: T threads each fill their own slice of a buffer with (a) non-temporal 16-byte
// stores, (b) plain stores (memset).  Prints GB/s.  The expander writes 100 B per quad, so its ceiling is the
// non-temporal figure.
#include <emmintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void fill_nt(char* p, size_t bytes, int v) {
    const __m128i x = _mm_set1_epi32(v);
    for (size_t i = 0; i + 64 <= bytes; i += 64) {
        _mm_stream_si128((__m128i*)(p + i), x);
        _mm_stream_si128((__m128i*)(p + i + 16), x);
        _mm_stream_si128((__m128i*)(p + i + 32), x);
        _mm_stream_si128((__m128i*)(p + i + 48), x);
    }
    _mm_sfence();
}

int main(int argc, char** argv) {
    const size_t total = (size_t)(argc > 1 ? atof(argv[1]) : 4.0) * (1ull << 30);
    const unsigned hw = std::thread::hardware_concurrency();
    char* buf = (char*)aligned_alloc(4096, total);
    if (!buf) return 1;
    memset(buf, 1, total);  // fault the pages in
    printf("hardware_concurrency %u, buffer %.1f GiB\n", hw, total / 1073741824.0);
    std::vector<unsigned> ts;
    for (unsigned t = 1; t < hw; t *= 2) ts.push_back(t);
    ts.push_back(hw);
    for (int mode = 0; mode < 2; ++mode) {
        for (unsigned T : ts) {
            double best = 1e30;
            for (int rep = 0; rep < 3; ++rep) {
                const size_t slice = (total / T) & ~(size_t)4095;
                std::vector<std::thread> th;
                const double t0 = now();
                for (unsigned i = 0; i < T; ++i)
                    th.emplace_back([=] {
                        if (mode == 0) fill_nt(buf + i * slice, slice, rep);
                        else memset(buf + i * slice, rep, slice);
                    });
                for (auto& t : th) t.join();
                const double dt = now() - t0;
                if (dt < best) best = dt;
            }
            const size_t slice = (total / T) & ~(size_t)4095;
            printf("%s threads %3u : %7.1f GB/s\n", mode == 0 ? "nt-store" : "memset  ", T, slice * T / best / 1e9);
        }
    }
    free(buf);
    return 0;
}
