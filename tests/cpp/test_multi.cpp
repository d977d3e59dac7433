// One process, several GPUs through hb_multi (include/herbgpu.h): the region meshed by N devices at once must be
// byte-identical, chunk by chunk, to the same region meshed by one device.  Exits 77 when the box has fewer than two
// GPUs (unless --allow-one, which still exercises the handle with a single device).
//   test_multi [--allow-one] [--columns N]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "../../include/herbgpu.h"

struct Collected {
    std::map<uint64_t, std::string> chunks;  // region chunk index -> raw Instance3d bytes
    std::map<uint64_t, std::string> ids;
    std::atomic<int> inside{0};
    int overlapped = 0, calls = 0;
};

static void sink(void* user, uint64_t first, uint64_t n, const uint64_t* off, const uint32_t* cnt, const hb_instance3d* inst,
                 uint64_t, const uint16_t* ids) {
    Collected* c = static_cast<Collected*>(user);
    if (c->inside.fetch_add(1) != 0) c->overlapped++;  // the handle promises serialised sink calls
    c->calls++;
    for (uint64_t k = 0; k < n; ++k) {
        c->chunks[first + k] = std::string(reinterpret_cast<const char*>(inst + off[k]), (size_t)cnt[k] * sizeof(hb_instance3d));
        if (ids) c->ids[first + k] = std::string(reinterpret_cast<const char*>(ids + k * HB_CHUNK_VOLUME), HB_CHUNK_VOLUME * 2);
    }
    c->inside.fetch_sub(1);
}

#define CHECK(x)                                                                  \
    do {                                                                          \
        int rc__ = (x);                                                           \
        if (rc__ != HB_OK) {                                                      \
            fprintf(stderr, "%s -> %d: %s\n", #x, rc__, hb_last_error());         \
            return 1;                                                             \
        }                                                                         \
    } while (0)

int main(int argc, char** argv) {
    bool allow_one = false;
    int cols = 24;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--allow-one")) allow_one = true;
        if (!strcmp(argv[i], "--columns") && i + 1 < argc) cols = atoi(argv[++i]);
    }
    hb_multi* m = nullptr;
    CHECK(hb_multi_create(nullptr, 0, &m));
    const int n = hb_multi_device_count(m);
    if (n < 2 && !allow_one) {
        printf("SKIP: %d GPU visible\n", n);
        hb_multi_destroy(m);
        return 77;
    }
    CHECK(hb_multi_set_world(m, 5, 0.001f, 6, 2.0f, 0.6f));
    hb_region region{-7, 3, cols, 9, -2, 4};  // cx0, cz0, ncx, ncz, cy0, ncy
    // the split is contiguous, balanced and covers the region
    int32_t prev = 0;
    for (int i = 0; i < n; ++i) {
        int32_t lo, hi;
        CHECK(hb_multi_slab(m, &region, i, &lo, &hi));
        if (lo != prev || hi < lo || hi - lo > cols / n + 1) { fprintf(stderr, "bad slab %d: [%d,%d)\n", i, lo, hi); return 1; }
        prev = hi;
    }
    if (prev != cols) { fprintf(stderr, "slabs do not cover the region\n"); return 1; }

    Collected multi, single;
    uint64_t tc = 0, tq = 0, sc = 0, sq = 0;
    CHECK(hb_multi_generate_mesh_region(m, &region, 1, sink, &multi, &tc, &tq));
    // the same region on one device through the plain entry point
    CHECK(hb_generate_mesh_region(hb_multi_ctx(m, 0), &region, 1, sink, &single, &sc, &sq));
    const uint64_t want = (uint64_t)cols * 9 * 4;
    if (tc != want || sc != want || tq != sq || tq == 0) { fprintf(stderr, "totals differ: %llu/%llu chunks, %llu/%llu quads\n",
        (unsigned long long)tc, (unsigned long long)sc, (unsigned long long)tq, (unsigned long long)sq); return 1; }
    if (multi.overlapped) { fprintf(stderr, "sink ran concurrently %d times\n", multi.overlapped); return 1; }
    if (multi.chunks.size() != want || single.chunks.size() != want) { fprintf(stderr, "missing chunks\n"); return 1; }
    for (uint64_t g = 0; g < want; ++g) {
        if (multi.chunks[g] != single.chunks[g]) { fprintf(stderr, "chunk %llu: quads differ\n", (unsigned long long)g); return 1; }
        if (multi.ids[g] != single.ids[g]) { fprintf(stderr, "chunk %llu: block ids differ\n", (unsigned long long)g); return 1; }
    }
    // both transports, and a second call on the warm handle
    CHECK(hb_multi_set_region_transport(m, HB_TRANSPORT_INSTANCES));
    Collected raw;
    CHECK(hb_multi_generate_mesh_region(m, &region, 0, sink, &raw, &tc, &tq));
    for (uint64_t g = 0; g < want; ++g)
        if (raw.chunks[g] != single.chunks[g]) { fprintf(stderr, "chunk %llu: raw transport differs\n", (unsigned long long)g); return 1; }
    // errors: a bad device list, a duplicate, a region outside the coordinate limit (message names the device)
    hb_multi* bad = nullptr;
    int32_t dup[2] = {0, 0}, oob[1] = {99};
    if (hb_multi_create(dup, 2, &bad) != HB_ERR_INVALID || hb_multi_create(oob, 1, &bad) != HB_ERR_INVALID) { fprintf(stderr, "bad device lists accepted\n"); return 1; }
    hb_region far{300000, 0, 2, 2, 0, 1};
    if (hb_multi_generate_mesh_region(m, &far, 0, nullptr, nullptr, nullptr, nullptr) != HB_ERR_RANGE || !strstr(hb_last_error(), "device")) {
        fprintf(stderr, "range error not reported: %s\n", hb_last_error()); return 1; }
    hb_multi_destroy(m);
    printf("OK: %d device(s), %llu chunks, %llu quads byte-exact vs one device, %d sink calls serialised\n", n,
           (unsigned long long)want, (unsigned long long)sq, multi.calls);
    return 0;
}
