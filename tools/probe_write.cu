// Write-only HBM bandwidth probes (sm_100a): what is the ceiling for a kernel that only writes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_probe_write tools/probe_write.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void k_st(uint4* p, size_t n, int mode) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (mode == 0) p[i] = v;
        else if (mode == 1) __stcs(p + i, v);
        else if (mode == 2) __stwt(p + i, v);
        else {  // incompressible-ish data: a hash of the index
            uint32_t h = (uint32_t)i * 2654435761u;
            __stcs(p + i, make_uint4(h, h ^ 0x9e3779b9u, h * 31u, ~h));
        }
    }
}

template <int BYTES>
__global__ void k_bulk(char* p, size_t nwin) {
    extern __shared__ __align__(128) char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    char* ws = sm + warp * BYTES;
    for (int i = lane * 16; i < BYTES; i += 32 * 16) *reinterpret_cast<uint4*>(ws + i) = make_uint4(lane, i, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ws);
    for (size_t w = (size_t)blockIdx.x * nw + warp; w < nwin; w += (size_t)gridDim.x * nw) {
        // refresh the window with index-dependent data (as a real producer would)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int i = lane * 16; i < BYTES; i += 32 * 16) { uint32_t h = (uint32_t)(w * 977 + i) * 2654435761u; *reinterpret_cast<uint4*>(ws + i) = make_uint4(h, ~h, h * 31u, h ^ 0x55u); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                         :: "l"(p + w * BYTES), "r"(saddr), "r"(BYTES) : "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The emit kernel's pattern: every CTA streams through its own contiguous region of `region` bytes (a chunk's
// instances), regions handed out CTA-strided; inside a region the CTA's warps take consecutive 3200-byte windows.
// `align` shifts every region start by a multiple of 400 bytes (16-byte but not 128-byte aligned, like chunk starts).
__global__ void k_bulk_regions(char* p, size_t nregions, size_t region, int misalign) {
    extern __shared__ __align__(128) char sm[];
    constexpr int BYTES = 3200;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    char* ws = sm + warp * BYTES;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ws);
    const size_t wins = region / BYTES;
    for (size_t r = blockIdx.x; r < nregions; r += gridDim.x) {
        const size_t shift = misalign ? ((r * 7) % 8) * 400 : 0;
        char* base = p + r * region + shift;
        if (misalign == 2) {
            // region still starts at a multiple of 400 B, but the windows are cut at absolute multiples of 3200 B:
            // a short first window, then 128-byte-aligned full windows (the last one short)
            char* end = base + wins * BYTES;
            char* first_full = p + ((size_t)(base - p) + BYTES - 1) / BYTES * BYTES;
            for (size_t w = warp;; w += nw) {
                char* a = w == 0 ? base : first_full + (w - 1) * BYTES;
                if (a >= end) break;
                char* b = w == 0 ? first_full : a + BYTES;
                if (b > end) b = end;
                const uint32_t nb = (uint32_t)(b - a);
                if (nb == 0) continue;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                for (int i = lane * 16; i < BYTES; i += 32 * 16) { uint32_t h = (uint32_t)(w * 977 + i + r) * 2654435761u; *reinterpret_cast<uint4*>(ws + i) = make_uint4(h, ~h, h * 31u, h ^ 0x55u); }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                                 :: "l"(a), "r"(saddr), "r"(nb) : "memory");
            }
            continue;
        }
        for (size_t w = warp; w < wins; w += nw) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            for (int i = lane * 16; i < BYTES; i += 32 * 16) { uint32_t h = (uint32_t)(w * 977 + i + r) * 2654435761u; *reinterpret_cast<uint4*>(ws + i) = make_uint4(h, ~h, h * 31u, h ^ 0x55u); }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                             :: "l"(base + w * BYTES), "r"(saddr), "r"(BYTES) : "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 10;
}

int main() {
    const size_t bytes = (size_t)8 << 30;
    char* d; cudaMalloc(&d, bytes);
    const char* names[4] = {"st.global.v4 (default)", "st.global.cs.v4", "st.global.wt.v4", "st.cs.v4 hashed data"};
    for (int ctas = 2; ctas <= 16; ctas *= 2)
        for (int m = 0; m < 4; ++m) {
            float ms = timeit([&] { k_st<<<148 * ctas, 256>>>((uint4*)d, bytes / 16, m); });
            printf("%-26s %2d CTA/SM: %.0f GB/s\n", names[m], ctas, bytes / ms / 1e6);
        }
    float ms = timeit([&] { cudaMemsetAsync(d, 1, bytes); });
    printf("cudaMemsetAsync: %.0f GB/s\n", bytes / ms / 1e6);
    cudaFuncSetAttribute(k_bulk<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384);
    for (int ctas = 1; ctas <= 8; ctas *= 2) {
        ms = timeit([&] { k_bulk<3200><<<148 * ctas, 256, 8 * 3200>>>(d, bytes / 3200); });
        printf("bulk 3200 B/warp, %d CTA/SM: %.0f GB/s\n", ctas, bytes / 3200 * 3200 / ms / 1e6);
        if (ctas > 3) continue;
        ms = timeit([&] { k_bulk<16384><<<148 * ctas, 128, 4 * 16384>>>(d, bytes / 16384); });
        printf("bulk 16 KB/warp,  %d CTA/SM: %.0f GB/s\n", ctas, bytes / ms / 1e6);
    }
    for (size_t region : {(size_t)102400, (size_t)163200, (size_t)1632000, (size_t)16320000})
        for (int mis = 0; mis < 3; ++mis) {
            const size_t nreg = (bytes - 4096) / region;
            ms = timeit([&] { k_bulk_regions<<<148 * 4, 256, 8 * 3200>>>(d, nreg, region, mis); });
            printf("bulk 3200 B/warp, 4 CTA/SM, per-CTA regions of %zu B, misalign %d: %.0f GB/s\n", region, mis,
                   (double)nreg * (region / 3200 * 3200) / ms / 1e6);
        }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
