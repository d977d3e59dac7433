// noise.cuh -- simplex / FBM device functions with the reference's exact f32 semantics.
//
// Follows the AVX2 column of crates/simd-noise (the backend an x86 host selects,
// simd/invoking.rs:169-199): every operator is a separate IEEE-754 round-to-nearest op
// (simd/ops/f32.rs:3-70) EXCEPT the six neg_mul_add sites of simplex_2d
// (functions/simplex_32.rs:144-146), which are true FMAs (simd/ops/f32.rs:141-145).
// All arithmetic below uses __fadd_rn/__fmul_rn/__fmaf_rn, which nvcc never contracts, so the
// results are bit-identical to that path (and this file is also built with -fmad=false).
#pragma once

#include "common.cuh"

namespace hb {

// functions/simplex_32.rs:29-46: the classic permutation; the reference stores it twice (512 x i32),
// we keep 256 bytes and mask the index (P[i] == P[i & 255] for i < 512).
__device__ const uint8_t kPerm[256] = {
    151, 160, 137, 91,  90,  15,  131, 13,  201, 95,  96,  53,  194, 233, 7,   225, 140, 36,  103, 30,  69,  142,
    8,   99,  37,  240, 21,  10,  23,  190, 6,   148, 247, 120, 234, 75,  0,   26,  197, 62,  94,  252, 219, 203,
    117, 35,  11,  32,  57,  177, 33,  88,  237, 149, 56,  87,  174, 20,  125, 136, 171, 168, 68,  175, 74,  165,
    71,  134, 139, 48,  27,  166, 77,  146, 158, 231, 83,  111, 229, 122, 60,  211, 133, 230, 220, 105, 92,  41,
    55,  46,  245, 40,  244, 102, 143, 54,  65,  25,  63,  161, 1,   216, 80,  73,  209, 76,  132, 187, 208, 89,
    18,  169, 200, 196, 135, 130, 116, 188, 159, 86,  164, 100, 109, 198, 173, 186, 3,   64,  52,  217, 226, 250,
    124, 123, 5,   202, 38,  147, 118, 126, 255, 82,  85,  212, 207, 206, 59,  227, 47,  16,  58,  17,  182, 189,
    28,  42,  223, 183, 170, 213, 119, 248, 152, 2,   44,  154, 163, 70,  221, 153, 101, 155, 167, 43,  172, 9,
    129, 22,  39,  253, 19,  98,  108, 110, 79,  113, 224, 232, 178, 185, 112, 104, 218, 246, 97,  228, 251, 34,
    242, 193, 238, 210, 144, 12,  191, 179, 162, 241, 81,  51,  145, 235, 249, 14,  239, 107, 49,  192, 214, 31,
    181, 199, 106, 157, 184, 84,  204, 176, 115, 121, 50,  45,  127, 4,   150, 254, 138, 236, 205, 93,  222, 114,
    67,  29,  24,  72,  243, 141, 128, 195, 78,  66,  215, 61,  156, 180};

// simplex_32.rs:6-15
#define HB_F2 0.36602540378f
#define HB_G2 0.2113248654f
#define HB_G22 (0.2113248654f * 2.0f)

// functions/gradient_32.rs:13-30: the gradient (gx, gy) selected by a hash value.
// mask.blendv(a, b) = mask ? b : a (simd/overloads.rs:588-591).
__device__ __forceinline__ float2 grad2_of(int seed_lo, int hash) {
    const int h = (hash ^ seed_lo) & 7;
    const bool lt4 = h < 4;
    const float xm = lt4 ? 1.0f : 2.0f;
    const float ym = lt4 ? 2.0f : 1.0f;
    const bool b1 = (h & 1) == 0;
    const bool b2 = (h & 2) == 0;
    return make_float2((lt4 ? b1 : b2) ? xm : -xm,   // 0.0 - m == -m exactly
                       (lt4 ? b2 : b1) ? ym : -ym);
}

// Shared-memory tables for the 2-D path: the permutation (256 B) and, per permutation slot, the
// gradient it selects for this seed (G[i] = grad2(P[i]), 2 KB).  A corner's gradient is then two
// dependent shared loads instead of a load plus ~12 select/logic instructions.
struct NoiseTables {
    __align__(16) uint8_t perm[256];
    __align__(16) float2 grad[256];
};

// Cooperative table setup (call from all threads, then __syncthreads()).
__device__ __forceinline__ void load_noise_tables(NoiseTables& t, int seed_lo) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int p = kPerm[i];
        t.perm[i] = (uint8_t)p;
        t.grad[i] = grad2_of(seed_lo, p);
    }
}

// functions/simplex_32.rs:103-170 (value half of simplex_2d_deriv)
__device__ __forceinline__ float simplex2(const NoiseTables& T, float x, float y) {
    const float s = __fmul_rn(HB_F2, __fadd_rn(x, y));
    const float ips = floorf(__fadd_rn(x, s));
    const float jps = floorf(__fadd_rn(y, s));
    const int i = __float2int_rz(ips);  // exact: ips is integral (cast_i32, simd/ops/f32.rs:716-739)
    const int j = __float2int_rz(jps);
    const float t = __fmul_rn(__int2float_rn(i + j), HB_G2);
    const float x0 = __fsub_rn(x, __fsub_rn(ips, t));
    const float y0 = __fsub_rn(y, __fsub_rn(jps, t));

    const bool xge = x0 >= y0;  // i1 = all-ones mask (-1) when x0 >= y0
    const bool ygt = y0 > x0;   // j1
    const float x1 = __fadd_rn(__fadd_rn(x0, xge ? -1.0f : 0.0f), HB_G2);
    const float y1 = __fadd_rn(__fadd_rn(y0, ygt ? -1.0f : 0.0f), HB_G2);
    const float x2 = __fadd_rn(__fadd_rn(x0, -1.0f), HB_G22);
    const float y2 = __fadd_rn(__fadd_rn(y0, -1.0f), HB_G22);

    // gi_k = P[(ii + di) + P[jj + dj]] (simplex_32.rs:137-141); indices are taken mod 256 because the
    // reference's 512-entry table is the permutation twice
    const int pj0 = T.perm[j & 255];
    const int pj1 = T.perm[(j + 1) & 255];
    const float2 g0 = T.grad[(i + pj0) & 255];
    const float2 g1 = T.grad[(i + (xge ? 1 : 0) + (ygt ? pj1 : pj0)) & 255];
    const float2 g2 = T.grad[(i + 1 + pj1) & 255];

    // neg_mul_add(a, b, c) = fma(-a, b, c): _mm256_fnmadd_ps (simd/ops/f32.rs:141-145)
    float t0 = __fmaf_rn(-y0, y0, __fmaf_rn(-x0, x0, 0.5f));
    float t1 = __fmaf_rn(-y1, y1, __fmaf_rn(-x1, x1, 0.5f));
    float t2 = __fmaf_rn(-y2, y2, __fmaf_rn(-x2, x2, 0.5f));
    // t &= t.cmp_gte(0): negative -> +0.  fmaxf(t, 0) differs only for t = -0 (keeps/loses the sign
    // of zero), and t is squared next, so the result is identical.
    t0 = fmaxf(t0, 0.0f);
    t1 = fmaxf(t1, 0.0f);
    t2 = fmaxf(t2, 0.0f);

    const float t20 = __fmul_rn(t0, t0), t40 = __fmul_rn(t20, t20);
    const float t21 = __fmul_rn(t1, t1), t41 = __fmul_rn(t21, t21);
    const float t22 = __fmul_rn(t2, t2), t42 = __fmul_rn(t22, t22);

    // g = gx*x + gy*y: two multiplies and one add, as the operator overloads do (never fused)
    const float n0 = __fmul_rn(t40, __fadd_rn(__fmul_rn(g0.x, x0), __fmul_rn(g0.y, y0)));
    const float n1 = __fmul_rn(t41, __fadd_rn(__fmul_rn(g1.x, x1), __fmul_rn(g1.y, y1)));
    const float n2 = __fmul_rn(t42, __fadd_rn(__fmul_rn(g2.x, x2), __fmul_rn(g2.y, y2)));
    return __fmul_rn(__fadd_rn(n0, __fadd_rn(n1, n2)), 45.26450774985561631259f);
}

// functions/fbm_32.rs:18-31
__device__ __forceinline__ float fbm2(const NoiseTables& T, float x, float y, const WorldDev& w) {
    float result = simplex2(T, x, y);
    float amp = 1.0f;
    for (int o = 1; o < w.octaves; ++o) {
        x = __fmul_rn(x, w.lacunarity);
        y = __fmul_rn(y, w.lacunarity);
        amp = __fmul_rn(amp, w.gain);
        result = __fadd_rn(__fmul_rn(simplex2(T, x, y), amp), result);
    }
    return result;
}

// `(noise * 32.0) as i32` (server/src/generator/mod.rs:79): truncating, saturating, NaN -> 0,
// which is exactly cvt.rzi.s32.f32.
__device__ __forceinline__ int height_from_noise(float v) { return __float2int_rz(__fmul_rn(v, 32.0f)); }

// ---- 3-D: functions/simplex_32.rs:197-282, hash3d_32.rs:21-36, gradient_32.rs:32-47 (no FMAs) ----
#define HB_F3 (1.0f / 3.0f)
#define HB_G3 (1.0f / 6.0f)
#define HB_G33 (3.0f / 6.0f - 1.0f)

__device__ __forceinline__ float grad3d_dot(int seed_lo, int i, int j, int k, float x, float y, float z) {
    uint32_t hash = (uint32_t)i ^ (uint32_t)seed_lo;
    hash = (uint32_t)j ^ hash;
    hash = (uint32_t)k ^ hash;
    hash = ((hash * hash) * 60493u) * hash;
    hash = (hash >> 13) ^ hash;
    const uint32_t a13 = hash & 13u;
    const float u = (a13 < 8u) ? x : y;
    const float v = (a13 < 2u) ? y : ((a13 == 12u) ? x : z);
    return __fadd_rn(__uint_as_float(__float_as_uint(u) ^ (hash << 31)),
                     __uint_as_float(__float_as_uint(v) ^ ((hash & 2u) << 30)));
}

__device__ __forceinline__ float falloff3(float x, float y, float z) {
    // 0.6 - x*x - y*y - z*z, left to right, unfused
    float t = __fsub_rn(__fsub_rn(__fsub_rn(0.6f, __fmul_rn(x, x)), __fmul_rn(y, y)), __fmul_rn(z, z));
    t = (t >= 0.0f) ? t : 0.0f;
    const float t2 = __fmul_rn(t, t);
    return __fmul_rn(t2, t2);
}

__device__ __forceinline__ float simplex3(float x, float y, float z, int seed_lo) {
    const float f = __fmul_rn(HB_F3, __fadd_rn(__fadd_rn(x, y), z));
    float x0 = floorf(__fadd_rn(x, f));
    float y0 = floorf(__fadd_rn(y, f));
    float z0 = floorf(__fadd_rn(z, f));
    const int i = __float2int_rz(x0) * 1619;  // X/Y/Z_PRIME_32, cellular_32.rs:8,11,14 (wrapping)
    const int j = __float2int_rz(y0) * 31337;
    const int k = __float2int_rz(z0) * 6971;
    const float g = __fmul_rn(HB_G3, __fadd_rn(__fadd_rn(x0, y0), z0));
    x0 = __fsub_rn(x, __fsub_rn(x0, g));
    y0 = __fsub_rn(y, __fsub_rn(y0, g));
    z0 = __fsub_rn(z, __fsub_rn(z0, g));

    const bool xy = x0 >= y0, yz = y0 >= z0, xz = x0 >= z0;
    const bool i1 = xy && xz, j1 = yz && !xy, k1 = !yz && !xz;
    const bool i2 = xy || xz, j2 = !xy || yz, k2 = !(xz && yz);

    const float x1 = __fadd_rn(__fsub_rn(x0, i1 ? 1.0f : 0.0f), HB_G3);
    const float y1 = __fadd_rn(__fsub_rn(y0, j1 ? 1.0f : 0.0f), HB_G3);
    const float z1 = __fadd_rn(__fsub_rn(z0, k1 ? 1.0f : 0.0f), HB_G3);
    const float x2 = __fadd_rn(__fsub_rn(x0, i2 ? 1.0f : 0.0f), HB_F3);
    const float y2 = __fadd_rn(__fsub_rn(y0, j2 ? 1.0f : 0.0f), HB_F3);
    const float z2 = __fadd_rn(__fsub_rn(z0, k2 ? 1.0f : 0.0f), HB_F3);
    const float x3 = __fadd_rn(x0, HB_G33);
    const float y3 = __fadd_rn(y0, HB_G33);
    const float z3 = __fadd_rn(z0, HB_G33);

    const float v0 = __fmul_rn(falloff3(x0, y0, z0), grad3d_dot(seed_lo, i, j, k, x0, y0, z0));
    const float v1 = __fmul_rn(falloff3(x1, y1, z1),
                               grad3d_dot(seed_lo, i + (i1 ? 1619 : 0), j + (j1 ? 31337 : 0), k + (k1 ? 6971 : 0), x1, y1, z1));
    const float v2 = __fmul_rn(falloff3(x2, y2, z2),
                               grad3d_dot(seed_lo, i + (i2 ? 1619 : 0), j + (j2 ? 31337 : 0), k + (k2 ? 6971 : 0), x2, y2, z2));
    const float v3 = __fmul_rn(falloff3(x3, y3, z3), grad3d_dot(seed_lo, i + 1619, j + 31337, k + 6971, x3, y3, z3));
    return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(v3, v2), v1), v0), 32.69587493801679f);
}

// functions/fbm_32.rs:33-47
__device__ __forceinline__ float fbm3(float x, float y, float z, const WorldDev& w) {
    float result = simplex3(x, y, z, w.seed_lo);
    float amp = 1.0f;
    for (int o = 1; o < w.octaves; ++o) {
        x = __fmul_rn(x, w.lacunarity);
        y = __fmul_rn(y, w.lacunarity);
        z = __fmul_rn(z, w.lacunarity);
        amp = __fmul_rn(amp, w.gain);
        result = __fadd_rn(__fmul_rn(simplex3(x, y, z, w.seed_lo), amp), result);
    }
    return result;
}

}  // namespace hb
