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

// Shared-memory tables for the 2-D path, laid out so that no index needs a mask or a shift:
//   perm4[k] = 4 * P[k & 255]  (u16, 258 entries used: k = jj and jj + 1 with jj in 0..255)
//   gx[k], gy[k] = grad2(P[k & 255]) for this seed, k in 0..511 (k = ii + P[..] (+1) with ii in 0..255)
// The reference's own table is the permutation twice (512 x i32, simplex_32.rs:29-46) for the same reason.
// A corner's gradient is two dependent shared loads instead of a load plus ~12 select/logic instructions.  gx and gy
// are separate arrays so that the values of two samples land in adjacent registers (operands of the packed ops).
struct NoiseTables {
    __align__(16) uint16_t perm4[512];
    __align__(16) float gx[512];
    __align__(16) float gy[512];
};

// Cooperative table setup (call from all threads, then __syncthreads()).
__device__ __forceinline__ void load_noise_tables(NoiseTables& t, int seed_lo) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const int p = kPerm[i & 255];
        const float2 g = grad2_of(seed_lo, p);
        t.perm4[i] = (uint16_t)(p * 4);
        t.gx[i] = g.x;
        t.gy[i] = g.y;
    }
}

// Packed helpers: Blackwell's FADD2 / FMUL2 / FFMA2 (add/mul/fma.rn.f32x2) process two f32 per instruction with the
// same per-component IEEE round-to-nearest results as the scalar ops, and fold operand negation and broadcast
// immediates, so a - b is written as a + (-b) (exactly the same value).
//
// CAUTION (measured with CUDA 12.9): ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 even with
// --fmad=false, which would break bit-exactness against the reference (its operator overloads never fuse).  It does
// not contract across a packed/scalar boundary.  So every product that feeds an add is either computed with scalar
// FMULs (smul2) or consumed by scalar FADDs (sadd2); the SASS must contain exactly six FFMA2 per simplex pair (the
// neg_mul_add sites), which tests/test_abi.py checks, and the GPU parity tests compare the noise bit for bit.
__device__ __forceinline__ float2 pk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fnma2(float2 a, float2 b, float2 c) { return __ffma2_rn(neg2(a), b, c); }  // c - a*b, fused
__device__ __forceinline__ float2 smul2(float2 a, float2 b) { return pk(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ float2 sadd2(float2 a, float2 b) { return pk(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
// neg_mul_add(a, b, c) = c - a*b (simd/ops/f32.rs:141-162): one fused operation on AVX2+FMA and NEON hosts
// (_mm256_fnmadd_ps / vfmsq_f32), a rounded product and a subtraction on SSE2 / SSE4.1 / scalar / wasm hosts.  FMA
// selects which of the reference's CPU columns is reproduced (hb_set_fma_mode).
template <bool FMA>
__device__ __forceinline__ float2 nma2(float2 a, float2 b, float2 c) {
    if constexpr (FMA) return fnma2(a, b, c);
    else return sub2(c, smul2(a, b));  // scalar multiplies: a packed product feeding a packed add would be contracted
}

// c - a*a, clamped to [0, 1] (callers guarantee c <= 0.5): the clamp rides on the producing instruction
template <bool FMA>
__device__ __forceinline__ float2 nma_sat(float2 a, float2 c) {
    if constexpr (FMA) return pk(__saturatef(__fmaf_rn(-a.x, a.x, c.x)), __saturatef(__fmaf_rn(-a.y, a.y, c.y)));
    else return pk(__saturatef(__fsub_rn(c.x, __fmul_rn(a.x, a.x))), __saturatef(__fsub_rn(c.y, __fmul_rn(a.y, a.y))));
}
// gx*x + gy*y for gradient components in {+-1, +-2}
template <bool FMA>
__device__ __forceinline__ float2 gdot(float2 gx, float2 x, float2 gy, float2 y) {
    if constexpr (FMA) return __ffma2_rn(gx, x, mul2(gy, y));
    else return sadd2(mul2(gx, x), mul2(gy, y));
}

// functions/simplex_32.rs:103-170 (value half of simplex_2d_deriv) for TWO samples at once: every variable is a
// float2 (.x = sample A, .y = sample B).  The arithmetic per sample is op for op the reference's: separate
// round-to-nearest add/mul everywhere, true FMA only at the six neg_mul_add sites
// (_mm256_fnmadd_ps, simd/ops/f32.rs:141-145).  x and y must not be the direct result of a packed multiply.
//
// FAST (every lattice coordinate |floor(x + s)| < 2^20, which the launch wrapper proves from the world parameters and
// the chunk-coordinate limit): the integer side never leaves the FP32/ALU pipes -- no F2I / I2F (quarter-rate XU pipe):
//   * table indices come from "magic" adds: ips is integral, so ips + 1.5*2^21 is exact and its mantissa holds
//     (2^20 + ips) << 2; `& 1020` is (i & 255) * 4, the byte offset into gx/gy.  Likewise jps + 1.5*2^22 gives
//     (j & 255) * 2, the byte offset into perm4.
//   * t = f32(i + j) * G2 is computed as (ips + jps) * G2: the float sum of two integral values below 2^21 is exact
//     and equals the converted integer sum.
//   * j1 = (y0 > x0) is taken as NOT (x0 >= y0): identical for every non-NaN input, and coordinates are finite by
//     construction.
// Bit-exactness of the result is what tests/test_gpu_parity.py::test_noise_tiles_bit_exact asserts (0 ulp).
template <bool FAST, bool FMA = true>
__device__ __forceinline__ float2 simplex2_pair(const NoiseTables& T, float2 x, float2 y) {
    const float2 s = smul2(bc(HB_F2), add2(x, y));  // feeds x + s, y + s
    const float2 xs = add2(x, s), ys = add2(y, s);
    const float2 ips = pk(floorf(xs.x), floorf(xs.y)), jps = pk(floorf(ys.x), floorf(ys.y));
    float2 t, x0, y0, x1, y1, g0x, g0y, g1x, g1y, g2x, g2y;
    if constexpr (FAST) {
        const float2 mi = add2(ips, bc(3145728.0f));  // 1.5 * 2^21
        const float2 mj = add2(jps, bc(6291456.0f));  // 1.5 * 2^22
        t = smul2(add2(ips, jps), bc(HB_G2));         // feeds ips - t, jps - t
        x0 = sub2(x, sub2(ips, t));
        y0 = sub2(y, sub2(jps, t));
        const bool xgeA = x0.x >= y0.x, xgeB = x0.y >= y0.y;  // i1 = all-ones mask (-1) when x0 >= y0
        // x1 = (x0 + f32(i1)) + G2 with i1 = -1 / 0: x0 - 1 is needed for x2 anyway and x0 + 0 is x0 (a -0 would
        // only differ before the + G2), so the masks select between values that exist already; j1 = NOT i1
        const float2 xm = add2(x0, bc(-1.0f)), ym = add2(y0, bc(-1.0f));
        x1 = add2(pk(xgeA ? xm.x : x0.x, xgeB ? xm.y : x0.y), bc(HB_G2));
        y1 = add2(pk(xgeA ? y0.x : ym.x, xgeB ? y0.y : ym.y), bc(HB_G2));
        // gi_k = P[(ii + di) + P[jj + dj]] (simplex_32.rs:137-141) as byte offsets; k1 is k0 + 1 when i1 is set
        // (then j1 is not: P[jj]), else P[jj + 1] without the +1
        const char* const pb = reinterpret_cast<const char*>(T.perm4);
        const char* const gxb = reinterpret_cast<const char*>(T.gx);
        const char* const gyb = reinterpret_cast<const char*>(T.gy);
        const uint32_t iiA = __float_as_uint(mi.x) & 1020u, iiB = __float_as_uint(mi.y) & 1020u;
        const uint32_t jjA = __float_as_uint(mj.x) & 510u, jjB = __float_as_uint(mj.y) & 510u;
        const uint32_t a0A = iiA + *reinterpret_cast<const uint16_t*>(pb + jjA);
        const uint32_t a2A = iiA + *reinterpret_cast<const uint16_t*>(pb + jjA + 2);
        const uint32_t a0B = iiB + *reinterpret_cast<const uint16_t*>(pb + jjB);
        const uint32_t a2B = iiB + *reinterpret_cast<const uint16_t*>(pb + jjB + 2);
        const uint32_t a1A = xgeA ? a0A + 4u : a2A, a1B = xgeB ? a0B + 4u : a2B;
#define HB_G(tab, off) (*reinterpret_cast<const float*>((tab) + (off)))
        g0x = pk(HB_G(gxb, a0A), HB_G(gxb, a0B)); g0y = pk(HB_G(gyb, a0A), HB_G(gyb, a0B));
        g1x = pk(HB_G(gxb, a1A), HB_G(gxb, a1B)); g1y = pk(HB_G(gyb, a1A), HB_G(gyb, a1B));
        g2x = pk(HB_G(gxb, a2A + 4u), HB_G(gxb, a2B + 4u)); g2y = pk(HB_G(gyb, a2A + 4u), HB_G(gyb, a2B + 4u));
#undef HB_G
    } else {
        const int iA = __float2int_rz(ips.x), iB = __float2int_rz(ips.y);  // exact: integral (cast_i32, simd/ops/f32.rs:716-739)
        const int jA = __float2int_rz(jps.x), jB = __float2int_rz(jps.y);
        t = smul2(pk(__int2float_rn(iA + jA), __int2float_rn(iB + jB)), bc(HB_G2));  // feeds ips - t, jps - t
        x0 = sub2(x, sub2(ips, t));
        y0 = sub2(y, sub2(jps, t));
        const bool xgeA = x0.x >= y0.x, xgeB = x0.y >= y0.y;  // i1 = all-ones mask (-1) when x0 >= y0
        const bool ygtA = y0.x > x0.x, ygtB = y0.y > x0.y;    // j1
        x1 = add2(add2(x0, pk(xgeA ? -1.0f : 0.0f, xgeB ? -1.0f : 0.0f)), bc(HB_G2));
        y1 = add2(add2(y0, pk(ygtA ? -1.0f : 0.0f, ygtB ? -1.0f : 0.0f)), bc(HB_G2));
        // indices are taken mod 256 because the reference's 512-entry table is the permutation twice
        const int pj0A = T.perm4[jA & 255] >> 2, pj1A = T.perm4[(jA + 1) & 255] >> 2;
        const int pj0B = T.perm4[jB & 255] >> 2, pj1B = T.perm4[(jB + 1) & 255] >> 2;
        const int k0A = (iA + pj0A) & 255, k0B = (iB + pj0B) & 255;
        const int k1A = (iA + (xgeA ? 1 : 0) + (ygtA ? pj1A : pj0A)) & 255;
        const int k1B = (iB + (xgeB ? 1 : 0) + (ygtB ? pj1B : pj0B)) & 255;
        const int k2A = (iA + 1 + pj1A) & 255, k2B = (iB + 1 + pj1B) & 255;
        g0x = pk(T.gx[k0A], T.gx[k0B]); g0y = pk(T.gy[k0A], T.gy[k0B]);
        g1x = pk(T.gx[k1A], T.gx[k1B]); g1y = pk(T.gy[k1A], T.gy[k1B]);
        g2x = pk(T.gx[k2A], T.gx[k2B]); g2y = pk(T.gy[k2A], T.gy[k2B]);
    }
    const float2 x2 = add2(add2(x0, bc(-1.0f)), bc(HB_G22));
    const float2 y2 = add2(add2(y0, bc(-1.0f)), bc(HB_G22));

    // t = (0.5 - x*x) - y*y, then t &= t.cmp_gte(0): negative -> +0.  t never exceeds 0.5, so the clamp is the
    // saturation modifier of the instruction that produces t (FFMA.SAT / FADD.SAT: [0, 1]) instead of a separate
    // FMNMX per value; it differs from the mask only for t = -0 (sign of zero), and t is squared next.
    float2 t0 = nma_sat<FMA>(y0, nma2<FMA>(x0, x0, bc(0.5f)));
    float2 t1 = nma_sat<FMA>(y1, nma2<FMA>(x1, x1, bc(0.5f)));
    float2 t2 = nma_sat<FMA>(y2, nma2<FMA>(x2, x2, bc(0.5f)));
    const float2 t20 = mul2(t0, t0), t40 = mul2(t20, t20);
    const float2 t21 = mul2(t1, t1), t41 = mul2(t21, t21);
    const float2 t22 = mul2(t2, t2), t42 = mul2(t22, t22);

    // g = gx*x + gy*y: two multiplies and one add in the reference (operator overloads, never fused).  The gradient
    // components are +-1 and +-2 (gradient_32.rs:13-30), so both products are EXACT, and a fused multiply-add of exact
    // terms rounds exactly once like the reference's add does: fma(gx, x, gy*y) is bit-identical (signed zeros
    // included) and one packed instruction shorter.  Used in the fused column only, so that the SSE2 / scalar column's
    // kernels stay free of any FMA (tests/test_abi.py).
    const float2 n0 = mul2(t40, gdot<FMA>(g0x, x0, g0y, y0));
    const float2 n1 = mul2(t41, gdot<FMA>(g1x, x1, g1y, y1));
    const float2 n2 = mul2(t42, gdot<FMA>(g2x, x2, g2y, y2));
    return mul2(sadd2(n0, sadd2(n1, n2)), bc(45.26450774985561631259f));
}

// x * lacunarity for the next octave; the product feeds adds inside simplex2_pair, so it must not be a packed
// multiply (see CAUTION above).  For lacunarity == 2 (the game's) x + x is the same value and packs.  LAC2 is a
// template parameter of the octave loops so that the choice is made once, outside them.
template <bool LAC2>
__device__ __forceinline__ float2 next_octave(float2 v, float lac) {
    return LAC2 ? add2(v, v) : smul2(v, bc(lac));
}

// functions/fbm_32.rs:18-31, two samples at once
template <bool FAST, bool LAC2, bool FMA>
__device__ __forceinline__ float2 fbm2_octaves(const NoiseTables& T, float2 x, float2 y, const WorldDev& w) {
    float2 result = simplex2_pair<FAST, FMA>(T, x, y);
    float amp = 1.0f;
    for (int o = 1; o < w.octaves; ++o) {
        x = next_octave<LAC2>(x, w.lacunarity);
        y = next_octave<LAC2>(y, w.lacunarity);
        amp = __fmul_rn(amp, w.gain);
        result = sadd2(mul2(simplex2_pair<FAST, FMA>(T, x, y), bc(amp)), result);
    }
    return result;
}
template <bool FAST, bool FMA>
__device__ __forceinline__ float2 fbm2_pair(const NoiseTables& T, float2 x, float2 y, const WorldDev& w) {
    return w.lacunarity == 2.0f ? fbm2_octaves<FAST, true, FMA>(T, x, y, w) : fbm2_octaves<FAST, false, FMA>(T, x, y, w);
}

__device__ __forceinline__ float2 abs2(float2 a) { return pk(fabsf(a.x), fabsf(a.y)); }  // andnot(-0.0, a), simd/ops/f32.rs:307-310

// The crate's other Sample32 implementations over the same simplex (noise/{ridge,turbulence,gradient}.rs):
//   ridge_2d      (functions/ridge_32.rs:18-31):      r = 1 - |s0|;  r = r + neg_mul_add(|s|, amp, 1)   (fused on AVX2)
//   turbulence_2d (functions/turbulence_32.rs:18-32): r = |s0|;      r = r + |s * amp|
//   gradient      (noise/gradient.rs:77-79):          r = s0
// KIND is a template parameter so the FBM kernel's code is untouched by the variants.
template <int KIND, bool FAST, bool FMA = true>
__device__ __forceinline__ float2 noise2_pair(const NoiseTables& T, float2 x, float2 y, const WorldDev& w) {
    if (KIND == HB_NOISE_FBM) return fbm2_pair<FAST, FMA>(T, x, y, w);
    const float2 s0 = simplex2_pair<FAST, FMA>(T, x, y);
    if (KIND == HB_NOISE_GRADIENT) return s0;
    float2 result = KIND == HB_NOISE_RIDGE ? sadd2(bc(1.0f), neg2(abs2(s0))) : abs2(s0);
    float amp = 1.0f;
    for (int o = 1; o < w.octaves; ++o) {
        x = next_octave<false>(x, w.lacunarity);
        y = next_octave<false>(y, w.lacunarity);
        amp = __fmul_rn(amp, w.gain);
        const float2 sv = simplex2_pair<FAST, FMA>(T, x, y);
        if (KIND == HB_NOISE_RIDGE) {
            const float2 a = abs2(sv);
            if constexpr (FMA) result = sadd2(result, pk(__fmaf_rn(-a.x, amp, 1.0f), __fmaf_rn(-a.y, amp, 1.0f)));
            else result = sadd2(result, pk(__fsub_rn(1.0f, __fmul_rn(a.x, amp)), __fsub_rn(1.0f, __fmul_rn(a.y, amp))));
        } else {
            result = sadd2(result, abs2(smul2(sv, bc(amp))));
        }
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

// Sample32::sample_3d of w.kind: fbm_32.rs:33-47, ridge_32.rs:33-47, turbulence_32.rs:34-48, gradient.rs:82-84
__device__ __forceinline__ float noise3(float x, float y, float z, const WorldDev& w) {
    if (w.kind == HB_NOISE_FBM) return fbm3(x, y, z, w);
    const float s0 = simplex3(x, y, z, w.seed_lo);
    if (w.kind == HB_NOISE_GRADIENT) return s0;
    const bool ridge = w.kind == HB_NOISE_RIDGE;
    float result = ridge ? __fsub_rn(1.0f, fabsf(s0)) : fabsf(s0);
    float amp = 1.0f;
    for (int o = 1; o < w.octaves; ++o) {
        x = __fmul_rn(x, w.lacunarity);
        y = __fmul_rn(y, w.lacunarity);
        z = __fmul_rn(z, w.lacunarity);
        amp = __fmul_rn(amp, w.gain);
        const float sv = simplex3(x, y, z, w.seed_lo);
        if (ridge) result = __fadd_rn(result, w.fma ? __fmaf_rn(-fabsf(sv), amp, 1.0f) : __fsub_rn(1.0f, __fmul_rn(fabsf(sv), amp)));
        else result = __fadd_rn(result, fabsf(__fmul_rn(sv, amp)));
    }
    return result;
}

}  // namespace hb
