/*
 * hb_oracle.c -- CPU restatement of herbolution's noise -> block fill -> face
 * culling -> chunk meshing path.  TEST INFRASTRUCTURE ONLY (see hb_oracle.h).
 *
 * Build: gcc -O2 -std=gnu11 -ffp-contract=off -mavx2 -mfma -fPIC -shared -pthread
 * -ffp-contract=off is REQUIRED: the reference's operator overloads are separate
 * IEEE mul/add (simd/ops/f32.rs:3-70); only the six neg_mul_add sites fuse.
 *
 * Citations are relative to /root/reference.
 */
#include "hb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* Noise: crates/simd-noise/src/functions/simplex_32.rs                       */
/* ------------------------------------------------------------------------- */

/* simplex_32.rs:6-19 */
static const float F2_32 = 0.36602540378f;
static const float G2_32 = 0.2113248654f;
static const float F3_32 = 1.0f / 3.0f;
static const float G3_32 = 1.0f / 6.0f;
static const float G33_32 = 3.0f / 6.0f - 1.0f;
/* cellular_32.rs:8,11,14 */
#define X_PRIME_32 1619
#define Y_PRIME_32 31337
#define Z_PRIME_32 6971

/* simplex_32.rs:29-46 (first 256; the table there is this sequence twice) */
static const uint8_t PERM256[256] = {
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148, 247, 120, 234,
    75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175, 74, 165, 71, 134, 139, 48,
    27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, 65, 25, 63, 161, 1, 216, 80, 73,
    209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226, 250, 124, 123, 5, 202, 38,
    147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213, 119, 248, 152, 2, 44, 154, 163, 70, 221, 153,
    101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104, 218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144,
    12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157, 184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254,
    138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180};

/* functions/ops.rs:3-11 gather_32 over PERM[512] */
static inline int32_t perm(int32_t idx) { return (int32_t)PERM256[idx & 255]; }

/* simd/ops/f32.rs:141-162 neg_mul_add(a,b,c) = c - a*b; fused on AVX2 (_mm256_fnmadd_ps) / NEON */
static inline float neg_mul_add(float a, float b, float c, int fused) {
    if (fused) return fmaf(-a, b, c);
    float p = a * b;
    return c - p;
}

/* t &= t.cmp_gte(0)  simplex_32.rs:148-150 */
static inline float mask_ge0(float t) { return (t >= 0.0f) ? t : 0.0f; }

/* functions/gradient_32.rs:13-30.  mask.blendv(a,b) = mask ? b : a (simd/overloads.rs:588-591) */
static inline void grad2(int64_t seed, int32_t hash, float *gx, float *gy) {
    int32_t h = (hash ^ (int32_t)seed) & 7;
    int mask = 4 > h;
    float x_magnitude = mask ? 1.0f : 2.0f;
    float y_magnitude = mask ? 2.0f : 1.0f;
    int h_and_1 = (h & 1) == 0;
    int h_and_2 = (h & 2) == 0;
    *gx = (mask ? h_and_1 : h_and_2) ? x_magnitude : 0.0f - x_magnitude;
    *gy = (mask ? h_and_2 : h_and_1) ? y_magnitude : 0.0f - y_magnitude;
}

/* simplex_32.rs:103-170 (value half of simplex_2d_deriv) */
float hbo_simplex2(float x, float y, int64_t seed, int fused) {
    float s = F2_32 * (x + y);
    float ips = floorf(x + s);
    float jps = floorf(y + s);

    int32_t i = (int32_t)ips;
    int32_t j = (int32_t)jps;

    float t = (float)(int32_t)((uint32_t)i + (uint32_t)j) * G2_32;

    float x0 = x - (ips - t);
    float y0 = y - (jps - t);

    int32_t i1 = (x0 >= y0) ? -1 : 0;
    int32_t j1 = (y0 > x0) ? -1 : 0;

    float x1 = (x0 + (float)i1) + G2_32;
    float y1 = (y0 + (float)j1) + G2_32;
    const float G22_32 = G2_32 * 2.0f;
    float x2 = (x0 + -1.0f) + G22_32;
    float y2 = (y0 + -1.0f) + G22_32;

    int32_t ii = i & 0xff;
    int32_t jj = j & 0xff;

    int32_t gi0 = perm(ii + perm(jj));
    int32_t gi1 = perm((ii - i1) + perm(jj - j1));
    int32_t gi2 = perm((ii - -1) + perm(jj - -1));

    float t0 = neg_mul_add(y0, y0, neg_mul_add(x0, x0, 0.5f, fused), fused);
    float t1 = neg_mul_add(y1, y1, neg_mul_add(x1, x1, 0.5f, fused), fused);
    float t2 = neg_mul_add(y2, y2, neg_mul_add(x2, x2, 0.5f, fused), fused);

    t0 = mask_ge0(t0);
    t1 = mask_ge0(t1);
    t2 = mask_ge0(t2);

    float t20 = t0 * t0;
    float t40 = t20 * t20;
    float t21 = t1 * t1;
    float t41 = t21 * t21;
    float t22 = t2 * t2;
    float t42 = t22 * t22;

    float gx, gy;
    grad2(seed, gi0, &gx, &gy);
    float a0 = gx * x0, b0 = gy * y0;
    float g0 = a0 + b0;
    float n0 = t40 * g0;
    grad2(seed, gi1, &gx, &gy);
    float a1 = gx * x1, b1 = gy * y1;
    float g1 = a1 + b1;
    float n1 = t41 * g1;
    grad2(seed, gi2, &gx, &gy);
    float a2 = gx * x2, b2 = gy * y2;
    float g2 = a2 + b2;
    float n2 = t42 * g2;

    const float scale = 45.26450774985561631259f;
    float n12 = n1 + n2;
    return (n0 + n12) * scale;
}

/* functions/fbm_32.rs:18-31 */
float hbo_fbm2(float x, float y, const hbo_world *w) {
    float lac = w->lacunarity, gain = w->gain;
    float result = hbo_simplex2(x, y, w->seed, w->fused);
    float amp = 1.0f;
    for (int o = 1; o < (w->octaves & 0xff); ++o) {
        x = x * lac;
        y = y * lac;
        amp = amp * gain;
        float sv = hbo_simplex2(x, y, w->seed, w->fused) * amp;
        result = sv + result;
    }
    return result;
}

/* functions/ridge_32.rs:18-31 */
float hbo_ridge2(float x, float y, const hbo_world *w) {
    float lac = w->lacunarity, gain = w->gain;
    float result = 1.0f - fabsf(hbo_simplex2(x, y, w->seed, w->fused));
    float amp = 1.0f;
    for (int o = 1; o < (w->octaves & 0xff); ++o) {
        x = x * lac;
        y = y * lac;
        amp = amp * gain;
        result = result + neg_mul_add(fabsf(hbo_simplex2(x, y, w->seed, w->fused)), amp, 1.0f, w->fused);
    }
    return result;
}

/* functions/turbulence_32.rs:18-32 */
float hbo_turbulence2(float x, float y, const hbo_world *w) {
    float lac = w->lacunarity, gain = w->gain;
    float result = fabsf(hbo_simplex2(x, y, w->seed, w->fused));
    float amp = 1.0f;
    for (int o = 1; o < (w->octaves & 0xff); ++o) {
        x = x * lac;
        y = y * lac;
        amp = amp * gain;
        float sv = hbo_simplex2(x, y, w->seed, w->fused) * amp;
        result = result + fabsf(sv);
    }
    return result;
}

/* Sample32::sample_2d: noise/fbm.rs, ridge.rs:100-102, turbulence.rs:100-102, gradient.rs:77-79 */
float hbo_sample2(float x, float y, const hbo_world *w) {
    switch (w->kind) {
        case HBO_NOISE_RIDGE: return hbo_ridge2(x, y, w);
        case HBO_NOISE_TURBULENCE: return hbo_turbulence2(x, y, w);
        case HBO_NOISE_GRADIENT: return hbo_simplex2(x, y, w->seed, w->fused);
        default: return hbo_fbm2(x, y, w);
    }
}

/* noise/f32.rs:69-129 get_2d_noise_helper_f32 with the transform from
 * server/src/generator/mod.rs:74,98-101: start = (cx*32 as f32, cz*32 as f32).
 * Lane coordinates are start + i (then += WIDTH): integers, exact in f32 for |coord| < 2^24,
 * so the vector width does not matter. */
void hbo_noise_tile2(const hbo_world *w, int32_t cx, int32_t cz, float out[HBO_CHUNK_AREA]) {
    float start_x = (float)cx * (float)HBO_CHUNK_LENGTH;
    float start_y = (float)cz * (float)HBO_CHUNK_LENGTH;
    float freq_x = w->freq, freq_y = w->freq;
    float y = start_y;
    int i = 0;
    for (int row = 0; row < HBO_CHUNK_LENGTH; ++row) {
        float x = start_x;
        for (int col = 0; col < HBO_CHUNK_LENGTH; ++col) {
            out[i++] = hbo_sample2(x * freq_x, y * freq_y, w);
            x = x + 1.0f;
        }
        y = y + 1.0f;
    }
}

/* ---- 3-D (crate feature, unused by the game): simplex_32.rs:197-282 ---- */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* hash3d_32.rs:21-36 + gradient_32.rs:32-47 (grad3d_dot) */
static inline float grad3d_dot(int64_t seed, int32_t i, int32_t j, int32_t k, float x, float y, float z) {
    uint32_t hash = (uint32_t)i ^ (uint32_t)(int32_t)seed;
    hash = (uint32_t)j ^ hash;
    hash = (uint32_t)k ^ hash;
    hash = ((hash * hash) * 60493u) * hash; /* wrapping i32 mul, simd/ops/i32.rs:48-77 */
    hash = (hash >> 13) ^ hash;             /* logical shift, simd/ops/i32.rs:383-406 */
    uint32_t hasha13 = hash & 13u;
    int l8 = (int32_t)hasha13 < 8;
    int l4 = (int32_t)hasha13 < 2;
    int h12_or_14 = hasha13 == 12u;
    uint32_t h1 = hash << 31;
    uint32_t h2 = (hash & 2u) << 30;
    float u = l8 ? x : y;
    float v = l4 ? y : (h12_or_14 ? x : z);
    return u2f(f2u(u) ^ h1) + u2f(f2u(v) ^ h2);
}

float hbo_simplex3(float x, float y, float z, int64_t seed) {
    float f = F3_32 * ((x + y) + z);
    float x0 = floorf(x + f);
    float y0 = floorf(y + f);
    float z0 = floorf(z + f);

    int32_t i = (int32_t)((uint32_t)(int32_t)x0 * (uint32_t)X_PRIME_32);
    int32_t j = (int32_t)((uint32_t)(int32_t)y0 * (uint32_t)Y_PRIME_32);
    int32_t k = (int32_t)((uint32_t)(int32_t)z0 * (uint32_t)Z_PRIME_32);

    float g = G3_32 * ((x0 + y0) + z0);
    x0 = x - (x0 - g);
    y0 = y - (y0 - g);
    z0 = z - (z0 - g);

    int x0_ge_y0 = x0 >= y0;
    int y0_ge_z0 = y0 >= z0;
    int x0_ge_z0 = x0 >= z0;

    int i1 = x0_ge_y0 & x0_ge_z0;
    int j1 = y0_ge_z0 & !x0_ge_y0;          /* a.and_not(b) = a & !b, simd/overloads.rs:576-586 */
    int k1 = (!y0_ge_z0) & !x0_ge_z0;

    int i2 = x0_ge_y0 | x0_ge_z0;
    int j2 = (!x0_ge_y0) | y0_ge_z0;
    int k2 = !(x0_ge_z0 & y0_ge_z0);

    float x1 = (x0 - (i1 ? 1.0f : 0.0f)) + G3_32;
    float y1 = (y0 - (j1 ? 1.0f : 0.0f)) + G3_32;
    float z1 = (z0 - (k1 ? 1.0f : 0.0f)) + G3_32;

    float x2 = (x0 - (i2 ? 1.0f : 0.0f)) + F3_32;
    float y2 = (y0 - (j2 ? 1.0f : 0.0f)) + F3_32;
    float z2 = (z0 - (k2 ? 1.0f : 0.0f)) + F3_32;

    float x3 = x0 + G33_32;
    float y3 = y0 + G33_32;
    float z3 = z0 + G33_32;

    float t0 = ((0.6f - (x0 * x0)) - (y0 * y0)) - (z0 * z0);
    float t1 = ((0.6f - (x1 * x1)) - (y1 * y1)) - (z1 * z1);
    float t2 = ((0.6f - (x2 * x2)) - (y2 * y2)) - (z2 * z2);
    float t3 = ((0.6f - (x3 * x3)) - (y3 * y3)) - (z3 * z3);

    t0 = mask_ge0(t0);
    t1 = mask_ge0(t1);
    t2 = mask_ge0(t2);
    t3 = mask_ge0(t3);

    float t20 = t0 * t0, t21 = t1 * t1, t22 = t2 * t2, t23 = t3 * t3;
    float t40 = t20 * t20, t41 = t21 * t21, t42 = t22 * t22, t43 = t23 * t23;

    float g0 = grad3d_dot(seed, i, j, k, x0, y0, z0);
    float v0 = t40 * g0;

    int32_t v1x = (int32_t)((uint32_t)i + (i1 ? (uint32_t)X_PRIME_32 : 0u));
    int32_t v1y = (int32_t)((uint32_t)j + (j1 ? (uint32_t)Y_PRIME_32 : 0u));
    int32_t v1z = (int32_t)((uint32_t)k + (k1 ? (uint32_t)Z_PRIME_32 : 0u));
    float g1 = grad3d_dot(seed, v1x, v1y, v1z, x1, y1, z1);
    float v1 = t41 * g1;

    int32_t v2x = (int32_t)((uint32_t)i + (i2 ? (uint32_t)X_PRIME_32 : 0u));
    int32_t v2y = (int32_t)((uint32_t)j + (j2 ? (uint32_t)Y_PRIME_32 : 0u));
    int32_t v2z = (int32_t)((uint32_t)k + (k2 ? (uint32_t)Z_PRIME_32 : 0u));
    float g2 = grad3d_dot(seed, v2x, v2y, v2z, x2, y2, z2);
    float v2 = t42 * g2;

    int32_t v3x = (int32_t)((uint32_t)i + (uint32_t)X_PRIME_32);
    int32_t v3y = (int32_t)((uint32_t)j + (uint32_t)Y_PRIME_32);
    int32_t v3z = (int32_t)((uint32_t)k + (uint32_t)Z_PRIME_32);
    float g3 = grad3d_dot(seed, v3x, v3y, v3z, x3, y3, z3);
    float v3 = t43 * g3;

    float p1 = v3 + v2;
    float p2 = p1 + v1;
    const float scale = 32.69587493801679f;
    return (p2 + v0) * scale;
}

/* functions/fbm_32.rs:33-47 */
float hbo_fbm3(float x, float y, float z, const hbo_world *w) {
    float lac = w->lacunarity, gain = w->gain;
    float result = hbo_simplex3(x, y, z, w->seed);
    float amp = 1.0f;
    for (int o = 1; o < (w->octaves & 0xff); ++o) {
        x = x * lac;
        y = y * lac;
        z = z * lac;
        amp = amp * gain;
        float sv = hbo_simplex3(x, y, z, w->seed) * amp;
        result = sv + result;
    }
    return result;
}

/* functions/ridge_32.rs:33-47, turbulence_32.rs:34-48; Sample32::sample_3d */
float hbo_sample3(float x, float y, float z, const hbo_world *w) {
    if (w->kind == HBO_NOISE_GRADIENT) return hbo_simplex3(x, y, z, w->seed);
    if (w->kind != HBO_NOISE_RIDGE && w->kind != HBO_NOISE_TURBULENCE) return hbo_fbm3(x, y, z, w);
    const int ridge = w->kind == HBO_NOISE_RIDGE;
    float lac = w->lacunarity, gain = w->gain;
    float s0 = fabsf(hbo_simplex3(x, y, z, w->seed));
    float result = ridge ? 1.0f - s0 : s0;
    float amp = 1.0f;
    for (int o = 1; o < (w->octaves & 0xff); ++o) {
        x = x * lac;
        y = y * lac;
        z = z * lac;
        amp = amp * gain;
        float sv = hbo_simplex3(x, y, z, w->seed);
        if (ridge) {
            result = result + neg_mul_add(fabsf(sv), amp, 1.0f, w->fused);
        } else {
            float p = sv * amp;
            result = result + fabsf(p);
        }
    }
    return result;
}

/* noise/f32.rs:131-198 */
void hbo_noise_tile3(const hbo_world *w, float sx, float sy, float sz, int nx, int ny, int nz,
                     float fx, float fy, float fz, float *out) {
    size_t i = 0;
    float z = sz;
    for (int iz = 0; iz < nz; ++iz) {
        float y = sy;
        for (int iy = 0; iy < ny; ++iy) {
            float x = sx;
            for (int ix = 0; ix < nx; ++ix) {
                out[i++] = hbo_sample3(x * fx, y * fy, z * fz, w);
                x = x + 1.0f;
            }
            y = y + 1.0f;
        }
        z = z + 1.0f;
    }
}

/* ------------------------------------------------------------------------- */
/* Generation: server/src/generator/mod.rs + server/src/chunk/mesh.rs         */
/* ------------------------------------------------------------------------- */

/* lib/src/vector.rs:1274-1276 */
static inline int lin(int x, int y, int z) { return x * 1024 + z * 32 + y; }

/* `(noise * 32.0) as i32` (generator/mod.rs:79): Rust float->int casts truncate and saturate, NaN -> 0 */
int32_t hbo_height(float noise) {
    float v = noise * (float)HBO_CHUNK_LENGTH;
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}

/* material.rs:27-61 */
void hbo_default_palette(hbo_palette *p) {
    static const float c[3][3][4] = {
        {{0.5f, 0.5f, 0.5f, 1.0f}, {0.6f, 0.6f, 0.6f, 1.0f}, {0.7f, 0.7f, 0.7f, 1.0f}},
        {{0.4f, 0.3f, 0.2f, 1.0f}, {0.5f, 0.4f, 0.3f, 1.0f}, {0.6f, 0.5f, 0.4f, 1.0f}},
        {{0.1f, 0.8f, 0.1f, 1.0f}, {0.2f, 0.9f, 0.2f, 1.0f}, {0.3f, 1.0f, 0.3f, 1.0f}},
    };
    memset(p, 0, sizeof(*p));
    p->n_materials = 3; /* insertion order stone, dirt, grass: generator/mod.rs:64-72 */
    for (int m = 0; m < 3; ++m) {
        p->mat[m].n_colors = 3;
        memcpy(p->mat[m].rgba, c[m], sizeof(c[m]));
    }
}

/* mesh.rs:22-30: Cube::new(None) everywhere, flags 0 */
void hbo_mesh_init(hbo_cube_mesh *m, int32_t cx, int32_t cy, int32_t cz) {
    memset(m, 0, sizeof(*m));
    m->pos[0] = cx; m->pos[1] = cy; m->pos[2] = cz;
}

/* cube.rs:35-49,69-72 */
static inline uint32_t flags_faces(uint32_t v) { return (v & 1u) == 0 ? 0u : ((v >> 1) & 0x3Fu); }
static inline uint32_t flags_set_opaque(uint32_t faces) { return (faces << 1) | 1u; }
/* remove_faces(f): set_opaque(faces() - f)  (Into<CubeFaces> is the correct 1<<f mapping, face.rs:181-192) */
static inline void cube_remove_face(hbo_cube *c, int f) {
    c->flags = flags_set_opaque(flags_faces(c->flags) & ~(1u << f) & 0x3Fu);
}
static inline void cube_insert_face(hbo_cube *c, int f) {
    c->flags = flags_set_opaque(flags_faces(c->flags) | (1u << f));
}

/* every Material has cullable_faces = all() (material.rs:31,43,55) => non-empty iff Some */
static inline int cullable_nonempty(uint16_t material) { return material != 0; }

/* face.rs:47-58 */
static const int NORMAL[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
static inline int inverse_face(int f) { return f ^ 1; } /* face.rs:71-81 */

/* updated_positions.push(vec3u5::new(x, y, z)) */
static inline void push_updated(hbo_cube_mesh *m, int x, int y, int z) {
    if (m->upd && m->n_updated < m->upd_cap)
        m->upd[m->n_updated] = (uint16_t)(1u | ((unsigned)x << 11) | ((unsigned)y << 6) | ((unsigned)z << 1));
    m->n_updated++;
}

/* mesh.rs:209-232 */
static void remove_neighboring_faces(hbo_cube_mesh *m, uint32_t faces, int x, int y, int z) {
    for (int f = 0; f < 6; ++f) { /* FaceIter order, face.rs:313-329 */
        if (!(faces & (1u << f))) continue;
        int nx = x + NORMAL[f][0], ny = y + NORMAL[f][1], nz = z + NORMAL[f][2];
        if (nx < 0 || ny < 0 || nz < 0) continue;       /* try_cast::<u8>() fails */
        if (nx >= 32 || ny >= 32 || nz >= 32) continue; /* vec3u5::try_from fails */
        int index = lin(nx, ny, nz);
        push_updated(m, nx, ny, nz);
        if (cullable_nonempty(m->data[index].material)) cube_remove_face(&m->data[lin(x, y, z)], f);
        cube_remove_face(&m->data[index], inverse_face(f));
    }
}

/* mesh.rs:234-251.  NB: iterates faces.not() as written in the reference. */
static void add_neighboring_faces(hbo_cube_mesh *m, uint32_t faces, int x, int y, int z) {
    uint32_t it = ~faces & 0x3Fu;
    for (int f = 0; f < 6; ++f) {
        if (!(it & (1u << f))) continue;
        int nx = x + NORMAL[f][0], ny = y + NORMAL[f][1], nz = z + NORMAL[f][2];
        if (nx < 0 || ny < 0 || nz < 0 || nx >= 32 || ny >= 32 || nz >= 32) continue;
        push_updated(m, nx, ny, nz);
        cube_insert_face(&m->data[lin(nx, ny, nz)], inverse_face(f));
    }
}

/* mesh.rs:125-186 (exposed_faces bookkeeping at :159-183 is omitted: it only feeds
 * set_rendered(), and with the discriminant-as-mask bug in face.rs:163-173 bits 3..5 of
 * exposed_faces can never clear, so it is constant-true) */
void hbo_mesh_set(hbo_cube_mesh *m, int x, int y, int z, uint16_t new_material) {
    int i = lin(x, y, z);
    uint16_t old_material = m->data[i].material;
    if (new_material == old_material) return;
    m->data[i].material = new_material;

    if (!cullable_nonempty(old_material) && cullable_nonempty(new_material)) {
        uint32_t missing = ~flags_faces(m->data[i].flags) & 0x3Fu;
        m->data[i].flags = flags_set_opaque(0x3Fu);
        remove_neighboring_faces(m, missing, x, y, z);
    }
    if (cullable_nonempty(old_material) && !cullable_nonempty(new_material)) {
        uint32_t present = flags_faces(m->data[i].flags);
        m->data[i].flags = flags_set_opaque(0u);
        add_neighboring_faces(m, present, x, y, z);
    }
    push_updated(m, x, y, z);
}

/* generator/mod.rs:63-95 through CubeMesh::set */
void hbo_generate_literal(const hbo_world *w, hbo_cube_mesh *m) {
    const uint16_t stone = 1, dirt = 2, grass = 3; /* generator/mod.rs:64-72, material.rs:162-173 */
    float noise[HBO_CHUNK_AREA];
    hbo_noise_tile2(w, m->pos[0], m->pos[2], noise);
    for (int x = 0; x < 32; ++x) {
        for (int z = 0; z < 32; ++z) {
            int32_t h = hbo_height(noise[x + z * 32]);
            for (int chunk_y = 0; chunk_y < 32; ++chunk_y) {
                int64_t y = (int64_t)m->pos[1] * 32 + chunk_y;
                if (y < (int64_t)h - 6) hbo_mesh_set(m, x, chunk_y, z, stone);
                else if (y < (int64_t)h - 1) hbo_mesh_set(m, x, chunk_y, z, dirt);
                else if (y < (int64_t)h) hbo_mesh_set(m, x, chunk_y, z, grass);
            }
        }
    }
}

void hbo_generate_ids(const hbo_world *w, int32_t cx, int32_t cy, int32_t cz, uint16_t *ids) {
    float noise[HBO_CHUNK_AREA];
    hbo_noise_tile2(w, cx, cz, noise);
    for (int x = 0; x < 32; ++x)
        for (int z = 0; z < 32; ++z) {
            int32_t h = hbo_height(noise[x + z * 32]);
            for (int yl = 0; yl < 32; ++yl) {
                int64_t y = (int64_t)cy * 32 + yl;
                uint16_t id = 0;
                if (y < (int64_t)h - 6) id = 1;
                else if (y < (int64_t)h - 1) id = 2;
                else if (y < (int64_t)h) id = 3;
                ids[lin(x, yl, z)] = id;
            }
        }
}

/* Closed form of the state CubeMesh::set leaves behind: a solid voxel keeps
 * face f iff the in-chunk neighbour in direction f is air or f leaves the chunk; an air voxel
 * has flags=1 iff at least one in-chunk neighbour is solid, else 0. */
void hbo_cubes_from_ids(const uint16_t *ids, hbo_cube *out) {
    for (int x = 0; x < 32; ++x)
        for (int z = 0; z < 32; ++z)
            for (int y = 0; y < 32; ++y) {
                int i = lin(x, y, z);
                uint32_t faces = 0;
                int any_solid_nbr = 0;
                for (int f = 0; f < 6; ++f) {
                    int nx = x + NORMAL[f][0], ny = y + NORMAL[f][1], nz = z + NORMAL[f][2];
                    if (nx < 0 || ny < 0 || nz < 0 || nx >= 32 || ny >= 32 || nz >= 32) {
                        faces |= 1u << f;
                        continue;
                    }
                    if (ids[lin(nx, ny, nz)] == 0) faces |= 1u << f;
                    else any_solid_nbr = 1;
                }
                out[i].material = ids[i];
                out[i].pad = 0;
                if (ids[i] != 0) out[i].flags = flags_set_opaque(faces);
                else out[i].flags = any_solid_nbr ? 1u : 0u;
            }
}

/* mesh.rs:254-264 boundary(face): the fixed coordinate of the slab */
static inline int boundary_fixed(int f) { return (f & 1) ? 0 : 31; }

/* mesh.rs:76-123 */
void hbo_cull_shared_faces(hbo_cube_mesh *a, hbo_cube_mesh *b) {
    int d[3] = {b->pos[0] - a->pos[0], b->pos[1] - a->pos[1], b->pos[2] - a->pos[2]};
    int shared = -1;
    for (int f = 0; f < 6; ++f)
        if (d[0] == NORMAL[f][0] && d[1] == NORMAL[f][1] && d[2] == NORMAL[f][2]) shared = f;
    if (shared < 0) return; /* from_normal -> None */
    int other_shared = inverse_face(shared);
    int axis = shared >> 1; /* 0:x 1:y 2:z */
    int fa = boundary_fixed(shared), fb = boundary_fixed(other_shared);
    for (int u = 0; u < 32; ++u)
        for (int v = 0; v < 32; ++v) {
            int pa[3], pb[3];
            int k = 0;
            for (int ax = 0; ax < 3; ++ax) {
                if (ax == axis) { pa[ax] = fa; pb[ax] = fb; }
                else { int c = (k++ == 0) ? u : v; pa[ax] = c; pb[ax] = c; }
            }
            hbo_cube *ca = &a->data[lin(pa[0], pa[1], pa[2])];
            hbo_cube *cb = &b->data[lin(pb[0], pb[1], pb[2])];
            if (cullable_nonempty(ca->material) && cullable_nonempty(cb->material)) {
                cube_remove_face(ca, shared);
                cube_remove_face(cb, other_shared);
                push_updated(a, pa[0], pa[1], pa[2]);
                push_updated(b, pb[0], pb[1], pb[2]);
            }
        }
}

/* ------------------------------------------------------------------------- */
/* Meshing: client/src/world/chunk.rs                                         */
/* ------------------------------------------------------------------------- */

/* std::hash::DefaultHasher == SipHash-1-3 with k0=k1=0; #[derive(Hash)] on ChunkPt(Vec3<i32>)
 * (lib/src/point.rs:34-35, lib/src/vector.rs:74) feeds x, y, z as 4 native-endian bytes each. */
static inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
#define SIPROUND                                                   \
    do {                                                           \
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32); \
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                   \
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                   \
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32); \
    } while (0)

uint64_t hbo_siphash13_chunk(int32_t x, int32_t y, int32_t z) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    uint64_t m0 = (uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32);
    v3 ^= m0; SIPROUND; v0 ^= m0;
    uint64_t b = ((uint64_t)12 << 56) | (uint64_t)(uint32_t)z;
    v3 ^= b; SIPROUND; v0 ^= b;
    v2 ^= 0xff;
    SIPROUND; SIPROUND; SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

/* fastrand 2.3.0 (Cargo.lock:712-715; NOT vendored -> restated, unpinned):
 * Rng::with_seed(s) = Rng(s); gen_u64: s += 0x2d358dccaa6c78a5; t = s as u128 * (s ^ 0x8bb84b93962eacc9);
 * (t as u64) ^ (t >> 64) as u64.  Rng::f32: from_bits(0x3F800000 | (gen_u32() >> 9)) - 1.0 with
 * gen_u32 = gen_u64 as u32. */
uint64_t hbo_wyrand_next(uint64_t *state) {
    uint64_t s = *state + 0x2d358dccaa6c78a5ULL;
    *state = s;
    __uint128_t t = (__uint128_t)s * (__uint128_t)(s ^ 0x8bb84b93962eacc9ULL);
    return (uint64_t)t ^ (uint64_t)(t >> 64);
}
float hbo_wyrand_f32(uint64_t *state) {
    uint32_t u = (uint32_t)hbo_wyrand_next(state);
    return u2f(0x3F800000u | (u >> 9)) - 1.0f;
}

/* CubeFace::rotation (face.rs:60-69) -> Quat::from_euler (rotation.rs:77-88) -> to_axes (:109-116)
 * -> * Mat3::from(Vec3::ONE) (mat3.rs:63-90) -> columns extended with 0.0 (vertex.rs:53-62).
 * sin/cos come from the platform libm, as f32::sin_cos does on linux. */
void hbo_face_matrix(int face, float out12[12]) {
    const float FRAC_PI_2 = 1.57079632679489661923132169163975144f;
    const float PI = 3.14159265358979323846264338327950288f;
    float yaw = 0.0f, pitch = 0.0f, roll = 0.0f; /* Euler::new(yaw, pitch, roll) */
    switch (face) {
        case HBO_EAST: pitch = FRAC_PI_2; break;
        case HBO_WEST: pitch = -FRAC_PI_2; break;
        case HBO_UP: yaw = -FRAC_PI_2; break;
        case HBO_DOWN: yaw = FRAC_PI_2; break;
        case HBO_NORTH: break;
        case HBO_SOUTH: pitch = PI; break;
    }
    float sx = sinf(yaw / 2.0f), cx = cosf(yaw / 2.0f);
    float sy = sinf(pitch / 2.0f), cy = cosf(pitch / 2.0f);
    float sz = sinf(roll / 2.0f), cz = cosf(roll / 2.0f);
    float qx = sx * cy * cz - cx * sy * sz;
    float qy = cx * sy * cz + sx * cy * sz;
    float qz = cx * cy * sz - sx * sy * cz;
    float qw = cx * cy * cz + sx * sy * sz;
    float x = qx, y = qy, z = qz, w = qw;
    float ax[3][3] = {
        {1.f - 2.f * (y * y + z * z), 2.f * (x * y + z * w), 2.f * (x * z - y * w)},
        {2.f * (x * y - z * w), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + x * w)},
        {2.f * (x * z + y * w), 2.f * (y * z - x * w), 1.f - 2.f * (x * x + y * y)},
    };
    /* rotation_matrix * Mat3::from(ONE): column c = ax.x*s.x + ax.y*s.y + ax.z*s.z with s = e_c */
    static const float S[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            float p0 = ax[0][r] * S[c][0];
            float p1 = ax[1][r] * S[c][1];
            float p2 = ax[2][r] * S[c][2];
            out12[c * 4 + r] = (p0 + p1) + p2;
        }
    out12[3] = out12[7] = out12[11] = 0.0f;
}

/* chunk.rs:271-275: (Vec4<u8>.cast() / 3.0) is f64 by literal fallback, then cast::<f32>() */
float hbo_ao_value(int k) {
    double amount = (double)k / 3.0;
    float a = (float)amount;
    float factor = -a + 1.0f;
    return factor > 0.15f ? factor : 0.15f;
}

/* face.rs:84-93 orthonormal_basis -> (u, v, n) */
static const int BASIS[6][3][3] = {
    {{0, 0, -1}, {0, 1, 0}, {1, 0, 0}},   /* East  (-Z, Y, X) */
    {{0, 0, 1}, {0, 1, 0}, {-1, 0, 0}},   /* West  ( Z, Y,-X) */
    {{1, 0, 0}, {0, 0, -1}, {0, 1, 0}},   /* Up    ( X,-Z, Y) */
    {{1, 0, 0}, {0, 0, 1}, {0, -1, 0}},   /* Down  ( X, Z,-Y) */
    {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}},    /* North ( X, Y, Z) */
    {{-1, 0, 0}, {0, 1, 0}, {0, 0, -1}},  /* South (-X, Y,-Z) */
};

static inline int div_euclid32(int a) { return (a >= 0) ? a / 32 : -((-a + 31) / 32); }
static inline int rem_euclid32(int a) { int r = a % 32; return r < 0 ? r + 32 : r; }

/* chunk.rs:225-249; `solid(shell_entry, L)` abstracts cubes[..].material.is_some() */
typedef int (*solid_fn)(const void *chunk, int L);
static int is_cube_present(const void *const shell[27], solid_fn solid, int px, int py, int pz) {
    int ox = div_euclid32(px), oy = div_euclid32(py), oz = div_euclid32(pz);
    if (ox < -1 || ox > 1 || oy < -1 || oy > 1 || oz < -1 || oz > 1) return 0;
    int index = (ox + 1) * 9 + (oy + 1) * 3 + (oz + 1); /* Vec3::linearize(3), vector.rs:1010-1018 */
    const void *target = shell[index];
    if (!target) return 0;
    return solid(target, lin(rem_euclid32(px), rem_euclid32(py), rem_euclid32(pz)));
}

static inline int vertex_ao(int s1, int s2, int c) { return (s1 && s2) ? 3 : s1 + s2 + c; } /* chunk.rs:251-253 */

/* chunk.rs:255-276 */
static void facial_ao(const void *const shell[27], solid_fn solid, int face, int px, int py, int pz, float ao[4]) {
    const int *u = BASIS[face][0], *v = BASIS[face][1], *n = BASIS[face][2];
#define PROBE(a, b) is_cube_present(shell, solid, px + n[0] + (a) * u[0] + (b) * v[0], \
                                    py + n[1] + (a) * u[1] + (b) * v[1], pz + n[2] + (a) * u[2] + (b) * v[2])
    int tl = PROBE(-1, -1), tc = PROBE(0, -1), tr = PROBE(1, -1);
    int ml = PROBE(-1, 0), mr = PROBE(1, 0);
    int bl = PROBE(-1, 1), bc = PROBE(0, 1), br = PROBE(1, 1);
#undef PROBE
    ao[0] = hbo_ao_value(vertex_ao(ml, tc, tl));
    ao[1] = hbo_ao_value(vertex_ao(ml, bc, bl));
    ao[2] = hbo_ao_value(vertex_ao(mr, tc, tr));
    ao[3] = hbo_ao_value(vertex_ao(mr, bc, br));
}

/* material.rs:67-71 */
static const float *get_color(const hbo_palette *pal, uint16_t material, float p) {
    const hbo_material *m = &pal->mat[material - 1];
    int len = m->n_colors;
    float f = (float)(len > 0 ? len - 1 : 0) * p;
    int idx = (int)f;
    return m->rgba[idx];
}

static int solid_cube(const void *chunk, int L) { return ((const hbo_cube *)chunk)[L].material != 0; }
static int solid_id(const void *chunk, int L) { return ((const uint16_t *)chunk)[L] != 0; }

static void push_instance(hbo_instance *o, const float mat12[12], const int32_t cpos[3], int x, int y, int z,
                          const float *rgba, const float ao[4]) {
    memcpy(o->model, mat12, 12 * sizeof(float));
    o->model[12] = (float)(cpos[0] * 32 + x); /* (chunk_position + position).cast::<f32>() chunk.rs:204 */
    o->model[13] = (float)(cpos[1] * 32 + y);
    o->model[14] = (float)(cpos[2] * 32 + z);
    o->model[15] = 1.0f;
    memcpy(o->rgba, rgba, 4 * sizeof(float));
    o->light = 0;
    memcpy(o->ao, ao, 4 * sizeof(float));
}

/* chunk.rs:176-223.  mode 0: faces from cube.flags (literal); mode 1: faces recomputed from ids */
static size_t mesh_impl(const void *const shell[27], int ids_mode, const int32_t cpos[3],
                        const hbo_palette *pal, hbo_instance *out, size_t cap) {
    float mats[6][12];
    for (int f = 0; f < 6; ++f) hbo_face_matrix(f, mats[f]);
    solid_fn solid = ids_mode ? solid_id : solid_cube;
    uint64_t rng = hbo_siphash13_chunk(cpos[0], cpos[1], cpos[2]); /* Rng::with_seed(hasher.finish()) */
    size_t n = 0;
    const void *center = shell[13];
    for (int x = 0; x < 32; ++x)
        for (int z = 0; z < 32; ++z)
            for (int y = 0; y < 32; ++y) {
                int L = lin(x, y, z);
                float perms[6];
                for (int f = 0; f < 6; ++f) perms[f] = hbo_wyrand_f32(&rng); /* PerFace::mapped, face.rs:430-442 */
                uint16_t material;
                uint32_t faces;
                if (ids_mode) {
                    material = ((const uint16_t *)center)[L];
                    faces = 0;
                    if (material)
                        for (int f = 0; f < 6; ++f)
                            if (!is_cube_present(shell, solid, x + NORMAL[f][0], y + NORMAL[f][1], z + NORMAL[f][2]))
                                faces |= 1u << f;
                } else {
                    const hbo_cube *c = &((const hbo_cube *)center)[L];
                    material = c->material;
                    faces = flags_faces(c->flags);
                }
                if (!material) continue;
                if (material > pal->n_materials) continue; /* get_by_id -> None => `continue` chunk.rs:190-192 */
                for (int f = 0; f < 6; ++f) {
                    if (!(faces & (1u << f))) continue;
                    float ao[4];
                    facial_ao(shell, solid, f, x, y, z, ao);
                    if (out && n < cap) push_instance(&out[n], mats[f], cpos, x, y, z, get_color(pal, material, perms[f]), ao);
                    n++;
                }
            }
    return n;
}

size_t hbo_generate_mesh(const hbo_cube *const shell[27], const int32_t center_pos[3],
                         const hbo_palette *pal, hbo_instance *out, size_t cap) {
    return mesh_impl((const void *const *)shell, 0, center_pos, pal, out, cap);
}

size_t hbo_mesh_ids(const uint16_t *const shell[27], const int32_t center_pos[3],
                    const hbo_palette *pal, hbo_instance *out, size_t cap) {
    return mesh_impl((const void *const *)shell, 1, center_pos, pal, out, cap);
}

/* ------------------------------------------------------------------------- */
/* Region pipeline: the CPU baseline driver                                   */
/* ------------------------------------------------------------------------- */

typedef struct region_job {
    const hbo_world *w;
    const hbo_palette *pal;
    int32_t cx0, cz0, ncx, ncz, cy0, ncy;
    int literal, threads, phase;
    hbo_cube_mesh **meshes; /* literal */
    uint16_t *ids;          /* closed form: n * 32768 */
    uint32_t *counts;
    hbo_instance *out;
    const uint64_t *offsets;
    size_t cap;
    volatile int64_t next;
    int64_t n;
    pthread_mutex_t mu;
} region_job;

static inline int64_t chunk_index(const region_job *j, int ix, int iz, int iy) {
    if (ix < 0 || iz < 0 || iy < 0 || ix >= j->ncx || iz >= j->ncz || iy >= j->ncy) return -1;
    return ((int64_t)ix * j->ncz + iz) * j->ncy + iy;
}

static int64_t take(region_job *j) {
    return __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
}

static void *region_worker(void *arg) {
    region_job *j = (region_job *)arg;
    for (;;) {
        int64_t c = take(j);
        if (c >= j->n) break;
        int iy = (int)(c % j->ncy);
        int iz = (int)((c / j->ncy) % j->ncz);
        int ix = (int)(c / ((int64_t)j->ncy * j->ncz));
        int32_t pos[3] = {j->cx0 + ix, j->cy0 + iy, j->cz0 + iz};
        if (j->phase == 0) { /* ChunkGenerator::request task, generator/mod.rs:36-43 */
            if (j->literal) {
                hbo_cube_mesh *m = (hbo_cube_mesh *)malloc(sizeof(hbo_cube_mesh)); /* Box::new, mesh.rs:25 */
                hbo_mesh_init(m, pos[0], pos[1], pos[2]);
                hbo_generate_literal(j->w, m);
                j->meshes[c] = m;
            } else {
                hbo_generate_ids(j->w, pos[0], pos[1], pos[2], j->ids + (size_t)c * HBO_CHUNK_VOLUME);
            }
        } else if (j->phase == 1) { /* load_provided culls, map.rs:213-224: each adjacent pair once.
                                       Work item = chunk; it culls against its +x,+y,+z neighbours.  Pairs
                                       touching the same chunk are serialised by the parity sweep below. */
        } else { /* client remesh task, chunk.rs:66-75 */
            const void *shell[27];
            int k = 0;
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dz = -1; dz <= 1; ++dz) {
                        int64_t o = chunk_index(j, ix + dx, iz + dz, iy + dy);
                        if (o < 0) shell[k++] = NULL;
                        else shell[k++] = j->literal ? (const void *)j->meshes[o]->data
                                                     : (const void *)(j->ids + (size_t)o * HBO_CHUNK_VOLUME);
                    }
            hbo_instance *dst = NULL;
            size_t cap = 0;
            if (j->out && j->offsets) {
                dst = j->out + j->offsets[c];
                cap = j->offsets[c] < j->cap ? j->cap - j->offsets[c] : 0;
            }
            size_t q = mesh_impl(shell, !j->literal, pos, j->pal, dst, cap);
            j->counts[c] = (uint32_t)q;
        }
    }
    return NULL;
}

/* cull phase worker: for a fixed axis and parity, pairs (c, c+e_axis) with coordinate parity p are disjoint */
typedef struct cull_job {
    region_job *j;
    int axis, parity;
    volatile int64_t next;
} cull_job;

static void *cull_worker(void *arg) {
    cull_job *cj = (cull_job *)arg;
    region_job *j = cj->j;
    for (;;) {
        int64_t c = __atomic_fetch_add(&cj->next, 1, __ATOMIC_RELAXED);
        if (c >= j->n) break;
        int iy = (int)(c % j->ncy);
        int iz = (int)((c / j->ncy) % j->ncz);
        int ix = (int)(c / ((int64_t)j->ncy * j->ncz));
        int coord = cj->axis == 0 ? ix : (cj->axis == 1 ? iy : iz);
        if ((coord & 1) != cj->parity) continue;
        int64_t o = chunk_index(j, ix + (cj->axis == 0), iz + (cj->axis == 2), iy + (cj->axis == 1));
        if (o < 0) continue;
        hbo_cull_shared_faces(j->meshes[c], j->meshes[o]);
    }
    return NULL;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void run_threads(void *(*fn)(void *), void *arg, int threads) {
    if (threads <= 1) { fn(arg); return; }
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; ++i) pthread_create(&t[i], NULL, fn, arg);
    for (int i = 0; i < threads; ++i) pthread_join(t[i], NULL);
    free(t);
}

uint64_t hbo_region_pipeline(const hbo_world *w, const hbo_palette *pal,
                             int32_t cx0, int32_t cz0, int32_t ncx, int32_t ncz,
                             int32_t cy0, int32_t ncy, int literal, int threads,
                             uint32_t *counts, hbo_instance *out, size_t cap,
                             uint16_t *ids_out, double secs[3]) {
    region_job j;
    memset(&j, 0, sizeof(j));
    j.w = w; j.pal = pal;
    j.cx0 = cx0; j.cz0 = cz0; j.ncx = ncx; j.ncz = ncz; j.cy0 = cy0; j.ncy = ncy;
    j.literal = literal; j.threads = threads;
    j.n = (int64_t)ncx * ncz * ncy;
    uint32_t *own_counts = NULL;
    if (!counts) counts = own_counts = (uint32_t *)calloc((size_t)j.n, sizeof(uint32_t));
    j.counts = counts;
    if (literal) j.meshes = (hbo_cube_mesh **)calloc((size_t)j.n, sizeof(hbo_cube_mesh *));
    else j.ids = ids_out ? ids_out : (uint16_t *)malloc((size_t)j.n * HBO_CHUNK_VOLUME * sizeof(uint16_t));

    double t0 = now_s();
    j.phase = 0; j.next = 0;
    run_threads(region_worker, &j, threads);
    double t1 = now_s();
    if (literal) {
        for (int axis = 0; axis < 3; ++axis)
            for (int parity = 0; parity < 2; ++parity) {
                cull_job cj = {&j, axis, parity, 0};
                run_threads(cull_worker, &cj, threads);
            }
    }
    double t2 = now_s();
    double cull_s = t2 - t1;
    /* The reference meshes each chunk into its own Vec (chunk.rs:61-75).  To concatenate into one
       caller buffer we need offsets first: a counting pass (excluded from secs), then the writing pass. */
    j.phase = 2; j.next = 0; j.out = NULL; j.offsets = NULL;
    uint64_t total = 0;
    uint64_t *offsets = NULL;
    if (out) {
        run_threads(region_worker, &j, threads);
        offsets = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)j.n);
        for (int64_t c = 0; c < j.n; ++c) { offsets[c] = total; total += counts[c]; }
        j.out = out; j.offsets = offsets; j.cap = cap; j.next = 0;
    }
    double t3 = now_s();
    run_threads(region_worker, &j, threads);
    double t4 = now_s();
    if (!out) for (int64_t c = 0; c < j.n; ++c) total += counts[c];
    if (secs) { secs[0] = t1 - t0; secs[1] = cull_s; secs[2] = t4 - t3; }

    if (literal) {
        if (ids_out)
            for (int64_t c = 0; c < j.n; ++c)
                for (int L = 0; L < HBO_CHUNK_VOLUME; ++L)
                    ids_out[(size_t)c * HBO_CHUNK_VOLUME + L] = j.meshes[c]->data[L].material;
        for (int64_t c = 0; c < j.n; ++c) free(j.meshes[c]);
        free(j.meshes);
    } else if (!ids_out) {
        free(j.ids);
    }
    free(offsets);
    free(own_counts);
    return total;
}

/* ------------------------------------------------------------------------- */
/* Server ChunkMap, CubeUpdate wire format, client apply  */
/* ------------------------------------------------------------------------- */

struct hbo_map {
    hbo_world world;
    hbo_cube_mesh **chunks; /* by handle; NULL after unload */
    int n, cap;
};

hbo_map *hbo_map_create(const hbo_world *w) {
    hbo_map *m = (hbo_map *)calloc(1, sizeof(*m));
    if (m) m->world = *w;
    return m;
}

static void map_free_chunk(hbo_cube_mesh *c) {
    if (!c) return;
    free(c->upd);
    free(c);
}

void hbo_map_destroy(hbo_map *m) {
    if (!m) return;
    for (int i = 0; i < m->n; ++i) map_free_chunk(m->chunks[i]);
    free(m->chunks);
    free(m);
}

int hbo_map_find(const hbo_map *m, int32_t cx, int32_t cy, int32_t cz) {
    for (int i = 0; i < m->n; ++i) {
        const hbo_cube_mesh *c = m->chunks[i];
        if (c && c->pos[0] == cx && c->pos[1] == cy && c->pos[2] == cz) return i;
    }
    return -1;
}

int hbo_map_len(const hbo_map *m) { return m->n; }

/* the log must hold every push a chunk can see between two drains: generation pushes at most
 * 7 per voxel, six culls 1024 each; edits add a handful.  Grown on demand at drain time is not
 * possible (pushes happen inside set()), so it is sized generously and pushes past it are an error
 * the tests check for via hbo_map_pending() > stored. */
#define HBO_UPD_CAP ((uint64_t)HBO_CHUNK_VOLUME * 8u)

int hbo_map_load(hbo_map *m, int32_t cx, int32_t cy, int32_t cz) {
    int h = hbo_map_find(m, cx, cy, cz);
    if (h >= 0) return h; /* map.rs:69-75 */
    hbo_cube_mesh *c = (hbo_cube_mesh *)malloc(sizeof(*c));
    if (!c) return -1;
    hbo_mesh_init(c, cx, cy, cz);
    c->upd = (uint16_t *)malloc(sizeof(uint16_t) * HBO_UPD_CAP);
    c->upd_cap = c->upd ? HBO_UPD_CAP : 0;
    hbo_generate_literal(&m->world, c);
    /* map.rs:203-228: for face in CubeFace::values(): neighbour loaded -> mesh.cull_shared_faces(adj) */
    for (int f = 0; f < 6; ++f) {
        int o = hbo_map_find(m, cx + NORMAL[f][0], cy + NORMAL[f][1], cz + NORMAL[f][2]);
        if (o >= 0) hbo_cull_shared_faces(c, m->chunks[o]);
    }
    if (m->n == m->cap) {
        int ncap = m->cap ? m->cap * 2 : 64;
        hbo_cube_mesh **nc = (hbo_cube_mesh **)realloc(m->chunks, sizeof(*nc) * (size_t)ncap);
        if (!nc) { map_free_chunk(c); return -1; }
        m->chunks = nc;
        m->cap = ncap;
    }
    m->chunks[m->n] = c;
    return m->n++;
}

void hbo_map_unload(hbo_map *m, int32_t cx, int32_t cy, int32_t cz) {
    int h = hbo_map_find(m, cx, cy, cz);
    if (h < 0) return;
    map_free_chunk(m->chunks[h]);
    m->chunks[h] = NULL;
}

/* map.rs:81-151 */
void hbo_map_set_cube(hbo_map *m, int32_t wx, int32_t wy, int32_t wz, uint16_t material) {
    /* lib/src/point.rs:25-32: chunk = p >> 5 (arithmetic), local = p & 31 */
    const int32_t c[3] = {wx >> 5, wy >> 5, wz >> 5};
    const int l[3] = {wx & 31, wy & 31, wz & 31};
    /* edges, in the reference's order: (-x, EAST), (+x, WEST), (-y, UP), (+y, DOWN), (-z, NORTH), (+z, SOUTH) */
    static const int off[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    static const int face_of[6] = {0, 1, 2, 3, 4, 5};
    for (int e = 0; e < 6; ++e) {
        const int axis = e >> 1, neg = (e & 1) == 0;
        if (!(neg ? l[axis] == 0 : l[axis] == 31)) continue;
        int o = hbo_map_find(m, c[0] + off[e][0], c[1] + off[e][1], c[2] + off[e][2]);
        if (o < 0) continue;
        int nl[3] = {l[0], l[1], l[2]};
        nl[axis] = neg ? 31 : 0;
        hbo_cube_mesh *nm = m->chunks[o];
        cube_insert_face(&nm->data[lin(nl[0], nl[1], nl[2])], face_of[e]); /* unconditional, map.rs:113 */
        push_updated(nm, nl[0], nl[1], nl[2]);
    }
    int h = hbo_map_find(m, c[0], c[1], c[2]);
    if (h < 0) return;
    /* palette.get_id_by_key: stone/dirt/grass are in every generated chunk's palette (ids 1..3) */
    hbo_mesh_set(m->chunks[h], l[0], l[1], l[2], material);
}

const hbo_cube *hbo_map_cubes(const hbo_map *m, int handle) {
    if (handle < 0 || handle >= m->n || !m->chunks[handle]) return NULL;
    return m->chunks[handle]->data;
}

void hbo_map_position(const hbo_map *m, int handle, int32_t pos[3]) {
    pos[0] = pos[1] = pos[2] = 0;
    if (handle < 0 || handle >= m->n || !m->chunks[handle]) return;
    memcpy(pos, m->chunks[handle]->pos, sizeof(int32_t) * 3);
}

uint64_t hbo_map_pending(const hbo_map *m, int handle) {
    if (handle < 0 || handle >= m->n || !m->chunks[handle]) return 0;
    return m->chunks[handle]->n_updated;
}

/* server/src/chunk/mod.rs:35-67 */
uint64_t hbo_map_drain(hbo_map *m, int handle, hbo_chunk_cube *out, uint64_t cap) {
    if (handle < 0 || handle >= m->n || !m->chunks[handle]) return 0;
    hbo_cube_mesh *c = m->chunks[handle];
    const uint64_t n = c->n_updated < c->upd_cap ? c->n_updated : c->upd_cap;
    for (uint64_t i = 0; i < n && i < cap; ++i) {
        const unsigned p = c->upd[i];
        const int x = (p >> 11) & 31, y = (p >> 6) & 31, z = (p >> 1) & 31;
        out[i].pos = (uint16_t)p;
        out[i].pad = 0;
        out[i].cube = c->data[lin(x, y, z)];
    }
    const uint64_t total = c->n_updated;
    c->n_updated = 0;
    return total;
}

/* client/src/world/chunk.rs:132-149 */
void hbo_apply_updates(hbo_cube *client_cubes, const hbo_chunk_cube *entries, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        const unsigned p = entries[i].pos;
        const int x = (p >> 11) & 31, y = (p >> 6) & 31, z = (p >> 1) & 31;
        client_cubes[lin(x, y, z)] = entries[i].cube;
    }
}

/* ------------------------------------------------------------------------- */
/* On-disk cube codec: server/src/chunk/codec.rs              */
/* ------------------------------------------------------------------------- */

/* PaletteMaterialId::to_u16 (material.rs:224-226) is the 0-based palette index, and None is written as 0
 * (codec.rs:85-90,96-101): air and the first material share the wire value 0.  `raw` is the in-memory
 * NonZeroU16 niche value (0 = None, k = index k-1). */
static inline uint16_t wire_of_raw(uint16_t raw) { return raw ? (uint16_t)(raw - 1) : 0; }

/* encode_cubes (codec.rs:73-101): [u8 count][u16 LE id] records; a record is closed when the material changes
 * or the count reaches 255; the state starts as (None, 0), so a first voxel that is not None closes an empty
 * record first.  Returns the number of bytes the stream has; writes at most cap. */
size_t hbo_rle_encode(const uint16_t *raw_ids, size_t n, uint8_t *out, size_t cap) {
    size_t w = 0;
    uint8_t count = 0;
    uint16_t current = 0; /* None */
#define HBO_PUSH_REC()                                         \
    do {                                                       \
        const uint16_t v = wire_of_raw(current);               \
        if (w + 0 < cap) out[w + 0] = count;                   \
        if (w + 1 < cap) out[w + 1] = (uint8_t)(v & 0xFF);     \
        if (w + 2 < cap) out[w + 2] = (uint8_t)(v >> 8);       \
        w += 3;                                                \
    } while (0)
    for (size_t i = 0; i < n; ++i) {
        const uint16_t m = raw_ids[i];
        if (current != m || count == 255) {
            HBO_PUSH_REC();
            current = m;
            count = 1;
        } else {
            count++;
        }
    }
    HBO_PUSH_REC();
#undef HBO_PUSH_REC
    return w;
}

/* CubeGrid::decode record loop (codec.rs:52-66): whole 3-byte records only (array_chunks drops a tail);
 * PaletteMaterialId::new(v) = NonZeroU16::new(v + 1) (material.rs:219-222), i.e. wire 0 comes back as the first
 * material, never as None; v = 65535 wraps to None in a release build.  Voxels the stream does not reach stay
 * None.  Returns 0, or -1 where the reference indexes past the chunk and panics. */
int hbo_rle_decode(const uint8_t *bytes, size_t nbytes, uint16_t *raw_ids) {
    memset(raw_ids, 0, sizeof(uint16_t) * HBO_CHUNK_VOLUME);
    size_t i = 0;
    for (size_t r = 0; r + 3 <= nbytes; r += 3) {
        const uint8_t count = bytes[r];
        const uint16_t v = (uint16_t)(bytes[r + 1] | (bytes[r + 2] << 8));
        const uint16_t raw = (uint16_t)(v + 1u);
        for (uint8_t k = 0; k < count; ++k) {
            if (i >= HBO_CHUNK_VOLUME) return -1;
            raw_ids[i++] = raw;
        }
    }
    return 0;
}

/* encode_palette (codec.rs:70-75): for (i, material) in palette.materials().enumerate(): u16 LE i, then
 * Material::encode (material.rs:73-99): u8 group.len, u8 key.len, group bytes, key bytes,
 * u8 (cullable_faces.bits() << 6 | has_collider) -- a u8 shift, the upper four face bits fall off --,
 * Texture::Colors: u8 0, u16 LE vec.len, r g b a as f32 LE per colour; f32 LE toughness.
 * groups / keys / cullable / collider / toughness: one entry per material (the colours come from the palette).
 * Returns the number of bytes the palette has; writes at most cap. */
size_t hbo_encode_palette(const hbo_palette *pal, const char *const *groups, const char *const *keys, const uint8_t *cullable,
                          const uint8_t *collider, const float *toughness, uint8_t *out, size_t cap) {
    size_t n = 0;
#define HBO_PUT(ptr, len)                                                   \
    do {                                                                    \
        for (size_t q_ = 0; q_ < (size_t)(len); ++q_, ++n)                  \
            if (n < cap) out[n] = ((const uint8_t *)(ptr))[q_];             \
    } while (0)
    for (int32_t i = 0; i < pal->n_materials; ++i) {
        const uint8_t idx[2] = {(uint8_t)(i & 255), (uint8_t)(i >> 8)};
        HBO_PUT(idx, 2);
        const uint8_t gl = (uint8_t)strlen(groups[i]), kl = (uint8_t)strlen(keys[i]);
        HBO_PUT(&gl, 1);
        HBO_PUT(&kl, 1);
        HBO_PUT(groups[i], gl);
        HBO_PUT(keys[i], kl);
        const uint8_t e0 = (uint8_t)((uint8_t)(cullable[i] << 6) | (collider[i] ? 1 : 0));
        HBO_PUT(&e0, 1);
        const uint8_t tex = 0;
        HBO_PUT(&tex, 1);
        const uint8_t nc[2] = {(uint8_t)(pal->mat[i].n_colors & 255), (uint8_t)(pal->mat[i].n_colors >> 8)};
        HBO_PUT(nc, 2);
        for (int32_t c = 0; c < pal->mat[i].n_colors; ++c) HBO_PUT(pal->mat[i].rgba[c], 16);
        HBO_PUT(&toughness[i], 4);
    }
#undef HBO_PUT
    return n;
}
