"""Independent numpy float32 restatement of simplex_2d / fbm_2d (crates/simd-noise/src/functions/
simplex_32.rs:103-170, fbm_32.rs:18-31), vectorised, used to cross-check the C oracle.  Every op is a
separate float32 op; the six neg_mul_add sites are evaluated as fma via float64 (a*b is exact in
float64 for float32 inputs; the final rounding to float32 can double-round in ~1e-9 of cases)."""
import numpy as np

f32 = np.float32
PERM = np.array([
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21,
    10, 23, 190, 6, 148, 247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149,
    56, 87, 174, 20, 125, 136, 171, 168, 68, 175, 74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122,
    60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54, 65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132,
    187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64, 52, 217, 226, 250,
    124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183,
    170, 213, 119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110,
    79, 113, 224, 232, 178, 185, 112, 104, 218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81,
    51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157, 184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150,
    254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180], dtype=np.int64)
PERM512 = np.concatenate([PERM, PERM])
F2 = f32(0.36602540378)
G2 = f32(0.2113248654)
G22 = f32(G2 * f32(2.0))


def fnma(a, b, c):  # c - a*b, fused
    return (c.astype(np.float64) - a.astype(np.float64) * b.astype(np.float64)).astype(f32)


def grad2(seed, h):
    h = (h ^ int(np.int32(np.int64(seed) & 0xFFFFFFFF if seed >= 0 else np.int64(seed)))) & 7
    lt4 = h < 4
    xm = np.where(lt4, f32(1), f32(2)).astype(f32)
    ym = np.where(lt4, f32(2), f32(1)).astype(f32)
    b1 = (h & 1) == 0
    b2 = (h & 2) == 0
    gx = np.where(np.where(lt4, b1, b2), xm, -xm).astype(f32)
    gy = np.where(np.where(lt4, b2, b1), ym, -ym).astype(f32)
    return gx, gy


def simplex2(x, y, seed):
    x = x.astype(f32); y = y.astype(f32)
    s = F2 * (x + y)
    ips = np.floor(x + s); jps = np.floor(y + s)
    i = ips.astype(np.int64); j = jps.astype(np.int64)
    t = (i + j).astype(f32) * G2
    x0 = x - (ips - t); y0 = y - (jps - t)
    i1 = np.where(x0 >= y0, -1, 0); j1 = np.where(y0 > x0, -1, 0)
    x1 = (x0 + i1.astype(f32)) + G2; y1 = (y0 + j1.astype(f32)) + G2
    x2 = (x0 + f32(-1)) + G22; y2 = (y0 + f32(-1)) + G22
    ii = i & 255; jj = j & 255
    gi0 = PERM512[ii + PERM512[jj]]
    gi1 = PERM512[(ii - i1) + PERM512[jj - j1]]
    gi2 = PERM512[(ii + 1) + PERM512[jj + 1]]
    half = np.full_like(x0, 0.5, dtype=f32)
    t0 = fnma(y0, y0, fnma(x0, x0, half)); t1 = fnma(y1, y1, fnma(x1, x1, half)); t2 = fnma(y2, y2, fnma(x2, x2, half))
    t0 = np.where(t0 >= 0, t0, f32(0)).astype(f32); t1 = np.where(t1 >= 0, t1, f32(0)).astype(f32)
    t2 = np.where(t2 >= 0, t2, f32(0)).astype(f32)
    t40 = (t0 * t0) * (t0 * t0); t41 = (t1 * t1) * (t1 * t1); t42 = (t2 * t2) * (t2 * t2)
    gx, gy = grad2(seed, gi0); n0 = t40 * (gx * x0 + gy * y0)
    gx, gy = grad2(seed, gi1); n1 = t41 * (gx * x1 + gy * y1)
    gx, gy = grad2(seed, gi2); n2 = t42 * (gx * x2 + gy * y2)
    return ((n0 + (n1 + n2)) * f32(45.26450774985561631259)).astype(f32)


def fbm2(x, y, seed, octaves=6, lac=2.0, gain=0.6):
    x = x.astype(f32); y = y.astype(f32)
    lac = f32(lac); gain = f32(gain)
    result = simplex2(x, y, seed)
    amp = f32(1.0)
    for _ in range(1, octaves):
        x = x * lac; y = y * lac
        amp = f32(amp * gain)
        result = (simplex2(x, y, seed) * amp + result).astype(f32)
    return result


def ridge2(x, y, seed, octaves=6, lac=2.0, gain=0.6):
    """functions/ridge_32.rs:18-31 (neg_mul_add fused, the AVX2 column)."""
    x = x.astype(f32); y = y.astype(f32)
    lac = f32(lac); gain = f32(gain)
    result = (f32(1.0) - np.abs(simplex2(x, y, seed))).astype(f32)
    amp = f32(1.0)
    for _ in range(1, octaves):
        x = x * lac; y = y * lac
        amp = f32(amp * gain)
        result = (result + fnma(np.abs(simplex2(x, y, seed)), np.full_like(x, amp), np.full_like(x, f32(1.0)))).astype(f32)
    return result


def turbulence2(x, y, seed, octaves=6, lac=2.0, gain=0.6):
    """functions/turbulence_32.rs:18-32."""
    x = x.astype(f32); y = y.astype(f32)
    lac = f32(lac); gain = f32(gain)
    result = np.abs(simplex2(x, y, seed)).astype(f32)
    amp = f32(1.0)
    for _ in range(1, octaves):
        x = x * lac; y = y * lac
        amp = f32(amp * gain)
        result = (result + np.abs((simplex2(x, y, seed) * amp).astype(f32))).astype(f32)
    return result


def noise_tile(cx, cz, seed, freq=0.001, kind="fbm", **kw):
    """index x + z*32, like get_2d_noise_helper_f32 (noise/f32.rs:69-129)."""
    xs = (f32(cx) * f32(32) + np.arange(32, dtype=f32)).astype(f32)
    zs = (f32(cz) * f32(32) + np.arange(32, dtype=f32)).astype(f32)
    X, Z = np.meshgrid(xs, zs, indexing="xy")  # row = z
    fn = {"fbm": fbm2, "ridge": ridge2, "turbulence": turbulence2,
          "gradient": lambda x, y, seed, **_: simplex2(x.astype(f32), y.astype(f32), seed)}[kind]
    return fn((X * f32(freq)).ravel(), (Z * f32(freq)).ravel(), seed, **kw)
