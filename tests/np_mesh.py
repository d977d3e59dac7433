"""Independent Python restatement of the client mesher, transcribed from the Rust sources only
(NOT from oracle/hb_oracle.c) and used to cross-check the C oracle the way tests/np_noise.py does for noise:

    generate_mesh / is_cube_present / vertex_ao / facial_ao   client/src/world/chunk.rs:176-276
    create_chunk_shell order                                   client/src/world/chunk.rs:158-174
    CubeFace::rotation / orthonormal_basis, FaceIter            lib/src/spatial/face.rs:60-93, 319-333
    Quat::from_euler / to_axes                                   lib/src/rotation.rs:77-116
    Mat3 * Mat3, Mat3::from(Vec3)                                lib/src/matrix/mat3.rs:63-90
    Instance3d::new                                              client/src/video/world/vertex.rs:53-68
    Material::get_color                                          server/src/chunk/material.rs:67-71
    CubeFlags::faces                                             server/src/chunk/cube.rs:35-41
    std DefaultHasher = SipHash-1-3, keys (0, 0), over the derive(Hash) bytes of ChunkPt (x, y, z as i32)
    fastrand 2.3.0 (Cargo.lock:712-715, un-vendored): Rng::with_seed, gen_u64 (wyrand), f32()

fastrand is not under /root/reference and this image holds no copy or published test vector of it (searched
site-packages, /usr/include, /usr/share for the wyrand constants: nothing).  The generator is restated from its
published algorithm: wyhash final version 4's `wyrand` -- state += 0x2d358dccaa6c78a5, output = lo64 ^ hi64 of the
128-bit product state * (state ^ 0x8bb84b93962eacc9) -- which fastrand adopted in 2.0.0, and
`f32() = from_bits((1 << 30) - (1 << 23) + (u32() >> 9)) - 1.0` with `u32() = gen_u64() as u32`.  That part of the
parity chain stays UNPINNED by anything the reference issued.

Scalar loops on purpose (clarity over speed): a 32^3 chunk with a few thousand faces takes about a second."""
import ctypes
import ctypes.util
import math
import struct

import numpy as np

f32 = np.float32
MASK64 = (1 << 64) - 1
CHUNK_LENGTH = 32

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]


# ---------------------------------------------------------------- hashing and the colour stream
def _rotl(x, b):
    return ((x << b) | (x >> (64 - b))) & MASK64


def siphash13(data: bytes, k0: int = 0, k1: int = 0) -> int:
    """SipHash-c-d with c = 1, d = 3 (Aumasson & Bernstein); std::hash::DefaultHasher::new() uses keys (0, 0)."""
    v0 = k0 ^ 0x736F6D6570736575
    v1 = k1 ^ 0x646F72616E646F6D
    v2 = k0 ^ 0x6C7967656E657261
    v3 = k1 ^ 0x7465646279746573

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & MASK64; v1 = _rotl(v1, 13); v1 ^= v0; v0 = _rotl(v0, 32)
        v2 = (v2 + v3) & MASK64; v3 = _rotl(v3, 16); v3 ^= v2
        v0 = (v0 + v3) & MASK64; v3 = _rotl(v3, 21); v3 ^= v0
        v2 = (v2 + v1) & MASK64; v1 = _rotl(v1, 17); v1 ^= v2; v2 = _rotl(v2, 32)
        return v0, v1, v2, v3

    n = len(data)
    for i in range(0, n - n % 8, 8):
        m = int.from_bytes(data[i:i + 8], "little")
        v3 ^= m
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    b = ((n & 0xFF) << 56) | int.from_bytes(data[n - n % 8:], "little")
    v3 ^= b
    v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= b
    v2 ^= 0xFF
    for _ in range(3):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return v0 ^ v1 ^ v2 ^ v3


def chunk_seed(pos) -> int:
    """`center_chunk.position.hash(&mut hasher); hasher.finish()`: three write_i32 calls, native (little) endian."""
    return siphash13(struct.pack("<iii", *[int(v) for v in pos]))


class Rng:
    """fastrand::Rng (2.3.0)."""

    def __init__(self, seed: int):
        self.s = seed & MASK64  # Rng::with_seed

    def gen_u64(self) -> int:
        self.s = (self.s + 0x2D358DCCAA6C78A5) & MASK64
        t = self.s * (self.s ^ 0x8BB84B93962EACC9)
        return (t & MASK64) ^ (t >> 64)

    def f32(self):
        bits = (1 << 30) - (1 << 23) + ((self.gen_u64() & 0xFFFFFFFF) >> 9)
        return f32(np.uint32(bits).view(f32) - f32(1.0))


# ---------------------------------------------------------------- faces
FACES = ["East", "West", "Up", "Down", "North", "South"]  # CubeFace::VALUES, also the FaceIter bit order
X, Y, Z = np.array([1, 0, 0]), np.array([0, 1, 0]), np.array([0, 0, 1])
BASIS = {"East": (-Z, Y, X), "West": (Z, Y, -X), "Up": (X, -Z, Y), "Down": (X, Z, -Y), "North": (X, Y, Z),
         "South": (-X, Y, -Z)}
HALF_PI, PI = f32(math.pi / 2), f32(math.pi)  # std::f32::consts
EULER = {"East": (f32(0), HALF_PI, f32(0)), "West": (f32(0), f32(-HALF_PI), f32(0)), "Up": (f32(-HALF_PI), f32(0), f32(0)),
         "Down": (HALF_PI, f32(0), f32(0)), "North": (f32(0), f32(0), f32(0)), "South": (f32(0), PI, f32(0))}


def _sin_cos(a):
    return f32(_libm.sinf(float(a))), f32(_libm.cosf(float(a)))


def quat_from_euler(yaw, pitch, roll):
    sx, cx = _sin_cos(f32(yaw / f32(2)))
    sy, cy = _sin_cos(f32(pitch / f32(2)))
    sz, cz = _sin_cos(f32(roll / f32(2)))
    return (f32(f32(f32(sx * cy) * cz) - f32(f32(cx * sy) * sz)),
            f32(f32(f32(cx * sy) * cz) + f32(f32(sx * cy) * sz)),
            f32(f32(f32(cx * cy) * sz) - f32(f32(sx * sy) * cz)),
            f32(f32(f32(cx * cy) * cz) + f32(f32(sx * sy) * sz)))


def quat_to_axes(q):
    x, y, z, w = q
    one, two = f32(1), f32(2)
    mx = [f32(one - f32(two * f32(f32(y * y) + f32(z * z)))), f32(two * f32(f32(x * y) + f32(z * w))), f32(two * f32(f32(x * z) - f32(y * w)))]
    my = [f32(two * f32(f32(x * y) - f32(z * w))), f32(one - f32(two * f32(f32(x * x) + f32(z * z)))), f32(two * f32(f32(y * z) + f32(x * w)))]
    mz = [f32(two * f32(f32(x * z) + f32(y * w))), f32(two * f32(f32(y * z) - f32(x * w))), f32(one - f32(two * f32(f32(x * x) + f32(y * y))))]
    return np.array([mx, my, mz], dtype=f32)  # rows here = the Mat3's column vectors x, y, z


def mat3_mul_vec(m, v):
    # self.x * rhs.xxx() + self.y * rhs.yyy() + self.z * rhs.zzz(), every product and sum rounded to f32
    return ((m[0] * v[0]).astype(f32) + (m[1] * v[1]).astype(f32)).astype(f32) + (m[2] * v[2]).astype(f32)


def mat3_mul(a, b):
    return np.array([mat3_mul_vec(a, b[0]), mat3_mul_vec(a, b[1]), mat3_mul_vec(a, b[2])], dtype=f32)


def face_model(face):
    rot = quat_to_axes(quat_from_euler(*EULER[face]))
    scale = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=f32)  # Mat3::from(Vec3::ONE)
    return mat3_mul(rot, scale)


# ---------------------------------------------------------------- ambient occlusion
def is_cube_present(shell, p) -> bool:
    off = [int(c) // CHUNK_LENGTH for c in p]  # div_euclid by a positive divisor == floor division
    if any(o < -1 or o > 1 for o in off):
        return False
    index = (off[2] + 1) + (off[1] + 1) * 3 + (off[0] + 1) * 9  # Vec3::linearize(3): order (z, y, x)
    target = shell[index]
    if target is None:
        return False
    lx, ly, lz = [int(c) % CHUNK_LENGTH for c in p]
    return bool(target["material"][lx * 1024 + lz * 32 + ly] != 0)


def vertex_ao(s1, s2, c):
    return 3 if (s1 and s2) else int(s1) + int(s2) + int(c)


def facial_ao(shell, face, position):
    u, v, n = BASIS[face]
    p = np.asarray(position)
    tl = is_cube_present(shell, p + n - v - u)
    tc = is_cube_present(shell, p + n - v)
    tr = is_cube_present(shell, p + n - v + u)
    ml = is_cube_present(shell, p + n - u)
    mr = is_cube_present(shell, p + n + u)
    bl = is_cube_present(shell, p + n + v - u)
    bc = is_cube_present(shell, p + n + v)
    br = is_cube_present(shell, p + n + v + u)
    occ = [vertex_ao(ml, tc, tl), vertex_ao(ml, bc, bl), vertex_ao(mr, tc, tr), vertex_ao(mr, bc, br)]
    # `Vec4::new(u8..).cast() / 3.0` has no other constraint on its float type than the literal -> f64 (the
    # compiler's fallback), then `.cast::<f32>()`, negate, + 1.0, max_each(0.15)
    out = []
    for o in occ:
        a = f32(float(o) / 3.0)
        r = f32(f32(-a) + f32(1.0))
        out.append(r if r > f32(0.15) else f32(0.15))
    return out


# ---------------------------------------------------------------- generate_mesh
def generate_mesh(shell, palette):
    """shell: 27 entries (index 9*(dx+1) + 3*(dy+1) + (dz+1)), each None or
    {"position": (cx, cy, cz), "material": u16[32768] (0 = None, k = palette[k-1]), "flags": u32[32768]}, arrays indexed by
    vec3u5::linearize = x*1024 + z*32 + y.  palette: list of colour lists (r, g, b, a).
    Returns a list of 25-word records (24 floats + the u32 light at word 20) as raw 100-byte strings."""
    centre = shell[13]
    rng = Rng(chunk_seed(centre["position"]))
    base = [int(c) * CHUNK_LENGTH for c in centre["position"]]
    models = {f: face_model(f) for f in FACES}
    out = []
    for x in range(CHUNK_LENGTH):
        for z in range(CHUNK_LENGTH):
            for y in range(CHUNK_LENGTH):
                i = x * 1024 + z * 32 + y
                perms = [rng.f32() for _ in range(6)]  # PerFace::mapped draws for every voxel, solid or not
                mat = int(centre["material"][i])
                if mat == 0:
                    continue
                if mat - 1 >= len(palette):  # palette.get_by_id(..) is None -> `continue`
                    continue
                colours = palette[mat - 1]
                flags = int(centre["flags"][i])
                faces = 0 if flags & 1 == 0 else ((flags >> 1) & 0xFF) & 0b0011_1111
                for k, face in enumerate(FACES):
                    if not faces & (1 << k):
                        continue
                    sel = int(f32(f32(max(len(colours) - 1, 0)) * perms[k]))  # saturating_sub(1) as f32 * p, `as usize`
                    rgba = colours[sel]
                    ao = facial_ao(shell, face, (x, y, z))
                    m = models[face]
                    pos = [f32(base[0] + x), f32(base[1] + y), f32(base[2] + z)]
                    words = [m[0][0], m[0][1], m[0][2], f32(0), m[1][0], m[1][1], m[1][2], f32(0),
                             m[2][0], m[2][1], m[2][2], f32(0), pos[0], pos[1], pos[2], f32(1),
                             f32(rgba[0]), f32(rgba[1]), f32(rgba[2]), f32(rgba[3])]
                    out.append(struct.pack("<20f", *[float(w) for w in words]) + struct.pack("<I", 0) +
                               struct.pack("<4f", *[float(a) for a in ao]))
    return out
