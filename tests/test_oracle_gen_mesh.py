"""CPU tests of the oracle's generation / culling / meshing restatement (no GPU): the literal
transcription of CubeMesh::set + cull_shared_faces + generate_mesh must agree with the closed-form
visibility rule the GPU kernels implement; hashing is pinned against CPython's siphash13."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NORMALS = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]


def test_siphash13_golden_vectors():
    for v in json.load(open(os.path.join(GOLD, "siphash13_vectors.json"))):
        x, y, z = v["pt"]
        assert O.lib().hbo_siphash13_chunk(x, y, z) == v["hash"], v


def test_siphash13_live_against_cpython():
    assert sys.hash_info.algorithm == "siphash13"
    pts = [(3, 4, 5), (-100, 0, 100), (7, -7, 7)]
    code = "import struct,json;print(json.dumps([hash(struct.pack('<iii',*p)) & 0xFFFFFFFFFFFFFFFF for p in %r]))" % (pts,)
    out = subprocess.check_output([sys.executable, "-c", code], env={**os.environ, "PYTHONHASHSEED": "0"})
    for p, h in zip(pts, json.loads(out)):
        assert O.lib().hbo_siphash13_chunk(*p) == h


def test_wyrand_counter_form_and_range():
    """fastrand's wyrand is a counter generator: draw n uses state seed + (n+1)*C (what the CUDA kernel
    relies on to evaluate draw 6*L+face directly)."""
    seed = 0x1234_5678_9ABC_DEF0
    st = C.c_uint64(seed)
    seq = [O.lib().hbo_wyrand_next(C.byref(st)) for _ in range(50)]
    mask = (1 << 64) - 1
    for n, v in enumerate(seq):
        s = (seed + (n + 1) * 0x2D358DCCAA6C78A5) & mask
        t = s * (s ^ 0x8BB84B93962EACC9)
        assert v == ((t & mask) ^ (t >> 64))
    st = C.c_uint64(42)
    fs = [O.lib().hbo_wyrand_f32(C.byref(st)) for _ in range(2000)]
    assert 0.0 <= min(fs) and max(fs) < 1.0 and 0.4 < float(np.mean(fs)) < 0.6


def test_face_matrices_and_ao_table():
    """Instance3d model columns per face (vertex.rs:53-62): rotations of the +Z quad onto each face."""
    exact = {0: [[0, 0, -1], [0, 1, 0], [1, 0, 0]], 1: [[0, 0, 1], [0, 1, 0], [-1, 0, 0]],
             2: [[1, 0, 0], [0, 0, -1], [0, 1, 0]], 3: [[1, 0, 0], [0, 0, 1], [0, -1, 0]],
             4: [[1, 0, 0], [0, 1, 0], [0, 0, 1]], 5: [[-1, 0, 0], [0, 1, 0], [0, 0, -1]]}
    for f in range(6):
        m = np.zeros(12, dtype=np.float32)
        O.lib().hbo_face_matrix(f, m.ctypes.data)
        cols = m.reshape(3, 4)
        assert (cols[:, 3] == 0).all()
        assert np.abs(cols[:, :3] - np.array(exact[f], dtype=np.float32)).max() < 1e-6
        # third column (the quad's normal after rotation) points along the face normal
        assert np.allclose(cols[2, :3], NORMALS[f], atol=1e-6)
    ao = [O.lib().hbo_ao_value(k) for k in range(4)]
    assert ao[0] == 1.0 and ao[3] == np.float32(0.15) and ao[0] > ao[1] > ao[2] > ao[3]
    assert abs(ao[1] - 2 / 3) < 1e-6 and abs(ao[2] - 1 / 3) < 1e-6


def test_set_path_equals_closed_form_image():
    """CubeMesh::set bookkeeping (mesh.rs:125-232) vs the closed form 'face kept iff in-chunk neighbour
    is air or outside; air voxel flags=1 iff an in-chunk neighbour is solid'."""
    w = O.world(0)
    for p in [(0, 0, 0), (0, -1, 0), (1, 0, -1), (-3, 1, 2), (5, -1, 5), (-8, -2, 7)]:
        m = O.generate_literal(w, p)
        lit = O.mesh_cubes(m)
        ids = O.generate_ids(w, [p])[0]
        assert (lit["material"] == ids).all()
        assert lit.tobytes() == O.cubes_from_ids(ids).tobytes(), p
        assert m.n_updated >= int((ids != 0).sum())


def test_set_path_order_independent_and_removal():
    """Random insertion order and removals (dig/place) leave flags consistent with the closed form."""
    rng = np.random.default_rng(5)
    m = O.new_mesh((0, 0, 0))
    ids = np.zeros(32768, dtype=np.uint16)
    coords = rng.integers(0, 32, size=(6000, 3))
    for x, y, z in coords:
        mat = int(rng.integers(0, 4))  # 0 removes
        O.lib().hbo_mesh_set(C.byref(m), int(x), int(y), int(z), mat)
        ids[x * 1024 + z * 32 + y] = mat
    lit = O.mesh_cubes(m)
    ref = O.cubes_from_ids(ids)
    assert (lit["material"] == ids).all()
    solid = ids != 0
    assert (lit["flags"][solid] == ref["flags"][solid]).all()
    # air voxels: flags are either untouched (0) or an opaque tag with no faces... after removals the
    # reference leaves faces on air voxels (add_neighboring_faces inserts into neighbours regardless of
    # their material, mesh.rs:245-250); the mesher ignores air voxels, so only solid flags matter.


def test_literal_pipeline_equals_closed_form():
    for seed, reg in [(0, (-2, -2, 4, 4, -3, 6)), (5, (10, -20, 3, 2, -1, 3))]:
        w = O.world(seed)
        a = O.region_pipeline(w, reg, True, threads=4, want_ids=True)
        b = O.region_pipeline(w, reg, False, threads=4, want_ids=True)
        assert (a["counts"] == b["counts"]).all() and (a["ids"] == b["ids"]).all()
        assert a["instances"].tobytes() == b["instances"].tobytes()
        assert a["total"] > 0


def test_oracle_mesh_digests_regression():
    import hashlib
    for e in json.load(open(os.path.join(GOLD, "oracle_digests.json")))["mesh"]:
        r = O.region_pipeline(O.world(e["seed"]), tuple(e["region"]), literal=True, threads=4)
        assert r["total"] == e["quads"]
        assert hashlib.sha256(r["counts"].tobytes()).hexdigest() == e["counts_sha256"]
        geo = r["instances"]["model"].tobytes() + r["instances"]["ao"].tobytes()
        assert hashlib.sha256(geo).hexdigest() == e["geometry_sha256"]
        assert hashlib.sha256(r["instances"]["rgba"].tobytes()).hexdigest() == e["colour_sha256"]


def test_mesh_invariants():
    """Every emitted face has an air or absent neighbour; seam faces are symmetric; order is (L, face)."""
    reg = (0, 0, 2, 2, -1, 2)
    w = O.world(0)
    r = O.region_pipeline(w, reg, False, threads=2, want_ids=True)
    ids = r["ids"].reshape(2, 2, 2, 32, 32, 32)  # [ix][iz][iy][x][z][y]
    inst = r["instances"]
    pos = inst["model"][:, 12:15].astype(np.int64)
    assert (inst["model"][:, 15] == 1.0).all() and (inst["light"] == 0).all()

    def solid(wx, wy, wz):
        ix, iz, iy = wx // 32, wz // 32, wy // 32 + 1
        if not (0 <= ix < 2 and 0 <= iz < 2 and 0 <= iy < 2):
            return False
        return ids[ix, iz, iy, wx % 32, wz % 32, wy % 32] != 0

    normal_of = {}
    for f in range(6):
        m = np.zeros(12, dtype=np.float32)
        O.lib().hbo_face_matrix(f, m.ctypes.data)
        normal_of[m.tobytes()] = f
    sample = np.random.default_rng(1).integers(0, len(inst), size=400)
    for i in sample:
        f = normal_of[inst["model"][i, :12].tobytes()]
        x, y, z = pos[i]
        assert solid(x, y, z)
        n = NORMALS[f]
        assert not solid(x + n[0], y + n[1], z + n[2])
    # emission order inside each chunk: ascending L then face
    off = np.concatenate([[0], np.cumsum(r["counts"].astype(np.int64))])
    for c in range(len(r["counts"])):
        seg = slice(off[c], off[c + 1])
        p = pos[seg] & 31
        L = p[:, 0] * 1024 + p[:, 2] * 32 + p[:, 1]
        fidx = np.array([normal_of[m.tobytes()] for m in inst["model"][seg][:, :12]])
        key = L * 8 + fidx
        assert (np.diff(key) > 0).all()
    # colours come from the first two palette entries of the voxel's material (get_color, material.rs:67-71)
    pal = [[(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0)], [(0.4, 0.3, 0.2, 1.0), (0.5, 0.4, 0.3, 1.0)],
           [(0.1, 0.8, 0.1, 1.0), (0.2, 0.9, 0.2, 1.0)]]
    allowed = {np.array(c, dtype=np.float32).tobytes() for m in pal for c in m}
    assert {c.tobytes() for c in np.unique(inst["rgba"], axis=0)} <= allowed


def test_cull_shared_faces_pairwise():
    """cull_shared_faces (mesh.rs:76-123) clears exactly the faces between two solid border voxels."""
    w = O.world(0)
    a = O.generate_literal(w, (0, -1, 0))
    b = O.generate_literal(w, (1, -1, 0))
    before_a, before_b = O.mesh_cubes(a), O.mesh_cubes(b)
    O.lib().hbo_cull_shared_faces(C.byref(a), C.byref(b))
    after_a, after_b = O.mesh_cubes(a), O.mesh_cubes(b)
    L = np.arange(32768)
    x = L >> 10
    both = (before_a["material"][(31 << 10) + (L & 1023)] != 0) & (before_b["material"][L & 1023] != 0)
    ea = (before_a["flags"] != after_a["flags"])
    assert not ea[x != 31].any()
    changed = ea[x == 31]
    expect = both[:1024] & (((before_a["flags"][x == 31] >> 1) & 1) == 1)
    assert (changed == expect).all()
    assert (((after_a["flags"][x == 31] >> 1) & 1)[both[:1024]] == 0).all()       # East bit cleared
    assert (((after_b["flags"][x == 0] >> 2) & 1)[both[:1024]] == 0).all()        # West bit cleared
    # non-adjacent chunks: no-op
    c = O.generate_literal(w, (5, -1, 0))
    snap = O.mesh_cubes(a)
    O.lib().hbo_cull_shared_faces(C.byref(a), C.byref(c))
    assert O.mesh_cubes(a).tobytes() == snap.tobytes()
