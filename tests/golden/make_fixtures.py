"""Regenerates the JSON fixtures in tests/golden/.

    python tests/golden/make_fixtures.py

Provenance of each fixture (NONE is issued by the reference: it ships no tests or vectors, and its
Rust sources cannot be compiled in this image -- SURVEY.md s4, s8c):

  survey_fbm_values.json   copied from SURVEY.md s8(c): values of an independent numpy float32
                           restatement written during the survey (not by this script).
  siphash13_vectors.json   CPython's siphash13 with a zero key (PYTHONHASHSEED=0), i.e. an independent
                           implementation of the algorithm behind Rust's DefaultHasher::new().
  oracle_digests.json      sha256 digests of the ORACLE's own outputs (noise tiles, block ids, meshes);
                           regression anchors that also travel to the GPU box, where /root/reference
                           and the numpy cross-check are not needed.
"""
import hashlib
import json
import os
import struct
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def siphash_vectors():
    pts = [(0, 0, 0), (1, -2, 3), (-512, 5, 511), (2147483647, -2147483648, 7), (-1, -1, -1), (16, -5, -16),
           (262143, 0, -262143)]
    code = ("import struct,json;print(json.dumps([hash(struct.pack('<iii',*p)) & 0xFFFFFFFFFFFFFFFF for p in %r]))" % (pts,))
    out = subprocess.check_output([sys.executable, "-c", code], env={**os.environ, "PYTHONHASHSEED": "0"})
    assert sys.hash_info.algorithm == "siphash13", sys.hash_info
    return [{"pt": list(p), "hash": h} for p, h in zip(pts, json.loads(out))]


def oracle_digests():
    import numpy as np
    import oracle as O
    d = {"noise": [], "ids": [], "mesh": []}
    for seed in (0, 5):
        w = O.world(seed)
        for cx, cz in [(0, 0), (-1, -1), (3, -5), (-512, 511)]:
            d["noise"].append({"seed": seed, "col": [cx, cz],
                               "sha256": hashlib.sha256(O.noise_tile2(w, cx, cz).tobytes()).hexdigest()})
        for p in [(0, 0, 0), (0, -1, 0), (5, 1, -7)]:
            d["ids"].append({"seed": seed, "pt": list(p),
                             "sha256": hashlib.sha256(O.generate_ids(w, [p])[0].tobytes()).hexdigest()})
        reg = (-2, -2, 4, 4, -3, 6)
        r = O.region_pipeline(w, reg, literal=True, threads=4)
        d["mesh"].append({"seed": seed, "region": list(reg), "quads": r["total"],
                          "counts_sha256": hashlib.sha256(r["counts"].tobytes()).hexdigest(),
                          "geometry_sha256": hashlib.sha256(
                              r["instances"]["model"].tobytes() + r["instances"]["ao"].tobytes()).hexdigest(),
                          "colour_sha256": hashlib.sha256(r["instances"]["rgba"].tobytes()).hexdigest()})
    # the crate's other noise kinds
    d["noise_kinds"] = []
    for kind, name in ((O.NOISE_RIDGE, "ridge"), (O.NOISE_TURBULENCE, "turbulence"), (O.NOISE_GRADIENT, "gradient")):
        for seed in (0, 5):
            t = O.noise_tile2(O.world(seed, kind=kind), 3, -5)
            d["noise_kinds"].append({"kind": kind, "name": name, "seed": seed, "col": [3, -5],
                                     "sha256": hashlib.sha256(t.tobytes()).hexdigest()})
    # on-disk cube codec
    d["codec"] = []
    for seed in (0, 5):
        w = O.world(seed)
        for p in [(0, 0, 0), (0, -1, 0), (5, 1, -7), (-3, -2, 9)]:
            enc = O.rle_encode(O.generate_ids(w, [p])[0])
            d["codec"].append({"seed": seed, "pt": list(p), "bytes": int(enc.shape[0]),
                               "sha256": hashlib.sha256(enc.tobytes()).hexdigest()})
    # device-resident world: load, scripted edits, sync, remesh
    d["world"] = []
    for seed in (0, 5):
        m = O.ChunkMapOracle(O.world(seed))
        pts, edits = world_script(seed)
        for p in pts:
            m.load(p)
        m.apply(m.drain())
        for pos, mat in edits:
            m.set_cube(pos, mat)
        upd = m.drain()
        m.apply(upd)
        cubes = hashlib.sha256(b"".join(m.cubes(p).tobytes() for p in pts)).hexdigest()
        dirty = sorted(upd)
        mesh = hashlib.sha256()
        quads = 0
        for p in dirty:
            inst = m.remesh(p)
            quads += inst.shape[0]
            mesh.update(inst.tobytes())
        d["world"].append({"seed": seed, "cubes_sha256": cubes, "dirty": [list(p) for p in dirty],
                           "update_positions": int(sum(len(set(v["pos"].tolist())) for v in upd.values())),
                           "quads": quads, "mesh_sha256": mesh.hexdigest()})
        m.close()
    return d


def world_script(seed):
    """The fixed load set and edit list behind the "world" digests (also used by the GPU golden test)."""
    import numpy as np
    pts = [(cx, cy, cz) for cx in (0, 1) for cz in (-1, 0) for cy in (-1, 0)]
    rng = np.random.default_rng(4242 + seed)
    edits = []
    for _ in range(300):
        wx, wz = int(rng.integers(-2, 66)), int(rng.integers(-34, 34))
        wy = int(rng.integers(-20, 20))
        edits.append(((wx, wy, wz), int(rng.choice([0, 0, 1, 2, 3]))))
    return pts, edits


if __name__ == "__main__":
    with open(os.path.join(HERE, "siphash13_vectors.json"), "w") as f:
        json.dump(siphash_vectors(), f, indent=1)
    with open(os.path.join(HERE, "oracle_digests.json"), "w") as f:
        json.dump(oracle_digests(), f, indent=1)
    print("fixtures written")
