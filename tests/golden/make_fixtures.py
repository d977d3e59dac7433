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
    return d


if __name__ == "__main__":
    with open(os.path.join(HERE, "siphash13_vectors.json"), "w") as f:
        json.dump(siphash_vectors(), f, indent=1)
    with open(os.path.join(HERE, "oracle_digests.json"), "w") as f:
        json.dump(oracle_digests(), f, indent=1)
    print("fixtures written")
