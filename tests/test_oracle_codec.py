"""CPU checks of the oracle's restatement of the on-disk cube codec (server/src/chunk/codec.rs) against an
independent numpy statement of the same record rules."""
import numpy as np

import oracle as O


def np_encode(ids):
    """Independent statement: maximal runs of equal raw ids, cut every 255; the (None, 0) start state merges with a
    leading None run, otherwise it is closed as an empty record; wire value = index (raw - 1), None -> 0."""
    ids = np.asarray(ids, dtype=np.uint16)
    starts = np.flatnonzero(np.r_[True, ids[1:] != ids[:-1]])
    ends = np.r_[starts[1:], ids.shape[0]]
    recs = []
    if ids.shape[0] == 0 or ids[0] != 0:
        recs.append((0, 0))
    for s, e in zip(starts, ends):
        wire = int(ids[s]) - 1 if ids[s] else 0
        n = int(e - s)
        while n > 255:
            recs.append((255, wire))
            n -= 255
        recs.append((n, wire))
    out = np.zeros(3 * len(recs), dtype=np.uint8)
    for i, (c, w) in enumerate(recs):
        out[3 * i:3 * i + 3] = (c, w & 0xFF, w >> 8)
    return out


def cases():
    rng = np.random.default_rng(3)
    w = O.world(0)
    yield "air", np.zeros(O.CHUNK_VOLUME, dtype=np.uint16)
    yield "stone", np.ones(O.CHUNK_VOLUME, dtype=np.uint16)
    yield "checker", (np.arange(O.CHUNK_VOLUME) % 2 * 3).astype(np.uint16)
    yield "checker_solid_first", ((np.arange(O.CHUNK_VOLUME) + 1) % 2 * 2).astype(np.uint16)
    for k in (254, 255, 256, 510, 511):
        a = np.zeros(O.CHUNK_VOLUME, dtype=np.uint16)
        a[k:] = 2
        yield f"air{k}", a
        b = np.full(O.CHUNK_VOLUME, 3, dtype=np.uint16)
        b[k:2 * k] = 0
        yield f"solid{k}", b
    yield "terrain", O.generate_ids(w, [(0, 0, 0)])[0]
    yield "terrain_deep", O.generate_ids(w, [(3, -1, -2)])[0]
    yield "random", rng.integers(0, 4, O.CHUNK_VOLUME).astype(np.uint16)
    r = rng.integers(0, 4, O.CHUNK_VOLUME // 64).astype(np.uint16)
    yield "random_runs", np.repeat(r, 64)


def test_encode_matches_independent_statement():
    for name, ids in cases():
        assert O.rle_encode(ids).tobytes() == np_encode(ids).tobytes(), name


def test_decode_is_the_lossy_inverse():
    """decode(encode(x)): run structure survives, None comes back as the first material (wire 0 -> raw 1)."""
    for name, ids in cases():
        rc, back = O.rle_decode(O.rle_encode(ids))
        assert rc == 0, name
        want = np.where(ids == 0, 1, ids).astype(np.uint16)
        assert (back == want).all(), name


def test_decode_edges():
    rc, ids = O.rle_decode(np.zeros(0, dtype=np.uint8))
    assert rc == 0 and not ids.any()
    # partial tail ignored; short stream leaves None behind; 65535 wraps to None
    rc, ids = O.rle_decode(np.array([3, 1, 0, 2, 0xFF, 0xFF, 9, 9], dtype=np.uint8))
    assert rc == 0 and ids[:6].tolist() == [2, 2, 2, 0, 0, 0] and not ids[5:].any()
    # one voxel too many: the reference panics
    full = np.tile(np.array([255, 0, 0], dtype=np.uint8), 128)
    rc, _ = O.rle_decode(np.r_[full, np.array([128, 0, 0], dtype=np.uint8)])
    assert rc == 0
    rc, _ = O.rle_decode(np.r_[full, np.array([129, 0, 0], dtype=np.uint8)])
    assert rc == -1
