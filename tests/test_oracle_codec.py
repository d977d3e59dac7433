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


GAME_DESCS = [("herbolution", "stone", 0x3F, True, 10.0), ("herbolution", "dirt", 0x3F, True, 0.95),
              ("herbolution", "grass", 0x3F, True, 1.05)]  # Material::stone / dirt / grass, material.rs:27-61


def py_encode_palette(colors, descs) -> bytes:
    """encode_palette + Material::encode transcribed from codec.rs:70-75 / material.rs:73-99 (independent of the C oracle)."""
    import struct
    buf = bytearray()
    for i, (cols, (group, key, cull, coll, tough)) in enumerate(zip(colors, descs)):
        buf += struct.pack("<H", i)
        buf += bytes([len(group.encode()), len(key.encode())]) + group.encode() + key.encode()
        buf.append(((cull << 6) & 0xFF) | int(bool(coll)))  # u8 arithmetic
        buf.append(0)
        buf += struct.pack("<H", len(cols))
        for rgba in cols:
            buf += struct.pack("<4f", *rgba)
        buf += struct.pack("<f", tough)
    return bytes(buf)


def test_palette_bytes_oracle_vs_python_restatement():
    game = [[(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0), (0.7, 0.7, 0.7, 1.0)],
            [(0.4, 0.3, 0.2, 1.0), (0.5, 0.4, 0.3, 1.0), (0.6, 0.5, 0.4, 1.0)],
            [(0.1, 0.8, 0.1, 1.0), (0.2, 0.9, 0.2, 1.0), (0.3, 1.0, 0.3, 1.0)]]
    got = O.encode_palette(O.default_palette(), GAME_DESCS)
    assert got == py_encode_palette(game, GAME_DESCS)
    assert got[:2] == b"\x00\x00" and got[2:4] == bytes([11, 5]) and got[4:20] == b"herbolutionstone" and got[20] == 0xC1
    custom = [[(0.25, 0.5, 0.75, 1.0)], [(0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0, 0.5)]]
    descs = [("a", "", 0x15, False, -2.5), ("grp", "k" * 200, 0x02, True, 0.0)]
    assert O.encode_palette(O.make_palette(custom), descs) == py_encode_palette(custom, descs)
