"""GPU parity for the on-disk cube codec (hb_rle_encode / hb_rle_decode) against the oracle's
restatement of server/src/chunk/codec.rs.  Bar: byte-exact streams, bit-exact ids."""
import numpy as np
import pytest

import herbolution_b200 as hb
import oracle as O
from test_oracle_codec import cases

pytestmark = pytest.mark.gpu


def streams(enc):
    return [enc["bytes"][int(o):int(o) + int(s)] for o, s in zip(enc["offsets"], enc["sizes"])]


def test_encode_edge_cases_byte_exact(ctx):
    names, ids = zip(*cases())
    enc = ctx.rle_encode(np.stack(ids))
    assert (enc["offsets"] % 12 == 0).all()
    for name, a, got in zip(names, ids, streams(enc)):
        assert got.tobytes() == O.rle_encode(a).tobytes(), name
    # gaps between streams are zero-filled
    mask = np.ones(enc["bytes"].shape[0], dtype=bool)
    for o, s in zip(enc["offsets"], enc["sizes"]):
        mask[int(o):int(o) + int(s)] = False
    assert not enc["bytes"][mask].any()


@pytest.mark.parametrize("seed", [0, 5])
def test_encode_decode_config2(ctx, seed):
    """All 4096 chunks of BASELINE config 2: encode byte-exact, decode bit-exact, and the size-independent
    property decode(encode(x)) == x with None read back as the first material."""
    ctx.set_world(seed)
    pts = hb.Region(-16, -16, 32, 32, -2, 4).chunk_points()
    ids = ctx.generate(pts)["ids"]
    enc = ctx.rle_encode(ids)
    rng = np.random.default_rng(seed)
    for i in rng.choice(len(pts), 96, replace=False):
        o, s = int(enc["offsets"][i]), int(enc["sizes"][i])
        assert enc["bytes"][o:o + s].tobytes() == O.rle_encode(ids[i]).tobytes(), i
    back = ctx.rle_decode(enc["bytes"], enc["offsets"], enc["sizes"])
    assert (back == np.where(ids == 0, 1, ids)).all()
    for i in rng.choice(len(pts), 16, replace=False):
        o, s = int(enc["offsets"][i]), int(enc["sizes"][i])
        rc, want = O.rle_decode(enc["bytes"][o:o + s])
        assert rc == 0 and (back[i] == want).all()
    ctx.set_world(0)


def test_decode_edges_and_errors(ctx):
    full = np.tile(np.array([255, 0, 0], dtype=np.uint8), 128)
    blobs = [
        np.zeros(0, dtype=np.uint8),
        np.array([3, 1, 0, 2, 0xFF, 0xFF, 9, 9], dtype=np.uint8),      # partial tail, wire 65535 wraps to None
        np.r_[full, np.array([128, 0, 0], dtype=np.uint8)],            # exactly 32768 voxels
        np.tile(np.array([1, 2, 0, 0, 7, 7], dtype=np.uint8), 5000),   # empty records in between, 5000 voxels
    ]
    data = np.concatenate(blobs)
    sizes = np.array([b.shape[0] for b in blobs], dtype=np.uint32)
    offsets = np.concatenate([[0], np.cumsum(sizes[:-1])]).astype(np.uint64)
    got = ctx.rle_decode(data, offsets, sizes)
    for i, b in enumerate(blobs):
        rc, want = O.rle_decode(b)
        assert rc == 0 and (got[i] == want).all(), i
    too_long = np.r_[full, np.array([129, 0, 0], dtype=np.uint8)]
    with pytest.raises(hb.HerbGpuError) as e:
        ctx.rle_decode(too_long, [0], [too_long.shape[0]])
    assert e.value.code == hb._lib.HB_ERR_INVALID
    with pytest.raises(hb.HerbGpuError) as e:
        ctx.rle_decode(data, [4], [data.shape[0]])  # stream runs past the input
    assert e.value.code == hb._lib.HB_ERR_INVALID
    assert ctx.rle_encode(np.zeros((0, hb.CHUNK_VOLUME), dtype=np.uint16))["bytes"].shape[0] == 0


def test_encode_palette_matches_oracle(ctx):
    """hb_encode_palette = encode_palette + Material::encode (codec.rs:70-75, material.rs:73-99): the game's three
    materials by default, custom palettes with their own group keys; together with the RLE stream it is the byte
    string CubeGrid::encode produces."""
    from test_oracle_codec import GAME_DESCS, py_encode_palette
    got = ctx.encode_palette()
    assert got == O.encode_palette(O.default_palette(), GAME_DESCS)
    custom = [[(0.25, 0.5, 0.75, 1.0)], [(0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0, 0.5)], [(0.1 * k, 0.2, 0.3, 1.0) for k in range(8)],
              [(0.9, 0.9, 0.1, 1.0)]]
    descs = [("a", "", 0x15, False, -2.5), ("grp", "k" * 200, 0x02, True, 0.0), ("herbolution", "sand", 0x3F, True, 0.5),
             ("x" * 255, "y", 0x00, False, 1e9)]
    try:
        ctx.set_materials(custom)
        want = O.encode_palette(O.make_palette(custom), descs)
        assert ctx.encode_palette(descs) == want == py_encode_palette(custom, descs)
        with pytest.raises(hb.HerbGpuError):  # no default descriptions for a 4-material palette
            ctx.encode_palette()
        # CubeGrid::encode of one chunk = palette bytes + its record stream
        ids = np.zeros((1, 32768), dtype=np.uint16)
        ids[0, :5000] = 2
        enc = ctx.rle_encode(ids)
        stream = bytes(enc["bytes"][int(enc["offsets"][0]):int(enc["offsets"][0]) + int(enc["sizes"][0])])
        assert want + stream == want + O.rle_encode(ids[0]).tobytes()
    finally:
        ctx.set_materials([[(0.5, 0.5, 0.5, 1.0), (0.6, 0.6, 0.6, 1.0), (0.7, 0.7, 0.7, 1.0)],
                           [(0.4, 0.3, 0.2, 1.0), (0.5, 0.4, 0.3, 1.0), (0.6, 0.5, 0.4, 1.0)],
                           [(0.1, 0.8, 0.1, 1.0), (0.2, 0.9, 0.2, 1.0), (0.3, 1.0, 0.3, 1.0)]])
