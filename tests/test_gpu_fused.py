"""HB_STAGE_FUSED (one kernel: noise + count + chained-scan offsets + emit) must reproduce the staged pipeline
(HEIGHTS | COUNT | SCAN | EMIT) bit for bit: counts, offsets, totals, every instance byte, and the heights tiles
the FILL stage reads afterwards.  Through the C ABI."""
import numpy as np
import pytest

import herbolution_b200 as hb
import oracle as O

pytestmark = pytest.mark.gpu


def run_plan(ctx, region, stages, want_ids=False, capacity=0, **kw):
    plan = ctx.region_plan(region, quad_capacity=capacity, want_ids=want_ids, **kw)
    plan.enqueue(stages)
    res = plan.result()
    n = res["n_chunks"]
    out = {"n": n, "quads": res["quads"], "span": res["span"],
           "off": ctx.download(res["d_offsets"], np.uint64, n), "cnt": ctx.download(res["d_counts"], np.uint32, n),
           "inst": ctx.download(res["d_instances"], hb.INSTANCE_DTYPE, res["span"]).tobytes()}
    if want_ids:
        out["ids"] = ctx.download(res["d_ids"], np.uint16, n * hb.CHUNK_VOLUME).tobytes()
    plan.close()
    return out


@pytest.mark.parametrize("seed,reg", [(0, (-4, -3, 9, 7, -3, 6)), (5, (10, -20, 5, 3, -1, 2)), (0, (0, 0, 1, 1, 0, 1)),
                                      (5, (3, 3, 2, 2, -16, 32)), (0, (0, 0, 2, 2, 3, 2)), (5, (0, 0, 2, 2, -2, 1)),
                                      (12345, (-40, 17, 24, 20, -3, 6))])
def test_fused_equals_staged(ctx, seed, reg):
    ctx.set_world(seed)
    region = hb.Region(*reg)
    a = run_plan(ctx, region, hb.STAGE_ALL_MESH | hb.STAGE_FILL, want_ids=True)
    b = run_plan(ctx, region, hb.STAGE_FUSED | hb.STAGE_FILL, want_ids=True)
    assert (a["n"], a["quads"], a["span"]) == (b["n"], b["quads"], b["span"])
    assert (a["cnt"] == b["cnt"]).all() and (a["off"] == b["off"]).all()
    assert a["inst"] == b["inst"], "instances differ"
    assert a["ids"] == b["ids"], "FILL after FUSED reads different heights"
    ctx.set_world(0)


def test_fused_vs_oracle_and_slabs(ctx):
    """Directly against the oracle, and as x-slabs (apron recomputed at the slab seams)."""
    ctx.set_world(5)
    reg = (-3, -3, 6, 6, -3, 6)
    region = hb.Region(*reg)
    ref = O.region_pipeline(O.world(5), reg, literal=True, threads=8)
    want = ref["instances"]
    woff = np.concatenate([[0], np.cumsum(ref["counts"].astype(np.int64))])
    per_ix = region.ncz * region.ncy
    for (x0, x1) in [(0, 6), (0, 2), (2, 3), (3, 6)]:
        r = run_plan(ctx, region, hb.STAGE_FUSED, ix_begin=x0, ix_end=x1)
        inst = np.frombuffer(r["inst"], dtype=hb.INSTANCE_DTYPE)
        for k in range(r["n"]):
            g = x0 * per_ix + k
            got = inst[int(r["off"][k]):int(r["off"][k]) + int(r["cnt"][k])]
            assert got.tobytes() == want[woff[g]:woff[g + 1]].tobytes(), f"slab {x0}:{x1} chunk {k}"
    ctx.set_world(0)


def test_fused_fallbacks_and_capacity(ctx):
    """More than 32 layers or a non-FBM noise kind take the staged kernels (same results); a too small instance
    buffer is reported, not overrun."""
    ctx.set_world(5)
    tall = hb.Region(3, 3, 1, 2, -20, 40)
    a = run_plan(ctx, tall, hb.STAGE_ALL_MESH)
    b = run_plan(ctx, tall, hb.STAGE_FUSED)
    assert a["inst"] == b["inst"] and (a["off"] == b["off"]).all()
    ctx.set_noise_kind(hb._lib.NOISE_RIDGE)
    region = hb.Region(-2, -2, 4, 4, -3, 6)
    a = run_plan(ctx, region, hb.STAGE_ALL_MESH)
    b = run_plan(ctx, region, hb.STAGE_FUSED)
    assert a["inst"] == b["inst"] and a["quads"] == b["quads"]
    ctx.set_noise_kind(hb._lib.NOISE_FBM)
    plan = ctx.region_plan(region, quad_capacity=64)
    plan.enqueue(hb.STAGE_FUSED)
    with pytest.raises(hb.HerbGpuError) as e:
        plan.result()
    assert e.value.code == hb._lib.HB_ERR_CAPACITY
    plan.close()
    # custom world parameters outside the conversion-free index range use the exact-conversion variant
    ctx.set_world(7, freq=0.05, octaves=8, lacunarity=3.0, gain=0.5)
    a = run_plan(ctx, region, hb.STAGE_ALL_MESH)
    b = run_plan(ctx, region, hb.STAGE_FUSED)
    assert a["inst"] == b["inst"] and (a["cnt"] == b["cnt"]).all()
    ctx.set_world(0)


def test_fused_repeatable_at_size(ctx):
    """128x128 columns: thousands of columns chained through the look-back by hundreds of CTAs; two runs and the
    staged pipeline agree byte for byte."""
    ctx.set_world(0)
    region = hb.Region(-64, -64, 128, 128, -3, 6)
    a = run_plan(ctx, region, hb.STAGE_ALL_MESH)
    b = run_plan(ctx, region, hb.STAGE_FUSED, capacity=a["span"])
    c = run_plan(ctx, region, hb.STAGE_FUSED, capacity=a["span"])
    assert (a["cnt"] == b["cnt"]).all() and (a["off"] == b["off"]).all() and (a["quads"], a["span"]) == (b["quads"], b["span"])
    assert a["inst"] == b["inst"] == c["inst"]
