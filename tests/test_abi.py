"""CPU tests of the drop-in boundary: libherbgpu.so loads, exports every symbol include/herbgpu.h
declares, the POD layouts match the reference's repr(C) types, host-only helpers work, and the product
fails loudly without a GPU (no CPU fallback, no dependency on oracle/)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import herbolution_b200 as hb
from herbolution_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "herbgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n not in ("hb_region_sink", "hb_region_packed_sink")))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 20
    nm = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (hb_[a-z0-9_]+)", nm))
    for name in declared:
        assert name in exported, f"{name} declared in herbgpu.h but not exported"
        assert hasattr(lib, name)
    assert sorted(_lib.PROTOTYPES) == declared, "python binding table out of sync with the header"
    # C ABI only: nothing mangled is exported except what the C++ runtime drags in
    assert not [s for s in re.findall(r"\bT (\S+)", nm) if s.startswith("_ZN2hb")]


def test_pod_layouts_match_reference_types():
    assert hb.INSTANCE_DTYPE.itemsize == 100      # Instance3d, client/src/video/world/vertex.rs:41-51
    assert hb.INSTANCE_DTYPE.fields["rgba"][1] == 64 and hb.INSTANCE_DTYPE.fields["light"][1] == 80
    assert hb.INSTANCE_DTYPE.fields["ao"][1] == 84
    assert hb.CUBE_DTYPE.itemsize == 8            # PaletteCube, server/src/chunk/cube.rs:5-10
    assert hb.CUBE_DTYPE.fields["flags"][1] == 4
    assert hb.CHUNK_PT_DTYPE.itemsize == 12       # ChunkPt = Vec3<i32>
    assert C.sizeof(hb.Region) == 24


def test_version_and_error_channel():
    lib = _lib.load()
    assert b"sm_100a" in lib.hb_version()
    rc = lib.hb_create(0, None)
    assert rc == _lib.HB_ERR_INVALID and b"null" in lib.hb_last_error()


def test_build_neighbor_index_host_helper():
    """hb_build_neighbor_index is pure host code (no GPU): shell order of create_chunk_shell
    (client/src/world/chunk.rs:158-174), centre = 13, -1 = absent."""
    lib = _lib.load()
    pts = hb.Region(0, 0, 2, 2, 0, 2).chunk_points()
    out = np.empty((len(pts), 27), dtype=np.int32)
    assert lib.hb_build_neighbor_index(pts.ctypes.data, len(pts), out.ctypes.data) == 0
    lut = {(int(p["x"]), int(p["y"]), int(p["z"])): i for i, p in enumerate(pts)}
    for i, p in enumerate(pts):
        k = 0
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    want = lut.get((int(p["x"]) + dx, int(p["y"]) + dy, int(p["z"]) + dz), -1)
                    assert out[i, k] == want
                    k += 1
        assert out[i, 13] == i
    assert lib.hb_build_neighbor_index(None, 0, None) == 0


def test_region_chunk_points_order():
    r = hb.Region(-1, 5, 2, 3, -2, 2)
    pts = r.chunk_points()
    assert len(pts) == 12
    for ix in range(2):
        for iz in range(3):
            for iy in range(2):
                p = pts[(ix * 3 + iz) * 2 + iy]
                assert (p["x"], p["y"], p["z"]) == (-1 + ix, -2 + iy, 5 + iz)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(hb.HerbGpuError) as e:
        hb.Context()
    assert e.value.code == _lib.HB_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under herbolution_b200/ or include/ may import, link
    or execute it."""
    bad = []
    for base in ("herbolution_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, fn)).read()
                    for m in re.finditer(r"oracle", txt):
                        line = txt[txt.rfind("\n", 0, m.start()) + 1:txt.find("\n", m.end())]
                        if re.search(r"import|include|CDLL|dlopen|build_oracle|liboracle|hbo_", line):
                            bad.append((fn, line.strip()))
    assert not bad, bad
    ldd = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    assert "oracle" not in ldd
    nm = subprocess.check_output(["nm", "-D", _lib.LIB_PATH], text=True)
    assert "hbo_" not in nm


def test_host_mirror_api_surface():
    """The mirror keeps the reference's names (server/src/generator/mod.rs, client/src/world/chunk.rs)."""
    for name in ("ChunkGenerator", "GenerationParams", "CubeMesh", "ChunkPt", "ChunkMap", "ChunkData",
                 "create_chunk_shell", "generate_mesh"):
        assert hasattr(hb, name)
    assert {"request", "dequeue"} <= set(dir(hb.ChunkGenerator))
    assert {"generate", "get_noise"} <= set(dir(hb.GenerationParams))
    shell = hb.create_chunk_shell({hb.ChunkPt(0, 0, 0): hb.ChunkData(hb.ChunkPt(0, 0, 0), np.zeros(32768, np.uint16))},
                                  hb.ChunkPt(0, 0, 0))
    assert len(shell) == 27 and shell[13] is not None and sum(s is not None for s in shell) == 1


def test_noise_kernel_has_no_unintended_fma_contraction():
    """ptxas (CUDA 12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false; the reference's
    operator overloads never fuse (only the six neg_mul_add sites of simplex_2d do).  csrc/noise.cuh routes every
    product that feeds an add across a packed/scalar boundary; this pins the result in the shipped SASS.  Per simplex
    pair of the fused column (hb_set_fma_mode(1)) there are exactly
      * 6 FFMA2: the three inner neg_mul_add (0.5 - x*x) and the three gradient dot products gx*x + gy*y, whose
        products are exact (components +-1 / +-2), so the fused form is bit-identical to the reference's mul, mul, add;
      * 6 scalar FFMA.SAT: the three outer neg_mul_add (t - y*y) for both samples, carrying the t >= 0 clamp;
      * 13 FMUL2 (16 multiplies of the reference minus the three folded into the dot products);
    plus, per octave-loop pair, ridge's own neg_mul_add (functions/ridge_32.rs:27: one scalar FFMA per lane) and FBM's
    packed `simplex * amp`.  How many copies of the pair the compiler lays down (loop peeling, the two lacunarity
    variants) is its business.  The unfused column (SSE2 / scalar hosts) contains no FMA of any kind.  Every kind exists
    with the conversion-free index path (FAST) and with the exact-conversion fallback."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True)
    funcs = re.split(r"\n\s*Function : ", sass)
    heights = [f for f in funcs if "k_heights" in f.split("\n", 1)[0]]
    # {ColList, ColRegion} x 4 noise kinds x {FAST + fused, exact + fused, exact + unfused (hb_set_fma_mode(0))}
    assert len(heights) == 24
    unfused = [f for f in heights if re.search(r"ELi\dELb[01]ELb0EEEv", f.split("\n", 1)[0])]
    assert len(unfused) == 8
    for f in unfused:  # the SSE2 / scalar column: no fused multiply-add anywhere in the noise
        ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
        assert not any(o.startswith("FFMA") for o in ops), "FMA in the unfused noise column"
    heights = [f for f in heights if f not in unfused]
    # KIND -> (extra scalar FFMA per octave-loop pair: ridge only, extra FMUL2 per octave-loop pair: FBM's `* amp`)
    want = {0: (0, 1), 1: (2, 0), 2: (0, 0), 3: (0, 0)}
    conv = {}
    for f in heights:
        m = re.search(r"ELi(\d)ELb([01])ELb1EEEv", f.split("\n", 1)[0])
        kind, fast = int(m.group(1)), m.group(2) == "1"
        ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
        packed = sum(o.startswith("FFMA2") for o in ops)
        sat = sum(o.startswith("FFMA.SAT") for o in ops)
        scalar = sum(o.startswith("FFMA") and not o.startswith("FFMA2") for o in ops) - sat
        assert packed % 6 == 0 and packed >= 12, f"kind {kind}: FFMA2 count {packed} is not 6 per simplex pair"
        pairs, firsts = packed // 6, 2  # two first-octave evaluations (k = 0, 2) precede the loops
        assert sat == 6 * pairs, f"kind {kind}: {sat} FFMA.SAT for {pairs} simplex pairs"
        assert scalar == want[kind][0] * (pairs - firsts), f"kind {kind}: scalar FFMA count {scalar}, check for contraction"
        fmul2 = sum(o.startswith("FMUL2") for o in ops)
        # a contraction would turn one FMUL2 per site into an extra FFMA2: fewer than 13 FMUL2 per counted pair
        assert 13 * pairs <= fmul2 <= 13 * pairs + want[kind][1] * (pairs - firsts), f"kind {kind}: {fmul2} FMUL2 for {pairs} simplex pairs"
        assert sum(o.startswith("FMUL2") for o in ops) > 0 and sum(o.startswith("FADD2") for o in ops) > 0  # Blackwell packed f32x2
        conv[(f.split("\n", 1)[0].split("k_heights")[1][:16], kind, fast)] = (sum(o.startswith(("F2I", "I2F")) for o in ops), pairs)
    # the FAST index path stays off the quarter-rate conversion pipe: each simplex pair of the exact path converts
    # 4 floors to integers and 2 sums back (6 F2I / I2F); the FAST path has none (what is left sets up coordinates)
    for (name, kind, fast), (n_conv, pairs) in conv.items():
        if fast:
            n_exact, pairs_exact = conv[(name, kind, False)]
            assert n_exact - 6 * pairs_exact >= n_conv - 2 and n_conv <= 20, f"conversions crept back into the FAST noise path ({n_conv})"
    # the whole library is compiled without FMA contraction in the 3-D noise path too (reference: no FMAs there)
    # (reference: no FMAs there except ridge_3d's neg_mul_add, functions/ridge_32.rs:43)
    n3 = [f for f in funcs if "k_noise3d" in f.split("\n", 1)[0]]
    assert n3 and len(re.findall(r"\bFFMA\b", n3[0])) == 1


def test_mesher_kernels_use_the_hardware_paths_the_design_claims():
    """k_pack tests ids with the DPX packed 16-bit min/max (VIMNMX3.U16x2), k_emit_ids fetches the next
    chunk's occupancy words with cp.async (LDGSTS) and both emit kernels write their records with TMA bulk stores
    (UBLKCP).  Pinned in the shipped SASS so a refactor cannot silently fall back to the slow forms."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True)
    funcs = {f.split("\n", 1)[0]: f for f in re.split(r"\n\s*Function : ", sass)[1:]}

    def one(pred):
        hits = [body for name, body in funcs.items() if pred(name)]
        assert hits, "kernel not found in the library"
        return hits

    pack = one(lambda n: re.search(r"\d+k_packE", n))[0]
    assert len(re.findall(r"VIMNMX3\.U16x2", pack)) >= 6 * 16  # 6 per 8-id unit, 16 units per thread
    emit = one(lambda n: "k_emit_ids" in n)[0]
    assert "LDGSTS" in emit and "UBLKCP" in emit
    for body in one(lambda n: "k_mesh_region" in n and "Lb1ELb0E" in n):  # EMIT, 100-byte records
        assert "UBLKCP" in body
    assert one(lambda n: "k_count_border" in n)
