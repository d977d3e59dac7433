"""SASS opcode histogram of every kernel in libherbgpu.so (cuobjdump -sass), written to profiles/.

    python tools/sass_hist.py [out.txt]

Per kernel: instruction count, the packed-f32x2 arithmetic (FFMA2 / FADD2 / FMUL2), scalar FFMA / FADD / FMUL, the
conversion pipe (F2I / I2F / FRND), shared/global memory operations, the TMA bulk copy (UBLKCP) and the top opcodes.
There is no tensor-core (UTCMMA / HMMA) work on this path and the histogram shows none."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "herbolution_b200", "libherbgpu.so")


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:  # noqa: BLE001
        return {n: n for n in names}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_histogram.txt")
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    names = [f.split("\n", 1)[0].strip() for f in funcs]
    dm = demangle(names)
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()
    groups = [("FFMA2", r"^FFMA2"), ("FADD2", r"^FADD2"), ("FMUL2", r"^FMUL2"), ("FFMA", r"^FFMA(?!2)"), ("FADD", r"^FADD(?!2)"),
              ("FMUL", r"^FMUL(?!2)"), ("F2I/I2F/FRND", r"^(F2I|I2F|FRND|I2FP|F2IP)"), ("LDS", r"^LDS"), ("STS", r"^STS"),
              ("LDG", r"^(LDG|LD\.)"), ("STG", r"^(STG|ST\.)"), ("UBLKCP", r"^UBLKCP"), ("ATOM/RED", r"^(ATOM|RED|ATOMG|ATOMS)"),
              ("BAR", r"^BAR"), ("tensor (HMMA/UTC*MMA)", r"^(HMMA|IMMA|UTCHMMA|UTCQMMA|UTCIMMA|UTCMMA)")]
    lines = [f"# SASS opcode histogram of herbolution_b200/libherbgpu.so (sha256 {sha[:16]}...), sm_100a, cuobjdump -sass",
             f"# {len(funcs)} kernels; columns: total | " + " | ".join(g for g, _ in groups) + " | top opcodes", ""]
    for name, f in sorted(zip(names, funcs), key=lambda t: dm[t[0]]):
        ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", f)
        base = [o.split(".")[0] for o in ops]
        hist = collections.Counter(base)
        short = re.sub(r"hb::\(anonymous namespace\)::|\(anonymous namespace\)::", "", dm[name])
        short = re.sub(r"\(.*", "", short)
        cnt = [sum(v for k, v in hist.items() if re.match(rx, k)) for _, rx in groups]
        top = ", ".join(f"{k} {v}" for k, v in hist.most_common(8))
        lines.append(f"{short}\n    {len(ops)} | " + " | ".join(str(c) for c in cnt) + f" | {top}")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(out_path, len(funcs), "kernels")


if __name__ == "__main__":
    main()
