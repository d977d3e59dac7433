"""Recipe for the CPU oracle shared object (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY.

-ffp-contract=off is required (the reference's operator overloads never fuse); -mavx2 -mfma makes
fmaf() a real FMA instruction, matching the AVX2 column of the reference that an x86 host runs.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SRC = os.path.join(HERE, "hb_oracle.c")
ORACLE_LIB = os.path.join(HERE, "_build", "liboracle.so")
ORACLE_FLAGS = ["-O2", "-std=gnu11", "-ffp-contract=off", "-mavx2", "-mfma", "-fPIC", "-shared", "-pthread",
                "-Wall", "-Wextra"]


def build_oracle(force: bool = False) -> str:
    deps = [ORACLE_SRC, os.path.join(HERE, "hb_oracle.h")]
    if not all(os.path.exists(d) for d in deps):
        if os.path.exists(ORACLE_LIB):
            return ORACLE_LIB  # prebuilt copy travelling with the snapshot
        raise RuntimeError("oracle sources missing")
    if not force and os.path.exists(ORACLE_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(ORACLE_LIB) for d in deps):
        return ORACLE_LIB
    os.makedirs(os.path.dirname(ORACLE_LIB), exist_ok=True)
    cmd = ["gcc", *ORACLE_FLAGS, "-o", ORACLE_LIB, ORACLE_SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building the oracle")
    return ORACLE_LIB


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
