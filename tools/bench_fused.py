"""Staged (HEIGHTS|COUNT|SCAN|EMIT) vs single-launch (HB_STAGE_FUSED) region pipeline on config 4, kernel-only.
    python tools/bench_fused.py [--columns 256] [--steps 10]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import herbolution_b200 as hb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    n = a.columns
    ctx = hb.Context(0, seed=0)
    region = hb.Region(-n // 2, -n // 2, n, n, -3, 6)
    plan = ctx.region_plan(region)
    st = torch.cuda.Stream()
    out = {}
    for name, stages in (("staged", hb.STAGE_ALL_MESH), ("fused", hb.STAGE_FUSED), ("staged2", hb.STAGE_ALL_MESH), ("fused2", hb.STAGE_FUSED)):
        for _ in range(3):
            plan.enqueue(stages, st.cuda_stream)
        res = plan.result(st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.steps):
            plan.enqueue(stages, st.cuda_stream)
        e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[name] = {"ms_per_step": ms, "chunks_per_s": res["n_chunks"] / ms * 1e3, "quads": res["quads"],
                     "GBps_instances": res["quads"] * 100 / ms / 1e6}
    print(json.dumps(out))
    plan.close()
    ctx.close()


if __name__ == "__main__":
    main()
