"""BASELINE config 5: fused gen+mesh of a 1024x1024-chunk-column region (x cy in [-3,3) = 6 291 456 chunks) sharded
across the GPUs of one box by x-slabs, meshes returned through pinned async copies.

    python tools/run_c5.py [--columns 1024]                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_c5.py

Strong scaling (total work fixed).  Reports, per run: end-to-end chunks/s (every Instance3d landed in pinned host
memory and handed to a sink) and kernel-only chunks/s (device-resident, CUDA events), max over ranks.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import herbolution_b200 as hb  # noqa: E402
from herbolution_b200 import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=1024)
    ap.add_argument("--plan-width", type=int, default=64, help="ix per device-resident plan in the kernel-only leg")
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.columns
    region = hb.Region(-n // 2, -n // 2, n, n, -3, 6)
    lo, hi = sharding.slab_bounds(region.ncx, world, rank)
    ctx = hb.Context(local, seed=0)
    stats = {"chunks": 0, "quads": 0, "bytes": 0}

    def sink(first, off, cnt, inst, ids):
        stats["chunks"] += len(cnt)
        stats["quads"] += int(cnt.sum())
        stats["bytes"] += inst.nbytes

    # warm-up (allocates the pinned staging) on a thin slab, then the timed end-to-end run
    ctx.generate_mesh_region(region, sink=None, ix_begin=lo, ix_end=min(hi, lo + 8))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = ctx.generate_mesh_region(region, sink=sink, ix_begin=lo, ix_end=hi)
    torch.cuda.synchronize()
    t_e2e = sharding.reduce_max(time.perf_counter() - t0, device="cuda")
    assert tot["chunks"] == stats["chunks"] and tot["quads"] == stats["quads"]

    # kernel-only: device-resident plans of plan-width ix each, one after the other
    st = torch.cuda.Stream()
    ms = 0.0
    for a0 in range(lo, hi, a.plan_width):
        plan = ctx.region_plan(region, ix_begin=a0, ix_end=min(hi, a0 + a.plan_width))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        plan.enqueue(hb.STAGE_ALL_MESH, st.cuda_stream)
        e1.record(st)
        st.synchronize()
        ms += e0.elapsed_time(e1)
        plan.close()
    t_k = sharding.reduce_max(ms * 1e-3, device="cuda")
    chunks = sharding.reduce_sum(float(stats["chunks"]), device="cuda")
    quads = sharding.reduce_sum(float(stats["quads"]), device="cuda")
    if rank == 0:
        print(json.dumps({"config": f"C5: {n}x{n} columns x 6 = {int(chunks)} chunks, x-slabs over {world} GPU(s)",
                          "n_gpus": world, "quads": int(quads), "instance_GB": quads * 100 / 1e9,
                          "e2e_s": t_e2e, "e2e_chunks_per_s": chunks / t_e2e, "e2e_GBps_per_gpu": quads * 100 / 1e9 / t_e2e / world,
                          "kernel_s": t_k, "kernel_chunks_per_s": chunks / t_k}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
