#!/usr/bin/env python
"""bench.py -- chunks/sec generated+meshed (BASELINE.json metric) on N B200s of one node.

A step = one pass of the fused generate+mesh hot path over one region of chunk columns per GPU:
BASELINE config 4, "fused gen+mesh for a 256x256-chunk region" (256x256 columns x cy in [-3,3) =
393 216 chunks).  With N GPUs every rank takes its own 256x256 slab of a (256*N)x256 region
(x-slab partition, no collective on the data path) -> weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--columns C]

  value    : kernel-only throughput, outputs resident in HBM (CUDA events on the launching stream,
             max over ranks).
  e2e      : same metric through the C ABI's host entry point hb_generate_mesh_region, finished
             Instance3d buffers landed in pinned host memory (device->host copies in the timed region).
  roofline : the dominant kernel (k_mesh_region<EMIT>: 100 B per emitted quad) vs measured HBM peak.
  cpu_baseline / --impl reference : the CPU restatement of the reference path (oracle/, "port" --
             the Rust reference cannot be built here) on a bounded sample of the same region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SEED = 0
CY0, NCY = -3, 6
METRIC = "chunks/sec generated+meshed"
UNIT = "chunks/s"


def region_for_rank(columns: int, rank: int, world: int):
    # global region: x in [-columns*world/2, +columns*world/2), z in [-columns/2, columns/2)
    from herbolution_b200.sharding import weak_scaling_region
    r = weak_scaling_region(columns, rank, world, CY0, NCY)
    return (r.cx0, r.cz0, r.ncx, r.ncz, r.cy0, r.ncy)


def cpu_sample_region(columns: int):
    s = min(64, columns)  # 64 x 64 columns x 6 layers = 24 576 chunks: a few seconds on 16 cores
    return (-(columns // 2), -(columns // 2), s, s, CY0, NCY)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu(sample, threads: int, reps: int):
    """The reference's CPU path as restated in oracle/ (literal set() + cull_shared_faces + generate_mesh)."""
    import oracle as O
    w = O.world(SEED)
    n = sample[2] * sample[3] * sample[5]
    times = []
    quads = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        r = O.region_pipeline(w, sample, literal=True, threads=threads, want_instances=False)
        times.append(time.perf_counter() - t0)
        quads = r["total"]
    return n, quads, times


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region.

    nvidia-smi is started before the warm-up (it needs ~100 ms to come up) and polls every 20 ms; every line is
    time-stamped on arrival.  stop() reports the samples that arrived between begin() and end() -- the timed
    region -- and, if that region was too short to catch one, every sample since begin() up to stop() ("window"
    says which)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.t = None
        self.t_begin = self.t_end = None

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_stop = time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        t1 = self.t_end if self.t_end is not None else t_stop
        # a line printed at time t describes the GPU a few ms earlier: accept arrivals up to 25 ms after end()
        timed = [ln for ts, ln in self.lines if t0 <= ts <= t1 + 0.025]
        window = "timed region"
        if not timed:
            timed = [ln for ts, ln in self.lines if t0 <= ts <= t_stop]
            window = "begin() .. stop() (timed region shorter than one sample period)"
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in timed:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=256, help="region is columns x columns chunk columns per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (rank 0, N=1 only anyway)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-API leg (profiling runs only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cols = args.columns
    workload = (f"BASELINE config 4: fused gen+mesh, {cols}x{cols} chunk columns x cy[{CY0},{CY0 + NCY}) = "
                f"{cols * cols * NCY} chunks per GPU, seed {SEED}; region border = absent chunks")
    config = {"workload": workload, "columns_per_gpu": cols * cols, "chunks_per_gpu": cols * cols * NCY,
              "parallelism": f"x-slab partition over {world} GPU(s), no data-path collective",
              "l2": "no explicit flush: each step writes ~17 GB of instances and re-reads 268 MB of height tiles "
                    "(both >> 126 MB L2)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = host_threads()
        sample = cpu_sample_region(cols)
        for _ in range(min(warmup, 1)):
            run_cpu(sample, threads, 1)
        n, quads, times = run_cpu(sample, threads, steps)
        total_t = sum(times)
        val = n * steps / total_t
        sample_txt = (f"{sample[2]}x{sample[3]} columns x {NCY} = {n} chunks per step (sub-region of the workload), "
                      f"literal CubeMesh::set + cull_shared_faces + generate_mesh, {quads} quads")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": 1e3 * total_t / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "CPU restatement of the reference path (oracle/hb_oracle.c); the Rust reference cannot be "
                        "built in this image (no cargo/rustc, nightly-only, un-vendored deps)"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import numpy as np
    import torch

    import herbolution_b200 as hb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libherbgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from herbolution_b200.sharding import reduce_max, reduce_sum

    def max_over_ranks(x: float) -> float:
        return reduce_max(x, device="cuda")

    def sum_over_ranks(x: float) -> float:
        return reduce_sum(x, device="cuda")

    ctx = hb.Context(device=local_rank, seed=SEED)
    reg = hb.Region(*region_for_rank(cols, rank, world))
    plan = ctx.region_plan(reg)  # sizes the instance buffer with one synchronous count pass
    # a dedicated non-default stream: kernels are launched on it and the CUDA events are recorded on it
    tstream = torch.cuda.Stream(device=local_rank)
    stream = tstream.cuda_stream
    assert stream != 0
    ALL = hb.STAGE_ALL_MESH
    launches_per_step = plan.launches(ALL)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(warmup, 3)):
        plan.enqueue(ALL, stream)
    res = plan.result(stream)
    n_chunks, quads, span = res["n_chunks"], res["quads"], res["span"]

    # The timed region: K steps, each enqueued stage by stage with a CUDA event between the stages (all on the
    # launching stream), so the per-kernel durations behind the roofline come from the timed steps themselves.
    stage_names = [("heights", hb.STAGE_HEIGHTS), ("count", hb.STAGE_COUNT), ("scan", hb.STAGE_SCAN), ("emit", hb.STAGE_EMIT)]
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stage_names) + 1)] for _ in range(steps)]
    barrier()
    sampler.begin()
    for r in range(steps):
        evs[r][0].record(tstream)
        for i, (_, st) in enumerate(stage_names):
            plan.enqueue(st, stream)
            evs[r][i + 1].record(tstream)
    barrier()
    sampler.end()
    ev0, ev1 = evs[0][0], evs[-1][-1]
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    total_chunks = sum_over_ranks(float(n_chunks))
    value = total_chunks * steps / (ms_total * 1e-3)
    clocks = sampler.stop()
    reps = steps
    stage_ms = {nm: statistics.mean(evs[r][i].elapsed_time(evs[r][i + 1]) for r in range(reps))
                for i, (nm, _) in enumerate(stage_names)}
    peak, peak_src = measured_peaks()
    emit_bytes = 100.0 * quads  # fused path writes 100 B per emitted quad, no reads beyond coords
    emit_gbs = emit_bytes / (stage_ms["emit"] * 1e-3) / 1e9
    flops_per_column = 333824.0  # 1024 samples x 326 flops
    ncols_noise = reg.ncx * reg.ncz
    noise_tflops = flops_per_column * ncols_noise / (stage_ms["heights"] * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one `ncu --set full` capture of this exact
    # workload (profiles/r1_ncu_full_summary.csv: 0.274 GB read + 17.480 GB written per launch)
    traffic = 17.755e9 if (cols == 256 and world == 1) else None
    roofline = {"kernel": "k_mesh_region<EMIT=true>", "bound": "hbm", "achieved": emit_gbs, "peak": peak, "unit": "GB/s",
                "frac": emit_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": emit_bytes, "avg_launch_ms": stage_ms["emit"],
                "share_of_step": stage_ms["emit"] / sum(stage_ms.values())}
    kernels = {"stage_ms": stage_ms,
               "noise": {"kernel": "k_heights", "bound": "fp32", "achieved_tflops": noise_tflops, "peak_tflops": 74.4,
                         "frac": noise_tflops / 74.4, "flops_per_column": flops_per_column}}

    # ---- end to end through the host entry point: instances land in pinned host memory
    e2e = None
    stats = {"span": 0, "chunks": 0}

    def sink(first, off, cnt, inst, ids):
        stats["span"] += len(inst)
        stats["chunks"] += len(cnt)

    e2e_steps = max(1, min(args.e2e_steps, steps))
    e2e_alt = None
    if not args.no_e2e:
        host_thr = max(1, host_threads() // max(1, world))
        ctx.set_host_threads(host_thr)

        def run_e2e(transport):
            ctx.set_region_transport(transport)
            ctx.generate_mesh_region(reg, sink=sink)  # warm-up: staging buffers are allocated and touched here
            stats["span"] = stats["chunks"] = 0
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ctx.generate_mesh_region(reg, sink=sink)
            torch.cuda.synchronize()
            t = max_over_ranks(time.perf_counter() - t0)
            barrier()
            return t, stats["span"] / e2e_steps, stats["chunks"] / e2e_steps

        # headline: packed transport (4 B/quad over the link, Instance3d records rebuilt by the host threads)
        t_e2e, span_s, chunks_s = run_e2e(hb.TRANSPORT_PACKED)
        e2e = {"value": total_chunks * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 24,
               "d2h_bytes_per_step": int(span_s * 4 + chunks_s * 12), "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps,
               "api": "hb_generate_mesh_region (C ABI, host sink; HB_TRANSPORT_PACKED: 4 B/quad D2H into pinned memory, "
                      "100 B Instance3d records rebuilt by host threads with non-temporal stores)",
               "host_threads": host_thr, "instance_bytes_delivered_per_step": int(span_s * 100),
               "host_write_gbs": span_s * 100 * world / (t_e2e / e2e_steps) / 1e9,
               "note": "input is the 24-byte region descriptor; the sink receives every Instance3d (100 B/quad) + offsets/counts"}
        # the round-1 path beside it: the device writes the 100-byte records and they cross the link as they are
        t_raw, span_r, chunks_r = run_e2e(hb.TRANSPORT_INSTANCES)
        e2e_alt = {"transport": "HB_TRANSPORT_INSTANCES", "value": total_chunks * e2e_steps / t_raw, "unit": UNIT,
                   "d2h_bytes_per_step": int(span_r * 100 + chunks_r * 12), "ms_per_step": 1e3 * t_raw / e2e_steps,
                   "d2h_gbs_per_gpu": (span_r * 100 + chunks_r * 12) / (t_raw / e2e_steps) / 1e9}
        ctx.set_region_transport(hb.TRANSPORT_PACKED)

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        sample = cpu_sample_region(cols)
        n, q, times = run_cpu(sample, threads, 2)
        best = min(times)
        cpu = {"value": n / best, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{sample[2]}x{sample[3]} columns x {NCY} = {n} chunks (sub-region of the workload), oracle "
                         f"literal path, best of 2 passes, {best:.2f} s"}

    # columns/s and the split of chunks by what they contribute, so that trivial (all-air / buried)
    # layers cannot inflate the headline.  From this rank's per-chunk quad counts (device -> host, outside timing).
    split = None
    if rank == 0:
        import numpy as np
        per = ctx.download(res["d_counts"], np.uint32, n_chunks).astype(np.int64).reshape(-1, NCY)
        with_quads = int((per > 0).sum())
        slab_only = int((per[:, 0] == 1024).sum())  # region-bottom chunks whose only faces are their 1024 Down faces
        split = {"columns_per_s": value / NCY, "chunks_with_quads": with_quads, "chunks_without_quads": int(per.size - with_quads),
                 "region_bottom_slab_only_chunks": slab_only, "surface_chunks": with_quads - slab_only,
                 "surface_chunks_per_s": (with_quads - slab_only) / per.size * value,
                 "layers": [{"cy": CY0 + i, "chunks_with_quads": int((per[:, i] > 0).sum()), "quads": int(per[:, i].sum())}
                            for i in range(NCY)],
                 "note": "per GPU (rank 0's slab); rates scale the whole-job value by this rank's fractions"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
                "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches_per_step * steps, "roofline": roofline, "cpu_baseline": cpu,
                "e2e_instances_transport": e2e_alt, "kernels": kernels, "quads_per_step_per_gpu": quads, "instance_bytes_per_step_per_gpu": span * 100,
                "split": split}
        print(json.dumps(line))
    plan.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
