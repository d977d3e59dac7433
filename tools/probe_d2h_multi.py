"""Does pinned-buffer NUMA placement explain the ~100 GB/s aggregate device->host ceiling seen beyond 2 GPUs?

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/probe_d2h_multi.py

Every rank copies 2 GiB device->pinned host 6 times concurrently with the others, (a) with the pinned buffer allocated
wherever the process happens to run, (b) after pinning the process to the CPUs of its GPU's NUMA node (read from
sysfs) before allocating.  Rank 0 prints `nvidia-smi topo -m` and the per-rank / aggregate GB/s.  Plumbing probe."""
import glob
import os
import subprocess
import time

import torch
import torch.distributed as dist


def gpu_numa_node(index: int) -> int:
    bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
    try:
        out = subprocess.check_output(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip()
        bus = out.lower().replace("00000000:", "0000:")
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            return int(f.read())
    except Exception:
        return -1


def node_cpus(node: int):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            s = f.read().strip()
    except Exception:
        return None
    cpus = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def measure(local):
    n = 2 << 30
    d = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local}")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(6):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dist.barrier()
    del h
    return 6 * n / dt / 1e9


def measure_slabs(local, slab_mb=384, slabs=46, bufs=2):
    """The region streamer's pattern: `slabs` copies of slab_mb MiB, alternating between `bufs` pinned buffers."""
    n = slab_mb << 20
    d = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{local}") for _ in range(bufs)]
    h = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(bufs)]
    for b in range(bufs):
        h[b].copy_(d[b], non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for s in range(slabs):
        h[s % bufs].copy_(d[s % bufs], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dist.barrier()
    return slabs * n / dt / 1e9


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    if rank == 0:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
        print("numa nodes:", sorted(glob.glob("/sys/devices/system/node/node*")), "affinity", len(os.sched_getaffinity(0)))
    node = gpu_numa_node(local)
    res = {}
    res["default"] = measure(local)
    cpus = node_cpus(node) if node >= 0 else None
    allowed = os.sched_getaffinity(0)
    if cpus and set(cpus) & allowed:
        os.sched_setaffinity(0, set(cpus) & allowed)
        res["bound"] = measure(local)
    else:
        res["bound"] = None
    res["slabs384x2"] = measure_slabs(local, 384, 46, 2)
    res["slabs384x1"] = measure_slabs(local, 384, 46, 1)
    res["slabs2048x2"] = measure_slabs(local, 2048, 8, 2)
    res["slabs64x2"] = measure_slabs(local, 64, 256, 2)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, node, res))
    if rank == 0:
        for r, nd, rs in gathered:
            print(f"rank {r} gpu-numa {nd} default {rs['default']:.1f} GB/s bound {rs['bound'] if rs['bound'] is None else round(rs['bound'], 1)}")
        print("aggregate default", round(sum(g[2]["default"] for g in gathered), 1), "GB/s; bound",
              round(sum((g[2]["bound"] or 0) for g in gathered), 1), "GB/s")
        for k in ("slabs384x2", "slabs384x1", "slabs2048x2", "slabs64x2"):
            print(k, [round(g[2][k], 1) for g in gathered], "aggregate", round(sum(g[2][k] for g in gathered), 1), "GB/s")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
