#!/usr/bin/env python
"""Diagnostic (torchrun, one rank per GPU): device->host bandwidth of a 1 GiB pinned copy per rank, alone and with all
ranks copying at once, with and without binding the rank to the CPUs NVML reports as local to its GPU before the pinned
buffer is allocated (first touch decides the NUMA node).  Explains the end-to-end scaling of bench.py at N = 8."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    info = {"rank": rank, "aff0": sorted(os.sched_getaffinity(0))}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 4)
        cpus = [i for w, m in enumerate(mask) for i in range(w * 64, w * 64 + 64) if (m >> (i - w * 64)) & 1]
        info["gpu_cpus"] = cpus[:8] + ["..."] + cpus[-4:] if len(cpus) > 12 else cpus
    except Exception as e:
        cpus = []
        info["nvml_err"] = repr(e)
    nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.isdir("/sys/devices/system/node") else []
    info["numa_nodes"] = len(nodes)
    n = 1 << 27                                     # doubles = 1 GiB
    src = torch.rand(n, dtype=torch.float64, device="cuda")

    def bw(dst, together):
        torch.cuda.synchronize()
        if together:
            dist.barrier()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return n * 8 / (time.perf_counter() - t0) / 1e9

    dst = torch.empty(n, dtype=torch.float64).pin_memory()
    bw(dst, False)
    res = {}
    for r in range(world):                          # one rank at a time
        dist.barrier()
        if r == rank:
            res["alone"] = bw(dst, False)
    dist.barrier()
    res["together"] = bw(dst, True)
    usable = sorted(set(cpus) & os.sched_getaffinity(0))
    if usable:
        os.sched_setaffinity(0, usable)
        dst2 = torch.empty(n, dtype=torch.float64).pin_memory()
        bw(dst2, False)
        dist.barrier()
        res["together_numa_local"] = bw(dst2, True)
    info.update({k: round(v, 1) for k, v in res.items()})
    out = [None] * world
    dist.all_gather_object(out, info)
    if rank == 0:
        for o in out:
            print(o)
        print("aggregate together GB/s:", round(sum(o["together"] for o in out), 1),
              " numa-local:", round(sum(o.get("together_numa_local", 0) for o in out), 1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
