#!/usr/bin/env python
"""Concurrent pinned host->device copy rate per rank (torchrun), with and without binding the rank
to the NUMA node of its GPU before the pinned buffer is allocated.  Context for the e2e numbers of
bench.py at N > 1: the host side, not the GPUs, bounds them.

    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maveric_slam_b200  # noqa: F401,E402
from maveric_slam_b200 import tracking  # noqa: E402


def rate(dev, world):
    h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    del h, d
    return 5 * (1 << 30) / dt / 1e9


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    r0 = rate(dev, world)
    info = tracking.bind_to_gpu_numa_node(local)
    r1 = rate(dev, world)
    t = torch.tensor([r0, r1], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(out, t)
    else:
        out = [t]
    print("rank %d gpu %d numa %s: unbound %.1f GB/s, bound %.1f GB/s" % (rank, local, info, r0, r1), flush=True)
    if rank == 0:
        a = torch.stack(out).cpu()
        print("sum over %d ranks: unbound %.1f GB/s, bound %.1f GB/s" % (world, a[:, 0].sum(), a[:, 1].sum()), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
