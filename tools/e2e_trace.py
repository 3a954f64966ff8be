#!/usr/bin/env python
"""Where a rank's end-to-end step goes at N GPUs: every rank runs its shard of the bench sequence through
mv_track_sequence_host (pinned host buffers) concurrently; rank 0 runs one more step with MV_HOST_TRACE=1, which
makes the library print the device timestamps of every stage of every chunk (DMA begin/end, detector, row gather,
matcher, pose) to stderr.  Prints per-rank step times and rank 0's stage summary.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_trace.py 2> trace.txt
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maveric_slam_b200  # noqa: E402,F401
from maveric_slam_b200 import synth, tracking  # noqa: E402

N_FRAMES = 4541
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
node = tracking.bind_to_gpu_numa_node(local)
first, count, per = tracking.shard_pairs(N_FRAMES - 1, world, rank)
tr = tracking.Tracker(local)
params = tracking.kitti_track_params(first_pair=first)
offs = synth.default_offsets(N_FRAMES, 0)
semi, desc, depth = tr.synth_frames(0, 47, 155, first, offs[first:first + count + 1])
scale = torch.full((count + 1,), float(synth.SEMI_SCALE), device=dev)
h = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (semi, scale, desc, depth)]
for a, b in zip(h, (semi, scale, desc, depth)):
    a.copy_(b)
torch.cuda.synchronize()
out = np.zeros(count, tracking.PAIR_RESULT_DTYPE)
times = []
for i in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, up, _ = tr.track_sequence_host(params, h[0], h[1], h[2], h[3], out=out)
    times.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    os.environ["MV_HOST_TRACE"] = "1"
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
tr.track_sequence_host(params, h[0], h[1], h[2], h[3], out=out)
traced = (time.perf_counter() - t0) * 1e3
rec = torch.tensor([min(times[1:]), traced, up / 1e9, -1 if node is None else node], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(rec) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, rec)
else:
    allr = [rec]
if rank == 0:
    rows = [[float(x) for x in r.cpu()] for r in allr]
    print(json.dumps({"n_gpus": world, "pairs_per_rank": count,
                      "per_rank": [{"rank": i, "best_step_ms": r[0], "traced_step_ms": r[1], "h2d_GB": r[2],
                                    "GBps": r[2] / (r[0] * 1e-3), "numa_node": int(r[3])} for i, r in enumerate(rows)],
                      "sum_GBps": sum(r[2] / (r[0] * 1e-3) for r in rows)}))
if world > 1:
    dist.destroy_process_group()
