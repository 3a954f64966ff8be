#!/usr/bin/env python
"""The reference program's own computation at the reference's own shape, reference vs this library.

tracking_main.c (24x80 cells, top-100 queries, radius-4 window, at most 150 matches, RANSAC over the
identity model, pose from the essential matrix) is what the reference actually executes; it has no
Gauss-Newton PnP.  This tool runs exactly that computation
  * with the reference's unmodified sources (oracle/_ref, main() of tracking_main.c in process) on one
    host core, pair by pair -- the reference is single-threaded -- and
  * with the library's batched kernels (softmax, top-N, matcher, RANSAC-E + pose) on one B200 over a
    batch of synthetic pairs of that shape, CUDA events, 3 warm-ups,
checks that both return the same matches and pose on the timed reference pairs, and prints one JSON line.

    python tools/ref_shape_bench.py [--pairs P] [--ref-pairs R]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16384)
    ap.add_argument("--ref-pairs", type=int, default=64)
    args = ap.parse_args()
    import torch
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import synth, tracking
    from oracle import orc

    rows, cols, seed = 24, 80, 0
    n_frames = args.pairs + 1
    tr = tracking.Tracker(0)
    off = synth.default_offsets(n_frames, seed)
    semi, desc, depth = tr.synth_frames(seed, rows, cols, 0, off)
    scale = torch.full((n_frames,), float(synth.SEMI_SCALE), device=tr.device)
    mp = tracking.match_params(rows, cols, 4, 4, 4, 150)

    def step():
        idx, prob, _ = tr.softmax(semi, scale)
        qp, qi, _, qc, _ = tr.top_n(idx, prob, 100, 1000)
        pts, cnt, _, _, _ = tr.match(mp, desc, idx, prob, qp, qi, qc)
        ninl, _, pose = tr.ransac_identity(pts, cnt, 10, 1.1)
        return pts, cnt, ninl, pose

    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    pts, cnt, ninl, pose = [x.cpu().numpy() for x in out]

    ref_line = None
    if orc.have_ref():
        ref = orc.Reference()
        R = min(args.ref_pairs, args.pairs)
        hs, hd = semi[:R + 1].cpu().numpy(), desc[:R + 1].cpu().numpy()
        same = True
        t0 = time.perf_counter()
        res = [ref.tracking_main(synth.SEMI_SCALE, hs[p], hd[p], synth.SEMI_SCALE, hs[p + 1], hd[p + 1]) for p in range(R)]
        dt = (time.perf_counter() - t0) / R
        for p, r in enumerate(res):
            n = r["n"]
            same &= n == cnt[p] and r["num_inliers"] == ninl[p]
            same &= np.array_equal(r["pts0"], pts[p, :n, :2]) and np.array_equal(r["pts1"], pts[p, :n, 2:])
        ref_line = {"value": 1.0 / dt, "unit": "frame-pairs/s", "cores": 1, "kind": "reference",
                    "sample": "%d pairs, main() of src/tracking_main.c in process (setup + match + RANSAC-E + pose, "
                              "incl. loading the pair into its globals)" % R,
                    "same_matches_and_inliers_as_gpu": bool(same)}
    print(json.dumps({
        "metric": "frame-pairs/sec, the reference program's computation at its native shape (24x80 cells, N=100, "
                  "<=150 matches, RANSAC-E, no Gauss-Newton PnP)",
        "value": args.pairs / (ms * 1e-3), "unit": "frame-pairs/s", "ms_per_step": ms, "pairs": args.pairs,
        "mean_matches_per_pair": float(cnt.mean()), "cpu_baseline": ref_line}))


if __name__ == "__main__":
    main()
