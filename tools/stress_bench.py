#!/usr/bin/env python
"""BASELINE.json configs[4] end to end on one B200: 94x310 cells, ~12k keypoints per frame (16k
nominal), 33x33 search window, up to 16384 matches per pair, 4096 Gauss-Newton
hypotheses.  Device-resident frames, CUDA events, 2 warm-ups; prints one JSON line with the per-kernel
breakdown.  Also a scale check of the whole path: two runs must return identical bytes and both
matchers must agree.

    python tools/stress_bench.py [--frames F] [--steps K]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=33)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import synth, tracking

    rows, cols, seed = 94, 310, 0
    tr = tracking.Tracker(0)
    off = synth.default_offsets(args.frames, seed)
    semi, desc, depth = tr.synth_frames(seed, rows, cols, 0, off, keypoint_permille=550)
    scale = torch.full((args.frames,), float(synth.SEMI_SCALE), device=tr.device)

    def params(tc=True):
        return tracking.track_params(rows, cols, top_n=16000, max_valid=32768, max_matches=16384, hypotheses=4096,
                                     radius=16, shift=(4, 4), use_tensor_cores=tc)

    out = tr.track_sequence(params(), semi, scale, desc, depth)
    ref = out.cpu().numpy().tobytes()
    same = tr.track_sequence(params(), semi, scale, desc, depth).cpu().numpy().tobytes() == ref
    dp4a = tr.track_sequence(params(False), semi, scale, desc, depth).cpu().numpy().tobytes() == ref
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = tr.track_sequence(params(), semi, scale, desc, depth, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    tr.ctx.profile(True)
    tr.track_sequence(params(), semi, scale, desc, depth, out=out)
    tr.ctx.sync()
    prof = {t: tr.ctx.profile_read(t)[0] for t in ["detect", "topn", "match", "emit", "ransac", "gather", "pnp", "pnp_select"]}
    tr.ctx.profile(False)
    res = tracking.results_to_numpy(out)
    n_pairs = args.frames - 1
    print(json.dumps({
        "metric": "frame-pairs/sec (window match + PnP), BASELINE configs[4] stress shape, 1 GPU",
        "value": n_pairs / (ms * 1e-3), "unit": "frame-pairs/s", "ms_per_step": ms, "pairs": n_pairs,
        "config": {"grid": [rows, cols], "radius": 16, "top_n": 16000, "max_matches": 16384, "hypotheses": 4096,
                   "mean_matches_per_pair": float(res["num_matches"].mean()),
                   "mean_pnp_inliers": float(res["pnp_inliers"].mean())},
        "kernel_ms": prof, "deterministic": bool(same), "dp4a_matcher_same_bytes": bool(dp4a),
        "status_ok": bool((res["status"] == 0).all())}))


if __name__ == "__main__":
    main()
