#!/usr/bin/env python
"""Windowed-match sweep (BASELINE.json configs[1] and the matcher of configs[4]) on one B200.

For K = N in {1k, 2k, 4k, 8k, 16k} keypoints per frame (grid per SURVEY §8d: 47x155 up to 4k,
94x155 for 8k, 94x310 for 16k) at radius 4, plus the stress shape (16k, radius 16), runs the
detector once and then times `mv_match_batch` with the dp4a and the tcgen05 matcher on the same
device-resident frames: CUDA events on the launch stream, 3 warm-ups, inputs > L2.  Prints one
JSON line per (shape, matcher) with the algorithmic int8 op count of SURVEY §8d
(A_match = 2*256*sum_q |window_q|, window cells valid or not), the dense-equivalent tile ops the
tensor pipe actually executes, and checks that both matchers return identical bytes.

    python tools/match_sweep.py [--pairs P] [--only stress] [--matcher tc|dp4a|both] [--reps R]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (rows, cols, keypoint_permille, N, radius).  The density is set so that the detector really selects N
# query keypoints per frame (about 0.7-0.8 of the valid cells pass compute_top_N's cut): 2k / 4k on the
# 7 285-cell KITTI grid are 27 % / 55 % of all cells, which needs 33 % / 85 % of the synthetic world's cells
# to be keypoints.  Each output line carries the measured keypoints_per_frame.
SHAPES = {
    "1k": (47, 155, 140, 1000, 4),
    "2k": (47, 155, 330, 2000, 4),
    "4k": (47, 155, 850, 4000, 4),
    "8k": (94, 155, 850, 8000, 4),
    "16k": (94, 310, 850, 16000, 4),
    "stress": (94, 310, 850, 16000, 16),
}


def window_cells(rows, cols, q_patch, q_count, shift, r):
    """sum over queries of the clamped window size (cells, valid or not)."""
    tot = 0
    for f in range(1, q_patch.shape[0]):
        p = q_patch[f, :q_count[f]].astype(np.int64)
        x, y = p // rows, p % rows
        w = np.clip(np.minimum(x + shift + r, cols - 1) - np.maximum(x + shift - r, 0) + 1, 0, None)
        h = np.clip(np.minimum(y + shift + r, rows - 1) - np.maximum(y + shift - r, 0) + 1, 0, None)
        tot += int((w * h).sum())
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per launch (default: sized to ~1.5 GB of descriptors)")
    ap.add_argument("--only", default="")
    ap.add_argument("--matcher", default="both", choices=["both", "tc", "dp4a"])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()

    import torch
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import synth, tracking

    tr = tracking.Tracker(0)
    for name, (rows, cols, permille, N, r) in SHAPES.items():
        if args.only and name not in args.only.split(","):
            continue
        cells = rows * cols
        n_pairs = args.pairs or max(8, min(512, int(1.5e9 / (cells * 256))))
        n_frames = n_pairs + 1
        offs = synth.default_offsets(n_frames, 7)
        semi, desc, depth = tr.synth_frames(7, rows, cols, 0, offs, keypoint_permille=permille)
        scale = torch.full((n_frames,), float(synth.SEMI_SCALE), device=tr.device)
        idx, prob, nv = tr.softmax(semi, scale)
        qp, qi, qpr, qc, ov = tr.top_n(idx, prob, N, cells + 1)
        M = ((N + 255) // 256) * 256
        a_cells = window_cells(rows, cols, qp.cpu().numpy(), qc.cpu().numpy(), 4, r)
        a_ops = 2 * 256 * a_cells
        out = {}
        for m in (["dp4a", "tc"] if args.matcher == "both" else [args.matcher]):
            p = tracking.match_params(rows, cols, 4, 4, r, M, use_tensor_cores=(m == "tc"))
            for _ in range(3):
                res = tr.match(p, desc, idx, prob, qp, qi, qc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                res = tr.match(p, desc, idx, prob, qp, qi, qc)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            out[m] = (ms, [t.cpu().numpy().tobytes() for t in (res[0], res[1], res[2])], int(res[1].sum().item()))
            line = {"shape": name, "grid": [rows, cols], "keypoints_per_frame": float(qc[1:].float().mean().item()),
                    "radius": r, "pairs": n_pairs, "matcher": "tcgen05" if m == "tc" else "dp4a", "ms_per_launch": ms,
                    "pairs_per_s": n_pairs / (ms * 1e-3), "matches_per_pair": out[m][2] / n_pairs,
                    "algorithmic_int8_ops": a_ops, "algorithmic_TOPS": a_ops / (ms * 1e-3) / 1e12,
                    "descriptor_bytes": int(desc.numel())}
            if m == "tc":   # what the tensor pipe executed (128 x 256 x 64 chunks) and the tile kernel's own time
                tiles, chunks = tr.ctx.match_work()
                tr.ctx.profile(True)
                tr.match(p, desc, idx, prob, qp, qi, qc)
                tr.ctx.sync()
                parts = {t_: tr.ctx.profile_read(t_)[0] for t_ in ("match_compact", "match_lead", "match_gemm")}
                tr.ctx.profile(False)
                tile_ops = chunks * 2 * 128 * 256 * 64
                line.update({"executed_tile_int8_ops": tile_ops, "executed_over_algorithmic": tile_ops / a_ops,
                             "kernel_ms": parts,
                             "tile_kernel_TOPS": tile_ops / (parts["match_gemm"] * 1e-3) / 1e12 if parts["match_gemm"] else None})
                try:
                    peak = json.load(open(os.path.join(ROOT, "profiles", "int8_peak.json")))["int8_tops"]
                    line["tile_kernel_frac_of_measured_int8_peak"] = line["tile_kernel_TOPS"] / peak
                except Exception:
                    pass
            print(json.dumps(line), flush=True)
        if len(out) == 2:
            same = out["dp4a"][1] == out["tc"][1]
            print(json.dumps({"shape": name, "identical_results": bool(same),
                              "speedup_tc_over_dp4a": out["dp4a"][0] / out["tc"][0]}), flush=True)
            if not same:
                raise SystemExit("matchers disagree on shape " + name)
        del semi, desc, depth, idx, prob, qp, qi, qpr, qc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
