#!/usr/bin/env python
"""BASELINE.json configs[2] by itself: batched PnP RANSAC, 1024 hypotheses x 6-DoF Gauss-Newton (4 minimal-sample
+ 10 gated refinement passes + scoring) over ~1k correspondences per frame pair, on one B200.

    python tools/pnp_bench.py [--pairs 1024] [--n 1000] [--hyp 1024] [--reps 5]

Synthetic PnP problems (synth.synth_pnp_problem: pixels uniform in the image, depth U[4,40] m, KITTI intrinsics,
0.5 px noise, 20 % outliers), `pairs` of them per launch.  Prints one JSON line with the kernel time (CUDA events on
the launch stream, profile mode of the context), the executed FP32 work by the oracle's operation count (32 flop to
project and gate a correspondence in every pass, 95 more for each accepted one, counted by the kernel) against the
nominal FP32 peak (SMs x 128 x 2 x clocks.max.sm), and the selected poses' agreement with the generating poses.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import maveric_slam_b200  # noqa: E402,F401
from maveric_slam_b200 import lib, synth, tracking  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1024)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--hyp", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    tr = tracking.Tracker(0)
    stride = (a.n + 31) // 32 * 32
    base = 64                      # distinct problems, tiled over the launch (the sampler is keyed on the pair index)
    probs = np.zeros((base, 5, stride), np.float32)
    truth = []
    for p in range(base):
        c, pose, _ = synth.synth_pnp_problem(7000 + p, a.n, stride=stride)
        probs[p] = c
        truth.append(pose)
    corr = torch.from_numpy(np.tile(probs, ((a.pairs + base - 1) // base, 1, 1))[:a.pairs].copy()).to(tr.device)
    cnt = torch.full((a.pairs,), a.n, dtype=torch.int32, device=tr.device)
    prm = lib.PnpParams()
    lib.load().mv_pnp_params_default(C.byref(prm))
    prm.hypotheses = a.hyp
    for _ in range(2):
        pose, stats, _ = tr.pnp_gn(prm, corr, cnt)
    tr.ctx.sync()
    tr.ctx.profile(True)
    tr.ctx.pnp_work()
    for _ in range(a.reps):
        pose, stats, _ = tr.pnp_gn(prm, corr, cnt)
    tr.ctx.sync()
    ms = tr.ctx.profile_read("pnp")[0]
    accepted = tr.ctx.pnp_work() // a.reps
    tr.ctx.profile(False)
    sample_iters, refine_iters = prm.sample_iters, prm.refine_iters
    flops = (a.pairs * a.hyp * sample_iters * 8 * 127 + a.hyp * a.pairs * a.n * (refine_iters + 1) * 32 + accepted * 95)
    props = torch.cuda.get_device_properties(0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    fp32_peak = props.multi_processor_count * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    pose = pose.cpu().numpy()
    err = [float(np.linalg.norm(pose[p, 4:] - truth[p % base][4:])) for p in range(min(a.pairs, base))]
    print(json.dumps({
        "workload": "configs[2]: %d pairs x %d hypotheses x (%d+%d) GN passes over %d correspondences" % (a.pairs, a.hyp, sample_iters, refine_iters, a.n),
        "kernel": "pnp_gn_twophase_kernel" if a.n <= 480 else "pnp_gn_sorted_kernel (streaming; n > 480)",
        "ms_per_launch": ms, "pairs_per_s": a.pairs / (ms * 1e-3), "accepted_fraction": accepted / (a.hyp * a.pairs * a.n * refine_iters),
        "executed_flops": flops, "achieved_TFLOPs": flops / (ms * 1e-3) / 1e12, "fp32_peak_TFLOPs": fp32_peak,
        "frac": flops / (ms * 1e-3) / 1e12 / fp32_peak, "max_translation_error_m": max(err)}))


if __name__ == "__main__":
    main()
