#!/usr/bin/env python
"""Throughput of the batched local-BA Schur complement at the reference's shape
(1000 landmarks x 8 poses, chunks of 4): windows/s on one B200 (CUDA events, 3 warm-ups, inputs
larger than L2), achieved TFLOP/s on the reference's operation count against the FP32 rate that
separately rounded multiplies and adds can reach, the HBM GB/s on the algorithmic bytes (80 B per
factor in, (6P+1)^2 floats out), and the CPU restatement (oracle/, the reference's own loop) on this host beside it.

    python tools/lba_bench.py [--windows W] [--reps R]
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
    ap.add_argument("--windows", type=int, default=2368)   # 16 per SM; 1.5 GB of factors
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import tracking
    from oracle import orc

    L, P, ch = 1000, 8, 4
    tr = tracking.Tracker(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    J = torch.randn((args.windows, L, P, 20), generator=g, device=tr.device)
    for _ in range(3):
        C = tr.lba_schur(J, ch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        C = tr.lba_schur(J, ch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    in_bytes = J.numel() * 4
    out_bytes = C.numel() * 4
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    ach = (in_bytes + out_bytes) / (ms * 1e-3) / 1e9

    # the step of the reduced systems (mv_lba_solve_batch); C of random factors is SPD
    for _ in range(3):
        d, ok = tr.lba_solve(C, 1e-3)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.reps):
        d, ok = tr.lba_solve(C, 1e-3)
    e1.record()
    torch.cuda.synchronize()
    ms_solve = e0.elapsed_time(e1) / args.reps

    o = orc.Oracle()
    Jh = J[:4].cpu().numpy()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 5.0:
        ref = o.lba_schur(Jh[n % 4], ch)
        n += 1
    cpu_s = (time.perf_counter() - t0) / n
    same = np.array_equal(ref.view(np.int32), C[(n - 1) % 4].cpu().numpy().view(np.int32))
    # the reference's operation count per window (multiplications by the +-1 scale factors not counted):
    # per factor 100 x (2 mul + 2 add + the 0 * old) + 72 scatter adds; per chunk of 4 landmarks 4 inversions,
    # B A^-1 (12 x 48 x 12 x 2) and the update of C (48 x 48 x 12 x 2)
    flops = (L // ch) * (ch * P * 572 + 180 + 12 * 48 * 12 * 2 + 48 * 48 * 12 * 2)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    fp32_fma = sm * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    tf = flops * args.windows / (ms * 1e-3) / 1e12
    print(json.dumps({
        "metric": "local-BA windows/s (Schur complement, 1000 landmarks x 8 poses, chunks of 4)",
        "value": args.windows / (ms * 1e-3), "unit": "windows/s", "ms_per_launch": ms, "windows_per_launch": args.windows,
        "roofline": {"kernel": "lba_schur_p8c4_kernel", "bound": "fp32", "achieved": tf, "peak": fp32_fma / 2, "unit": "TFLOP/s",
                     "frac": tf / (fp32_fma / 2), "flops_per_window": flops,
                     "peak_note": "every product and sum is rounded separately, as in the reference (no FMA): half of the "
                                  "nominal FMA peak of %.1f TFLOP/s" % fp32_fma,
                     "hbm_achieved_gbs": ach, "hbm_peak_gbs": hbm, "algorithmic_bytes": in_bytes + out_bytes},
        "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "windows/s", "cores": 1, "kind": "port",
                         "sample": "%d windows, oracle/mv_oracle.c orc_lba_schur (-O2, the reference's loop order)" % n},
        "bit_identical_to_cpu": bool(same),
        "solve": {"kernel": "lba_solve_kernel", "ms_per_launch": ms_solve, "windows_solved": int(ok.sum()),
                  "windows_per_s": args.windows / (ms_solve * 1e-3),
                  "bit_identical_to_cpu": bool(all(
                      np.array_equal(o.lba_solve(C[w].cpu().numpy(), 1e-3)[1].view(np.int32),
                                     d[w].cpu().numpy().view(np.int32)) for w in range(4)))}}))


if __name__ == "__main__":
    main()
