#!/usr/bin/env python
"""Measured dense int8 tensor peak of this GPU, the denominator of the matcher's `bound: tensor` roofline
entry (SURVEY §6 / VERDICT r01 item 4: the driver's MEASURED_PEAKS.json has bf16 only).

A TOOL, not the product: a library s8 x s8 -> s32 GEMM (torch._int_mm = cuBLASLt) at 8192^3, timed like the
driver times bf16 (best of 10 = burst; back to back for 4 s = sustained), bf16 measured beside it with the
same code so the two are comparable.  Writes one JSON object (default: profiles/int8_peak.json, which
bench.py reads; without it bench.py falls back to 2 x the bf16 figure and says so).

    python tools/int8_peak.py [--out profiles/int8_peak.json] [--n 8192]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def time_op(fn, ops, seconds=4.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(10, int(seconds * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / reps
    return ops / (best * 1e-3) / 1e12, ops / (sustained * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "int8_peak.json"))
    ap.add_argument("--n", type=int, default=8192)
    a = ap.parse_args()
    if not torch.cuda.is_available():
        sys.exit("needs a GPU")
    n = a.n
    dev = torch.device("cuda", 0)
    A = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=dev)
    B = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=dev)
    ops = 2.0 * n * n * n
    i8_burst, i8_sus = time_op(lambda: torch._int_mm(A, B), ops)
    Ab, Bb = A.to(torch.bfloat16), B.to(torch.bfloat16)
    bf_burst, bf_sus = time_op(lambda: torch.matmul(Ab, Bb), ops)
    out = {"int8_tops": i8_burst, "int8_tops_sustained": i8_sus, "bf16_tflops": bf_burst,
           "bf16_tflops_sustained": bf_sus, "n": n, "gpu_name": torch.cuda.get_device_name(0),
           "how": "torch._int_mm (cuBLASLt s8 x s8 -> s32) and torch.matmul bf16 at %d^3, 2*N^3 ops: best of 10 (burst) "
                  "and back to back for 4 s (sustained), CUDA events" % n,
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    print(json.dumps(out))
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
