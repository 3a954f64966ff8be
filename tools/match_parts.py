#!/usr/bin/env python
"""Device time of the tensor-core matcher's three kernels (candidate compaction, leading candidates,
tile GEMM + epilogue) and of the emit kernel on the bench workload, CUDA events from the context's
profile mode.    python tools/match_parts.py [frames]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import maveric_slam_b200  # noqa: E402,F401
from maveric_slam_b200 import synth, tracking  # noqa: E402

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 4541
tr = tracking.Tracker(0)
semi, desc, depth = tr.synth_frames(0, 47, 155, 0, synth.default_offsets(NF, 0))
scale = torch.full((NF,), float(synth.SEMI_SCALE), device=tr.device)
idx, prob, _ = tr.softmax(semi, scale)
qp, qi, _, qc, _ = tr.top_n(idx, prob, 1000, 8192)
p = tracking.match_params(47, 155, 4, 4, 4, 1024, use_tensor_cores=True)
for _ in range(3):
    tr.match(p, desc, idx, prob, qp, qi, qc)
tr.ctx.sync()
tr.ctx.profile(True)
for _ in range(5):
    tr.match(p, desc, idx, prob, qp, qi, qc)
tr.ctx.sync()
out = {t: tr.ctx.profile_read(t)[0] for t in ("match", "match_compact", "match_lead", "match_gemm", "emit")}
tr.ctx.profile(False)
out["pairs"] = NF - 1
print(json.dumps(out))
