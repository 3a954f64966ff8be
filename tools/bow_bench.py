#!/usr/bin/env python
"""BoW word assignment on the bench sequence: every query of every frame gets its word.
CUDA events from the context's profile mode; roofline against measured HBM (256 B read per query + 8 B out)
and the word-operation count (10 x 64 dp4a + 1000 x 8 XOR/POPC per query).   python tools/bow_bench.py [frames]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import maveric_slam_b200  # noqa: E402,F401
from maveric_slam_b200 import synth, tracking  # noqa: E402

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 4541
tr = tracking.Tracker(0)
v = np.load(os.path.join(ROOT, "tests", "golden", "ref_vocab.npz"))
tr.bow_set_vocabulary(v["base_desc"], v["scale"], v["bias"], v["leaves"])
semi, desc, depth = tr.synth_frames(0, 47, 155, 0, synth.default_offsets(NF, 0))
scale = torch.full((NF,), float(synth.SEMI_SCALE), device=tr.device)
idx, prob, _ = tr.softmax(semi, scale)
qp, qi, _, qc, _ = tr.top_n(idx, prob, 1000, 8192)
dscale = torch.full((NF,), 4.3353, device=tr.device)
for _ in range(2):
    word, base = tr.bow_assign(desc, dscale, qp, qc)
tr.ctx.sync()
tr.ctx.profile(True)
for _ in range(5):
    word, base = tr.bow_assign(desc, dscale, qp, qc)
tr.ctx.sync()
ms = tr.ctx.profile_read("bow")[0]
tr.ctx.profile(False)
nq = int(qc.sum().item())
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
bytes_ = nq * (256 + 8)
w = word.cpu().numpy()
print(json.dumps({"frames": NF, "queries": nq, "ms_per_launch": ms, "queries_per_s": nq / (ms * 1e-3),
                  "hbm_GBps": bytes_ / (ms * 1e-3) / 1e9, "hbm_frac_of_measured": bytes_ / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                  "word_ops_per_query": 10 * 64 + 1000 * 8 * 2, "Tword_ops_per_s": nq * (640 + 16000) / (ms * 1e-3) / 1e12,
                  "distinct_words_frame0": int(len(set(w[0][w[0] >= 0].tolist())))}))
