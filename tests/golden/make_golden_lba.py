#!/usr/bin/env python
"""Writes tests/golden/ref_lba.npz by RUNNING THE REFERENCE (oracle/_ref, needs /root/reference):
local_bundle_adjustment.c's main() on its own input and on three substituted inputs; the matrix it
hands to cholesky() is the known answer.  Also invert_3x3 on random matrices."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

ref = orc.Reference()
rng = np.random.default_rng(2024)
flats = [None] + [rng.normal(size=640).astype(np.float32) for _ in range(3)]
C = np.stack([ref.lba_run(f) for f in flats])
flat_arr = np.stack([np.zeros(640, np.float32) if f is None else f for f in flats])
m = rng.normal(size=(64, 3, 3)).astype(np.float32)
inv = np.stack([ref.invert_3x3(x) for x in m])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_lba.npz"), flat=flat_arr, default_first=np.array(1),
                    C=C, inv_in=m, inv_out=inv)
print("wrote ref_lba.npz: C", C.shape, "nan per case", [int(np.isnan(c).sum()) for c in C])
