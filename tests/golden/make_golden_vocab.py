"""Writes tests/golden/ref_vocab.npz: the reference's BoW vocabulary (include/data/LCD/vocabulary.h: 10 base
nodes x 256 int8, their scale / bias, 10 x 1000 leaf descriptors of 4 ints) read out of the reference's own
bow_main.c compiled where it lies (oracle/_ref/libmaveric_ref_bow.so).  Run here, where /root/reference exists:

    python tests/golden/make_golden_vocab.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

base, scale, bias, leaves = orc.ReferenceBow().vocabulary()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_vocab.npz"), base_desc=base, scale=scale, bias=bias,
                    leaves=leaves)
print(base.shape, scale, bias, leaves.shape, leaves[0, 0], leaves[9, 999])
