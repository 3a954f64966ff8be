#!/usr/bin/env python
"""Writes the definitions behind oracle/ref_shim/quantized_pair0.h for the stand-alone drop-in
demonstration (oracle/Makefile: _ref/tracking_main_dropin): both images are the reference's own
fixture include/data/quantized/quantized_image0.h as stored in tests/golden/ref_image0.npz (the
self pair of SURVEY §8c-iii).  TEST INFRASTRUCTURE ONLY.  usage: gen_dropin_data.py out.c"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = np.load(os.path.join(ROOT, "tests", "golden", "ref_image0.npz"))


def arr(name, a):
    rows = ",\n".join("{" + ",".join(str(int(v)) for v in r) + "}" for r in a)
    return "int8_t %s[%d][%d] = {\n%s};\n" % (name, a.shape[0], a.shape[1], rows)


with open(sys.argv[1], "w") as f:
    f.write("#include <stdint.h>\nint cell_size = 8;\n")
    for k in (0, 1):
        f.write("int image%d_rows = 192, image%d_cols = 640, image%d_channels = 1;\n" % (k, k, k))
        f.write("int image%d_feature_rows = 24, image%d_feature_cols = 80;\n" % (k, k))
        f.write("float image%d_semi_scale = %.9gf;\nfloat image%d_desc_scale = %.9gf;\n"
                % (k, float(d["semi_scale"]), k, float(d["desc_scale"])))
        f.write(arr("image%d_semi" % k, d["semi"]))
        f.write(arr("image%d_desc" % k, d["desc"]))
