#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built objects (cuobjdump -sass), the evidence that the hot kernels are
Blackwell-native: tcgen05 MMA (UTCIMMA), TMEM loads (LDTM), TMA tensor / bulk copies (UTMALDG, UBLKCP), mbarrier
(SYNCS), packed FP32 (FFMA2 / FMUL2), dp4a (IDP.4A), POPC ...      python tools/sass_summary.py > profiles/r02/sass_opcounts.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "maveric-slam_b200", "build")
WANT = {"match_tc.o": ["match_tc_kernel", "lead_kernel", "compact_candidates_kernel"],
        "pnp_gn.o": ["pnp_gn_twophase_kernelILi2", "pnp_gn_sorted_kernelILi2", "pnp_gn_kernelILi32"],
        "api.o": ["gather_rows_kernel"], "bow.o": ["bow_assign_kernel"], "detector.o": ["softmax_cells_kernel", "top_n_kernel"],
        "match.o": ["match_queries_kernel", "emit_matches_kernel"]}
KEY = ["UTCIMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU",
       "IDP", "POPC", "LOP3", "IMAD", "ISETP", "FSETP", "SEL", "LDS", "STS", "LDG", "STG", "ATOMS", "SHFL", "BAR", "BRA", "FLO"]

for obj, kernels in WANT.items():
    path = os.path.join(OBJ, obj)
    if not os.path.exists(path):
        continue
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
            counts[cur]["__total__"] += 1
    for k in kernels:
        for fn, c in counts.items():
            if k in fn:
                print("%s :: %s  (%d SASS instructions)" % (obj, k, c["__total__"]))
                print("    " + "  ".join("%s %d" % (n, c[n]) for n in KEY if c[n]))
