#!/usr/bin/env python
"""Measured FP32 FMA throughput of this GPU (the denominator context for the PnP kernel's
roofline): chains of dependent FFMA / packed FFMA2 per thread, 4..32 warps per SM.
Builds tools/ffma_peak.cu with nvcc on first use."""
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
exe = os.path.join(here, "ffma_peak")
src = os.path.join(here, "ffma_peak.cu")
if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, src], check=True)
sys.exit(subprocess.run([exe]).returncode)
