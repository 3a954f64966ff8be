"""Evidence kept runnable: the reference's BoW program (src/bow_main.c) has no result to be faithful to, which is
why the word assignment of SURVEY section 8(f) rank 3 is built on a STATED definition (DESIGN.md section 4.6,
csrc/bow.cu) with only its two helper functions and its vocabulary pinned to the reference (tests/test_bow.py).
This compiles that program from the reference's own sources (the two files of its CMake target,
CMakeLists.txt:19-21, no build system) and shows that it dies with SIGSEGV before printing a word, at
every optimisation level -- its int8 arrays go to the float* matmul shim and to int* readers
(bow_main.c:81-86, :105, :115).  Needs /root/reference, so it runs in the build container only."""
import os
import signal
import subprocess

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="needs the reference sources")
@pytest.mark.parametrize("opt", ["-O0", "-O2"])
def test_reference_bow_main_crashes_as_shipped(tmp_path, opt):
    exe = str(tmp_path / "bow_main")
    inc = [f"-I{REF}/include", f"-I{REF}/include/data/LCD", f"-I{REF}/include/data/quantized"]
    r = subprocess.run(["/usr/bin/gcc", "-std=gnu11", opt, "-w", *inc, f"{REF}/src/bow_main.c", f"{REF}/src/top_N.c",
                        "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == -signal.SIGSEGV
    assert "word:" not in r.stdout      # bow_main.c:122 is never reached
