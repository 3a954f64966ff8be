"""The drop-in claim, end to end: the reference's own driver src/tracking_main.c, unmodified and
compiled where it lies under /root/reference (oracle/Makefile, target `dropin`), linked against
libmaveric_b200.so instead of the reference's pnp_solver.c and top_N.c.  Its compute_softmax,
compute_top_N, ransac_essential_matrix and recover_pose_from_essential_matrix calls then run on
the GPU; what it prints must be what the all-reference build of the same file prints (minus
call_svd's debug lines, which the library does not reproduce -- DESIGN.md §1)."""
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "tracking_main_reference")
DROPIN_BIN = os.path.join(ROOT, "oracle", "_ref", "tracking_main_dropin")
KEEP = re.compile(r"^(Number|R1|R2|t:|    -?[0-9])")


def _lines(binary):
    out = subprocess.run([binary], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return [l for l in out.stdout.splitlines() if KEEP.match(l)]


def _golden():
    return open(os.path.join(GOLDEN, "ref_dropin_stdout.txt")).read().splitlines()


def test_golden_is_what_the_reference_build_prints():
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/tracking_main_reference not built (needs /root/reference)")
    got = _lines(REF_BIN)
    assert got == _golden()
    assert got[0] == "Number of matches: 93" and len(got) == 8


def test_dropin_binary_links_against_the_library():
    if not os.path.exists(DROPIN_BIN):
        pytest.skip("oracle/_ref/tracking_main_dropin not built (needs /root/reference)")
    ldd = subprocess.run(["ldd", DROPIN_BIN], capture_output=True, text=True).stdout
    assert "libmaveric_b200.so" in ldd and "not found" not in ldd
    nm = subprocess.run(["nm", "-D", "--undefined-only", DROPIN_BIN], capture_output=True, text=True).stdout
    for sym in ("compute_softmax", "compute_top_N", "ransac_essential_matrix", "recover_pose_from_essential_matrix"):
        assert re.search(r"\bU %s\b" % sym, nm), sym      # resolved by the library, not compiled in


@pytest.mark.gpu
def test_reference_driver_runs_on_the_library():
    if not os.path.exists(DROPIN_BIN):
        pytest.skip("oracle/_ref/tracking_main_dropin not built")
    assert _lines(DROPIN_BIN) == _golden()


LBA_BIN = os.path.join(ROOT, "oracle", "_ref", "lba_dropin")


@pytest.mark.gpu
def test_reference_lba_driver_runs_its_matmul_shim_on_the_library():
    """src/local_bundle_adjustment.c, unmodified, compiled against this repository's
    include/gemmini_functions_cpu.h (declarations only): its 8 500 matmul2 calls execute in
    libmaveric_b200.so on the GPU and the matrix it hands to cholesky() is the one the all-reference
    run produces (tests/golden/ref_lba.npz, case 0: NaN in the pose block, finite gradient row)."""
    import numpy as np
    if not os.path.exists(LBA_BIN):
        pytest.skip("oracle/_ref/lba_dropin not built")
    nm = subprocess.run(["nm", "-D", "--undefined-only", LBA_BIN], capture_output=True, text=True).stdout
    assert re.search(r"\bU matmul2\b", nm)
    out = subprocess.run([LBA_BIN], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.split()
    assert lines[:2] == ["dim", "49"]
    got = np.array([int(x, 16) for x in lines[2:]], np.uint32).view(np.float32).reshape(49, 49)
    want = np.load(os.path.join(GOLDEN, "ref_lba.npz"))["C"][0]
    nan = np.isnan(want)
    assert (np.isnan(got) == nan).all()
    assert (got[~nan].view(np.int32) == want[~nan].view(np.int32)).all()
