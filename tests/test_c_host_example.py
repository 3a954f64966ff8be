"""examples/track_sequence_main.c: a plain C99 host program on the batched C ABI (the reference's host
language; its own driver is src/tracking_main.c, one pair from a compiled-in header).  Built with gcc
against include/maveric_b200.h and the shared library only -- no CUDA header, no C++ type."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "track_sequence")


def _build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def _write_frames(path, rows, cols, scale, semi, desc, depth):
    with open(path, "wb") as f:
        f.write(np.array([semi.shape[0], rows, cols, 0], np.int32).tobytes())
        for a, t in ((scale, np.float32), (semi, np.int8), (desc, np.int8), (depth, np.float32)):
            f.write(np.ascontiguousarray(a, t).tobytes())


def test_c_host_program_builds_and_refuses_to_run_without_a_gpu(tmp_path, synth):
    """-Wall -Werror C99 against the public header; without an sm_100 device it stops with the
    library's MV_ERR_NO_DEVICE text (there is no CPU path to fall back to)."""
    import torch
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    bad = tmp_path / "bad.bin"
    bad.write_bytes(np.array([1, 24, 80, 0], np.int32).tobytes())
    r = subprocess.run([exe, str(bad), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 2 and "bad header" in r.stderr
    if torch.cuda.is_available():
        return
    frames = [synth.synth_frame(1, 5, 7, f, 0, 0) for f in range(2)]
    _write_frames(tmp_path / "f.bin", 5, 7, np.full(2, synth.SEMI_SCALE, np.float32),
                  np.stack([x[0] for x in frames]), np.stack([x[1] for x in frames]), np.stack([x[2] for x in frames]))
    r = subprocess.run([exe, str(tmp_path / "f.bin"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 2 and "no usable CUDA device" in r.stderr
    assert not (tmp_path / "o.bin").exists()


@pytest.mark.gpu
def test_c_host_program_returns_the_library_bytes(tmp_path, tracker, synth):
    """The C program's result records for a short sequence equal what the Python binding gets from the same
    entry point and from the device-resident one."""
    import torch
    from maveric_slam_b200 import tracking
    exe = _build()
    rows, cols, n, seed, H = 24, 80, 5, 4, 64
    semi, desc, depth = tracker.synth_frames(seed, rows, cols, 0, synth.default_offsets(n, seed))
    scale = torch.full((n,), float(synth.SEMI_SCALE), device=tracker.device)
    p = tracking.track_params(rows, cols, top_n=100, max_valid=1000, max_matches=150, hypotheses=H)
    want = tracking.results_to_numpy(tracker.track_sequence(p, semi, scale, desc, depth))
    assert (want["num_matches"] > 0).all()
    _write_frames(tmp_path / "f.bin", rows, cols, scale.cpu().numpy(), semi.cpu().numpy(), desc.cpu().numpy(),
                  depth.cpu().numpy())
    r = subprocess.run([exe, str(tmp_path / "f.bin"), str(tmp_path / "o.bin"), "100", "1000", "150", str(H)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.bin").read_bytes() == want.tobytes()
    lines = r.stdout.strip().splitlines()
    assert len(lines) == n - 1 and lines[0].startswith("pair 0: num_matches = %d," % want["num_matches"][0])
