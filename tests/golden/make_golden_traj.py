"""Generates tests/golden/ref_traj.npz by running the REFERENCE's pose chaining here.

    python tests/golden/make_golden_traj.py        (needs /root/reference; numpy only)

Seeded relative transforms (KITTI-like forward motion, small rotations) are written as the
transform_XXXXXX_YYYYYY.npy files python/compute_trajectory.py reads; its unmodified main() then
writes frame-XXXXXX.pose.txt ('%.6f') and the trajectory PLY, which are stored verbatim.
"""
import contextlib
import importlib.util
import io
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference/python/compute_trajectory.py"


def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def main():
    spec = importlib.util.spec_from_file_location("ref_compute_trajectory", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20261018)
    start, n = 785, 37          # scripts/run_pairwise_pnp.sh starts at frame 785
    T = np.zeros((n, 3, 4))
    for k in range(n):
        T[k, :, :3] = rot(*(rng.normal(0, 0.01, 3) + [0, 0.02 * np.sin(k / 5.0), 0]))
        T[k, :, 3] = [rng.normal(0, 0.03), rng.normal(0, 0.02), -rng.uniform(0.5, 1.2)]
    with tempfile.TemporaryDirectory() as pose_dir, tempfile.TemporaryDirectory() as out_dir:
        for k in range(n):
            np.save(os.path.join(pose_dir, f"transform_{start + k:06d}_{start + k + 1:06d}.npy"), T[k])
        with contextlib.redirect_stdout(io.StringIO()):
            ref.main(start, start + n, pose_dir, out_dir)
        poses = np.stack([np.loadtxt(os.path.join(out_dir, f"frame-{start + k:06d}.pose.txt")) for k in range(n + 1)])
        pose_txt = [open(os.path.join(out_dir, f"frame-{start + k:06d}.pose.txt")).read() for k in range(n + 1)]
        ply = open(os.path.join(out_dir, f"trajectory_{start:06d}_{start + n:06d}.ply")).read()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_traj.npz")
    np.savez_compressed(out, transforms=T, start=start, poses_6dp=poses, pose_txt=np.array(pose_txt), ply=np.array(ply))
    print("wrote", out, poses.shape)


if __name__ == "__main__":
    sys.exit(main())
