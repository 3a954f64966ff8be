"""Pose chaining (python/compute_trajectory.py), SURVEY §8f rank 2.

CPU: the numpy oracle against what the reference's own main() wrote (tests/golden/ref_traj.npz,
'%.6f' files -> 5e-7), and the product's file writers against the reference's files byte for byte.
GPU: the parallel scan against the sequential oracle (float64, regrouped products -> 1e-12)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import traj_oracle


@pytest.fixture(scope="module")
def ref_traj():
    return np.load(os.path.join(GOLDEN, "ref_traj.npz"))


def test_oracle_matches_reference_files(ref_traj):
    poses = traj_oracle.chain_transforms(ref_traj["transforms"])
    assert poses.shape == ref_traj["poses_6dp"].shape
    assert np.abs(poses - ref_traj["poses_6dp"]).max() <= 5.0e-7 + 1e-12      # files hold 6 decimals
    assert (poses[0] == np.eye(4)[:3]).all()


def test_writers_reproduce_reference_files(ref_traj, tmp_path):
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import tracking
    poses = traj_oracle.chain_transforms(ref_traj["transforms"])
    start = int(ref_traj["start"])
    tracking.write_trajectory(str(tmp_path), start, poses)
    for k, txt in enumerate(ref_traj["pose_txt"]):
        assert open(tmp_path / f"frame-{start + k:06d}.pose.txt").read() == str(txt)
    ply = open(tmp_path / f"trajectory_{start:06d}_{start + len(poses) - 1:06d}.ply").read()
    assert ply == str(ref_traj["ply"])


def test_quat_to_transform_is_a_rotation():
    rng = np.random.default_rng(1)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    T = traj_oracle.quat_t_to_transform(q, [1, 2, 3])
    assert np.abs(T[:, :3] @ T[:, :3].T - np.eye(3)).max() < 1e-14 and (T[:, 3] == [1, 2, 3]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1023, 1024, 1025, 4540])
def test_chain_scan_vs_oracle(tracker, n):
    import torch
    rng = np.random.default_rng(n)
    T = np.zeros((n, 3, 4))
    for k in range(n):
        q = rng.normal(size=4) * [1, 0.01, 0.02, 0.01] + [1, 0, 0, 0]
        T[k] = traj_oracle.quat_t_to_transform(q / np.linalg.norm(q), rng.normal(0, 1, 3) + [0, 0, -1])
    got = tracker.chain_transforms(torch.from_numpy(T).to(tracker.device)).cpu().numpy()
    ref = traj_oracle.chain_transforms(T)
    assert got.shape == ref.shape and (got[0] == np.eye(4)[:3]).all()
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(got - ref).max() <= 1e-12 * scale * max(1, n) ** 0.5


@pytest.mark.gpu
def test_chain_golden_and_results_to_transforms(tracker, ref_traj):
    import torch
    from maveric_slam_b200 import tracking
    T = ref_traj["transforms"]
    got = tracker.chain_transforms(torch.from_numpy(T).to(tracker.device)).cpu().numpy()
    assert np.abs(got - ref_traj["poses_6dp"]).max() <= 5.0e-7 + 1e-12
    # result records -> transforms
    rng = np.random.default_rng(3)
    rec = np.zeros(50, tracking.PAIR_RESULT_DTYPE)
    q = rng.normal(size=(50, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    rec["q"] = q.astype(np.float32); rec["t"] = rng.normal(size=(50, 3)).astype(np.float32)
    dev = torch.from_numpy(rec.view(np.uint8).reshape(50, 64)).to(tracker.device)
    Tg = tracker.results_to_transforms(dev).cpu().numpy()
    for k in range(50):
        assert np.abs(Tg[k] - traj_oracle.quat_t_to_transform(rec["q"][k], rec["t"][k])).max() <= 1e-15
