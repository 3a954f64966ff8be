"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's pose chaining.

Follows /root/reference/python/compute_trajectory.py line for line (numpy float64, sequential):
  :51      current_pose = np.eye(4)
  :76      current_pose[:3, :3] = transform[:3, :3] @ current_pose[:3, :3]
  :77      current_pose[:3,  3] = transform[:3,  3] + current_pose[:3,  3]
Pinned by tests/golden/ref_traj.npz, produced by running the reference's own main() here
(tests/golden/make_golden_traj.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this file.
"""
import numpy as np


def chain_transforms(transforms: np.ndarray) -> np.ndarray:
    """float64 [n, 3, 4] relative transforms -> float64 [n+1, 3, 4] poses (pose 0 = identity)."""
    transforms = np.asarray(transforms, np.float64)
    current_pose = np.eye(4)
    out = [current_pose[:3, :].copy()]
    for k in range(transforms.shape[0]):
        transform = np.eye(4)
        transform[:3, :] = transforms[k]
        current_pose[:3, :3] = transform[:3, :3] @ current_pose[:3, :3]
        current_pose[:3, 3] = transform[:3, 3] + current_pose[:3, 3]
        out.append(current_pose[:3, :].copy())
    return np.stack(out)


def quat_t_to_transform(q, t) -> np.ndarray:
    """(w,x,y,z), t -> 3x4 [R|t] in float64 (rotation of src/types.c:62-68 for a unit quaternion)."""
    w, x, y, z = [np.float64(v) for v in q]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], np.float64)
    return np.concatenate([R, np.asarray(t, np.float64).reshape(3, 1)], axis=1)
