"""Synthetic KITTI-shaped frames (numpy, host side).

Counter-based generator: every byte is a pure function of (seed, world cell / frame,
channel group), so this numpy version, the CUDA generator in csrc/synth.cu
(``mv_synth_frames``) and the C one in the test oracle produce identical bytes.
The layout is the reference's (``src/tracking_main.c:59-66``; header writer
``python/superpoint_inference.py:630-664``): per frame ``semi`` int8 ``[cells][65]`` and
``desc`` int8 ``[cells][256]`` with ``cell = col*rows + row``.

A frame shows the window ``[off_x, off_x+cols) x [off_y, off_y+rows)`` of an unbounded
world grid of cells.  Keypoints and base descriptors belong to world cells, so frame
``f+1`` shifted by (4, 4) cells sees frame ``f``'s keypoints 4 cells up-left: the true
match of a query sits at the centre of the reference's search window
(``tracking_main.c:104-106,127-130``).
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_TAIL = np.array([-128, -111, -97, -85, -70, 70, 85, 97, 111, 127], dtype=np.int16)

KITTI_FX, KITTI_FY, KITTI_CX, KITTI_CY = 718.856, 718.856, 607.1928, 185.2157
SEMI_SCALE = np.float32(0.3562202453613281)   # include/data/quantized/quantized_image0.h:14
DESC_SCALE = np.float32(4.335296630859375)    # quantized_image0.h:1938


def sm64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def ctr(mixed_seed, tag, a, b, c) -> np.ndarray:
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    c = np.asarray(c, dtype=np.uint64)
    packed = ((np.uint64(tag) << np.uint64(56))
              | ((a & np.uint64(0xFFFFFF)) << np.uint64(32))
              | ((b & np.uint64(0xFFFFFF)) << np.uint64(8))
              | (c & np.uint64(0xFF)))
    with np.errstate(over="ignore"):
        return sm64((np.uint64(mixed_seed) + packed) & _M64)


def _bytes_of(h: np.ndarray) -> np.ndarray:
    """[..., G] uint64 -> [..., G*8] uint8, little-endian byte order."""
    sh = (np.arange(8, dtype=np.uint64) * np.uint64(8))
    return ((h[..., None] >> sh) & np.uint64(0xFF)).astype(np.int32).reshape(*h.shape[:-1], -1)


def default_offsets(n_frames: int, seed: int = 0, step: int = 4, jitter: int = 2) -> np.ndarray:
    """World-cell offset of every frame: a (step,step) cells/frame drift with a small
    per-frame jitter, so the true displacement stays inside a radius-4 window."""
    ms = sm64(np.array([seed], dtype=np.uint64))[0]
    f = np.arange(n_frames, dtype=np.uint64)
    h = ctr(ms, 6, f, 0, 0)
    span = 2 * jitter + 1
    jx = (h % np.uint64(span)).astype(np.int64) - jitter
    jy = ((h >> np.uint64(8)) % np.uint64(span)).astype(np.int64) - jitter
    dx = step + jx
    dy = step + jy
    dx[0] = 0
    dy[0] = 0
    # query (x1,y1) of frame f+1 is searched around (x1+step, y1+step) in frame f
    # (tracking_main.c:127-130), so frame f+1 looks at world cells `step` further on
    off = np.stack([np.cumsum(dx), np.cumsum(dy)], axis=1)
    return off.astype(np.int32)


def synth_frame(seed: int, rows: int, cols: int, frame: int, off_x: int, off_y: int,
                keypoint_permille: int = 140, noise_amp: int = 6):
    """Returns (semi int8 [cells,65], desc int8 [cells,256], depth f32 [cells])."""
    ms = sm64(np.array([seed], dtype=np.uint64))[0]
    x, y = np.meshgrid(np.arange(cols, dtype=np.int64), np.arange(rows, dtype=np.int64), indexing="ij")
    x = x.reshape(-1)
    y = y.reshape(-1)
    cell = (x * rows + y).astype(np.uint64)
    wx = (x + off_x + (1 << 23)).astype(np.uint64)
    wy = (y + off_y + (1 << 23)).astype(np.uint64)

    k = ctr(ms, 1, wx, wy, 0)
    is_kp = (k % np.uint64(1000)).astype(np.int64) < keypoint_permille
    kp_ch = ((k >> np.uint64(16)) & np.uint64(63)).astype(np.int64)
    kp_val = 14 + ((k >> np.uint64(24)) % np.uint64(29)).astype(np.int64)
    kp_dust = -20 + ((k >> np.uint64(32)) % np.uint64(45)).astype(np.int64)
    bg_dust = 10 + ((k >> np.uint64(32)) % np.uint64(30)).astype(np.int64)
    depth = (np.float32(4.0) + ((k >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float32)
             * np.float32(36.0 / 65536.0)).astype(np.float32)

    g9 = np.arange(9, dtype=np.uint64)
    hs = ctr(ms, 2, np.uint64(frame), cell[:, None], g9[None, :])
    u = _bytes_of(hs)[:, :65]
    semi = np.where(u == 255, (np.arange(65) % 6) * 3, -30 - (u % 70))
    semi[:, 64] = np.where(is_kp, kp_dust, bg_dust)
    rows_kp = np.nonzero(is_kp)[0]
    semi[rows_kp, kp_ch[rows_kp]] = kp_val[rows_kp]
    semi = semi.astype(np.int8)

    g32 = np.arange(32, dtype=np.uint64)
    hb = ctr(ms, 3, wx[:, None], wy[:, None], g32[None, :])
    hn = ctr(ms, 4, np.uint64(frame), cell[:, None], g32[None, :])
    ub = _bytes_of(hb)
    base = np.where(ub < 246, (ub % 67) - 33, _TAIL[np.clip(ub - 246, 0, 9)])
    if noise_amp > 0:
        noise = (_bytes_of(hn) % (2 * noise_amp + 1)) - noise_amp
    else:
        noise = 0
    desc = np.clip(base + noise, -128, 127).astype(np.int8)
    return semi, desc, depth


def synth_sequence(seed: int, rows: int, cols: int, n_frames: int, offsets: np.ndarray | None = None,
                   first_frame: int = 0, keypoint_permille: int = 140, noise_amp: int = 6):
    """Stacked frames: semi [n,cells,65], desc [n,cells,256], depth [n,cells], offsets [n,2]."""
    if offsets is None:
        offsets = default_offsets(first_frame + n_frames, seed)[first_frame:]
    cells = rows * cols
    semi = np.empty((n_frames, cells, 65), np.int8)
    desc = np.empty((n_frames, cells, 256), np.int8)
    depth = np.empty((n_frames, cells), np.float32)
    for i in range(n_frames):
        s, d, z = synth_frame(seed, rows, cols, first_frame + i, int(offsets[i, 0]), int(offsets[i, 1]),
                              keypoint_permille, noise_amp)
        semi[i], desc[i], depth[i] = s, d, z
    return semi, desc, depth, np.ascontiguousarray(offsets[:n_frames], dtype=np.int32)


def synth_pnp_problem(seed: int, n: int, outlier_frac: float = 0.2, noise_px: float = 0.5,
                      width: int = 1241, height: int = 376, stride: int | None = None):
    """A consistent PnP problem: pixels uniform in the image, depth U[4,40] m,
    a KITTI-like relative motion, Gaussian pixel noise, a fraction of gross outliers.
    Returns corr SoA float32 [5, stride] (X,Y,Z,u,v), the true pose (qw,qx,qy,qz,tx,ty,tz)."""
    rng = np.random.default_rng(seed)
    stride = stride or n
    u0 = rng.uniform(0, width, n)
    v0 = rng.uniform(0, height, n)
    d = rng.uniform(4.0, 40.0, n)
    X = np.stack([(u0 - KITTI_CX) / KITTI_FX * d, (v0 - KITTI_CY) / KITTI_FY * d, d], axis=1)
    ang = rng.normal(0, 0.01, 3)
    ang[1] += rng.normal(0, 0.02)
    t = np.array([rng.normal(0, 0.03), rng.normal(0, 0.02), -rng.uniform(0.5, 1.2)])
    th = np.linalg.norm(ang)
    ax = ang / th
    q = np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * ax])
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    Xc = X @ R.T + t
    uv = np.stack([KITTI_FX * Xc[:, 0] / Xc[:, 2] + KITTI_CX,
                   KITTI_FY * Xc[:, 1] / Xc[:, 2] + KITTI_CY], axis=1)
    uv += rng.normal(0, noise_px, uv.shape)
    n_out = int(round(outlier_frac * n))
    out_idx = rng.choice(n, n_out, replace=False)
    uv[out_idx] = np.stack([rng.uniform(0, width, n_out), rng.uniform(0, height, n_out)], axis=1)
    corr = np.zeros((5, stride), np.float32)
    corr[0, :n], corr[1, :n], corr[2, :n] = X[:, 0], X[:, 1], X[:, 2]
    corr[3, :n], corr[4, :n] = uv[:, 0], uv[:, 1]
    pose = np.concatenate([q, t]).astype(np.float32)
    inlier_mask = np.ones(n, bool)
    inlier_mask[out_idx] = False
    return corr, pose, inlier_mask
