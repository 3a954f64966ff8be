"""Frame pairs at the reference's native shape (24x80 cells) built to sit on the parity hazards of
SURVEY App. B.  Used by the T2-vs-T1 test (CPU) and by the GPU matchers' parity test, so the three
implementations see the same bytes."""
import numpy as np


def adversarial_pair(oracle, synth, seed):
    """-> (semi_scale, semi0, desc0, semi1, desc1).  seed % 4 picks the hazard:
    0  extreme entries (+-127/-128): both int32 products at tracking_main.c:154 wrap; scores come out
       negative, > 1, inf or NaN
    1  five distinct descriptors in all: exact score ties in every window (first candidate in x-outer /
       y-inner order must win)
    2  runs of all-zero descriptors at the head of the windows (the sticky 256-d branch of squared_dist)
    3  sign flips (the score loses the sign of the dot product) and saturated rows
    and in every case low-entropy logits on the query frame (probability ties at the top-N threshold)."""
    rng = np.random.default_rng(1000 + seed)
    off = synth.default_offsets(2, seed)
    s0, d0, _ = synth.synth_frame(seed, 24, 80, 0, int(off[0, 0]), int(off[0, 1]), 400, 6)
    s1, d1, _ = synth.synth_frame(seed, 24, 80, 1, int(off[1, 0]), int(off[1, 1]), 400, 6)
    d0, d1, s1 = d0.copy(), d1.copy(), s1.copy()
    kind = seed % 4
    if kind == 0:      # extreme entries everywhere
        d0[:] = rng.choice(np.array([-128, -127, 127], np.int8), size=d0.shape)
        d1[:] = rng.choice(np.array([-128, 127], np.int8), size=d1.shape)
    elif kind == 1:    # a handful of distinct descriptors: ties in every window; frame 1 holds the same ones
        pool = rng.integers(-128, 128, size=(5, 256)).astype(np.int8)
        d0[:] = pool[rng.integers(0, 5, size=d0.shape[0])]
        d1[:] = pool[rng.integers(0, 5, size=d1.shape[0])]
    elif kind == 2:    # zero descriptors in runs (columns of cells), the rest a copy of the query side
        d0[:] = np.roll(d1.reshape(80, 24, 256), (4, 4), axis=(0, 1)).reshape(1920, 256)
        z = d0.reshape(80, 24, 256)
        z[:, ::2] = 0
        z[::5] = 0
    else:              # sign flips (the score loses the sign of the dot product) and saturated rows
        d0[:] = np.roll(d1.reshape(80, 24, 256), (4, 4), axis=(0, 1)).reshape(1920, 256)
        d0[::2] = np.where(d0[::2] == -128, 127, -d0[::2])
        d0[1::7] = 127
    # low-entropy logits on the query frame: many cells share one probability
    scale = float(synth.SEMI_SCALE)
    ki, kp, _ = oracle.softmax(scale, s1)
    live = (ki != 64) & (kp > 0.2)           # keypoint cells only: the valid count stays below top_N.c:91's exit(1)
    s1[live, :64] = np.where(s1[live, :64] > 0, 40, s1[live, :64])
    return scale, s0, d0, s1, d1


def adversarial_logits(seed):
    """-> int8 [1920, 65] detector input for the order-dependent parts of top_N.c (seed in 0..5 picks the
    value palette): equal maxima inside a cell, logits of 127, dustbin-only and all-negative cells, zero
    logits, and a few distinct probabilities shared by hundreds of cells."""
    rng = np.random.default_rng(77 + seed)
    semi = np.full((1920, 65), -50, np.int8)
    semi[:, 64] = 60                                        # dustbin wins: not a keypoint
    active = rng.random(1920) < 0.3                         # ~576 valid cells, below top_N.c:91's exit(1)
    palette = np.array([[-5, 0, 3, 3, 120, 127], [0, 0, 1, 1, 2, 2], [-128, -1, 0, 127, 127, 127],
                        [5, 5, 5, 5, 5, 5], [-3, 7, 7, 90, 90, 126], [0, 0, 0, 0, 0, 0]], np.int8)[seed]
    n_act = int(active.sum())
    semi[active] = palette[rng.integers(0, 6, size=(n_act, 65))]
    semi[active, 64] = rng.choice(np.array([-9, 0, 3], np.int8), size=n_act)
    lone = np.flatnonzero(~active)[:40]
    semi[lone[:20], :] = -7                                 # all negative
    semi[lone[20:], :64] = -1                               # dustbin the only non-negative entry
    return semi
