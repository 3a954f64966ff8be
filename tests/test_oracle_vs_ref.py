"""Pins the CPU restatement (oracle/mv_oracle.c, "T2") to the reference:
 (a) against the committed golden vectors that were produced by running the reference's own
     sources on its own fixture (tests/golden/make_golden.py), and
 (b) live against oracle/_ref/libmaveric_ref.so (the unmodified reference, "T1") over many
     seeds at the reference's native 24x80 / N=100 / <=150 shape.
Bit-exact everywhere (integers, indices, tie order, fp32 bit patterns)."""
import numpy as np
import pytest

from conftest import bits
from oracle import orc


def test_softmax_matches_reference_fixture(oracle, image0, kat):
    idx, pr, nv = oracle.softmax(image0["semi_scale"], image0["semi"])
    assert nv == int(kat["img0_num_valid"]) == 410
    assert (idx == kat["img0_softmax_idx"]).all()
    assert (bits(pr) == bits(kat["img0_softmax_prob"])).all()


def test_softmax_argmax_vs_float_ground_truth(oracle, image0):
    # include/data/quantized/pair0_gt.h: argmax of the float softmax; the Taylor
    # approximation only changes probabilities, never the argmax of a valid cell
    idx, pr, nv = oracle.softmax(image0["semi_scale"], image0["semi"])
    valid = idx != 64
    assert valid.sum() == 410
    assert (idx[valid] == image0["indices_gt"][valid]).all()


def test_top_n_matches_reference_fixture(oracle, image0, kat):
    pa, ix, pr, ov = oracle.top_n(image0["semi_scale"], image0["semi"], 100)
    assert ov == 0 and len(pa) == 100
    assert (pa == kat["img0_top100_patch"]).all() and (ix == kat["img0_top100_idx"]).all()
    assert (bits(pr) == bits(kat["img0_top100_prob"])).all()
    # SURVEY §8c (ii)
    assert list(pa[:5]) == [154, 195, 214, 227, 246] and list(ix[:5]) == [17, 20, 57, 45, 19]


def _t2_pair(oracle, scale0, s0, d0, scale1, s1, d1):
    idx, pr, _ = oracle.softmax(scale0, s0)
    pa, ix, _, ov = oracle.top_n(scale1, s1, 100)
    cfg = orc.MatchCfg(24, 80, 4, 4, 4, 150, 0.9, 0.2)
    return oracle.match(cfg, d0, d1, idx, pr, pa, ix)


def test_self_pair_match_list_golden(oracle, image0, kat):
    m = _t2_pair(oracle, image0["semi_scale"], image0["semi"], image0["desc"],
                 image0["semi_scale"], image0["semi"], image0["desc"])
    assert m["n"] == 93 == len(kat["self_pts0"])
    assert (m["pts0"] == kat["self_pts0"]).all() and (m["pts1"] == kat["self_pts1"]).all()
    # the reference's own counters for this pair
    assert (m["window_cells"], m["pairs_256"] + m["pairs_64"]) == (7488, 1194)
    assert m["pairs_256"] == 99 and m["pairs_64"] == 1095
    E, inl, ni, wrote = oracle.ransac_identity(m["pts0"], m["pts1"])
    assert ni == int(kat["self_num_inliers"]) and (inl == kat["self_inliers"]).all()


def test_synthetic_pairs_golden(oracle, synth, kat):
    for seed in range(4):
        off = synth.default_offsets(2, seed)
        s0, d0, _ = synth.synth_frame(seed, 24, 80, 0, int(off[0, 0]), int(off[0, 1]))
        s1, d1, _ = synth.synth_frame(seed, 24, 80, 1, int(off[1, 0]), int(off[1, 1]))
        m = _t2_pair(oracle, synth.SEMI_SCALE, s0, d0, synth.SEMI_SCALE, s1, d1)
        assert (m["pts0"] == kat[f"syn{seed}_pts0"]).all() and (m["pts1"] == kat[f"syn{seed}_pts1"]).all()


def test_pose_constant_and_svd_golden(oracle, kat):
    R1, R2, t = oracle.recover_pose(np.eye(3, dtype=np.float32))
    assert (bits(R1) == bits(kat["pose_R1"])).all() and (bits(R2) == bits(kat["pose_R2"])).all()
    assert (bits(t) == bits(kat["pose_t"])).all()
    # the constant the reference prints for every input
    assert np.allclose(t, [0.996568, 0.0, -0.000495], atol=5e-7)
    for A, usv in zip(kat["svd_in"], kat["svd_usv"]):
        U, S, V = oracle.svd3(A)
        assert (bits(U) == bits(usv[0])).all() and (bits(S) == bits(usv[1])).all() and (bits(V) == bits(usv[2])).all()


def test_matmul_shim_golden(oracle, kat):
    A, B = kat["mm_A"], kat["mm_B"]
    C = kat["mm_C0"].copy()
    oracle.lib.orc_matmul(7, 6, 5, A, B, C, 5, 6, 6, 0.5, 1.25, 0, 0)
    assert (bits(C) == bits(kat["mm_C1"])).all()
    At, Bt = np.ascontiguousarray(A.T), np.ascontiguousarray(B.T)
    C2 = np.zeros((7, 6), np.float32)
    D = kat["mm_C0"].copy()
    oracle.lib.orc_matmul2(7, 6, 5, At, Bt, D.ctypes.data, C2, 7, 5, 6, 6, 1.5, -0.75, 2.0, 1, 1)
    assert (bits(C2) == bits(kat["mm2_C"])).all()


# ---------------------------------------------------------------- live T1 vs T2
@pytest.mark.parametrize("seed", range(12))
def test_t2_equals_t1_on_native_shape(oracle, reference, synth, seed):
    off = synth.default_offsets(2, seed)
    permille = [60, 140, 250, 400][seed % 4]
    s0, d0, _ = synth.synth_frame(seed, 24, 80, 0, int(off[0, 0]), int(off[0, 1]), permille, 6)
    s1, d1, _ = synth.synth_frame(seed, 24, 80, 1, int(off[1, 0]), int(off[1, 1]), permille, 6)
    if seed % 3 == 2:  # zero descriptors: exercises the sticky-zero-norm branch of squared_dist
        d0 = d0.copy()
        d0[::3] = 0
    scale = float(synth.SEMI_SCALE) * (1.0 + 0.05 * (seed % 3))
    i1, p1, n1 = reference.softmax(scale, s0)
    i2, p2, n2 = oracle.softmax(scale, s0)
    assert n1 == n2 and (i1 == i2).all() and (bits(p1) == bits(p2)).all()
    a = reference.top_n(scale, s1, 100)
    b = oracle.top_n(scale, s1, 100)
    assert all((x == y).all() for x, y in zip(a, b[:2])) and (bits(a[2]) == bits(b[2])).all()
    res = reference.tracking_main(scale, s0, d0, scale, s1, d1)
    m = _t2_pair(oracle, scale, s0, d0, scale, s1, d1)
    assert res["n"] == m["n"]
    assert (res["pts0"] == m["pts0"]).all() and (res["pts1"] == m["pts1"]).all()
    E, inl, ni, wrote = oracle.ransac_identity(m["pts0"], m["pts1"])
    assert ni == res["num_inliers"] and (inl == res["inliers"]).all()


@pytest.mark.parametrize("seed", range(8))
def test_t2_equals_t1_on_adversarial_descriptors(oracle, reference, synth, seed):
    """Inputs built to sit on the hazards of SURVEY App. B: descriptors of extreme entries (+-127/-128: both
    int32 products at tracking_main.c:154 wrap, scores come out negative, > 1, inf or NaN), runs of identical
    descriptors (exact score ties: the first candidate in x-outer / y-inner order must win), all-zero
    descriptors at the head of a window (the sticky 256-d branch of squared_dist) and low-entropy logits
    (probability ties at the top-N threshold)."""
    from adversarial import adversarial_pair
    scale, s0, d0, s1, d1 = adversarial_pair(oracle, synth, seed)
    a = reference.top_n(scale, s1, 100)
    b = oracle.top_n(scale, s1, 100)
    assert all((x == y).all() for x, y in zip(a, b[:2])) and (bits(a[2]) == bits(b[2])).all()
    res = reference.tracking_main(scale, s0, d0, scale, s1, d1)
    m = _t2_pair(oracle, scale, s0, d0, scale, s1, d1)
    assert res["n"] == m["n"]
    assert (res["pts0"] == m["pts0"]).all() and (res["pts1"] == m["pts1"]).all()
    E, inl, ni, wrote = oracle.ransac_identity(m["pts0"], m["pts1"])
    assert ni == res["num_inliers"] and (inl == res["inliers"]).all()


@pytest.mark.parametrize("seed", range(6))
def test_t2_equals_t1_on_adversarial_logits(oracle, reference, seed):
    """Detector inputs on the order-dependent parts of top_N.c: equal maxima inside a cell (the strict `>`
    at :39 keeps the first), logits of 127 under large scales (the int32 powers of approx_exp at their
    largest, exponentials up to ~1e9), cells whose only non-negative entry is the dustbin, all-negative
    cells (denominator FLT_MIN), zero logits, and a few distinct probabilities shared by hundreds of cells
    (the `prob >= threshold` cut at :116-133 then depends on patch order alone)."""
    from adversarial import adversarial_logits
    semi = adversarial_logits(seed)
    for scale in (0.01, 0.35622698, 1.0, 3.0):
        i1, p1, n1 = reference.softmax(scale, semi)
        i2, p2, n2 = oracle.softmax(scale, semi)
        assert n1 == n2 and (i1 == i2).all() and (bits(p1) == bits(p2)).all()
        for N in (100, 37, 1):
            a = reference.top_n(scale, semi, N)
            b = oracle.top_n(scale, semi, N)
            assert len(a[0]) == len(b[0]) and b[3] == 0
            assert all((x == y).all() for x, y in zip(a, b[:2])) and (bits(a[2]) == bits(b[2])).all()


def test_t2_equals_t1_shuffled_real_descriptors(oracle, reference, image0):
    # real descriptors (int32 wrap territory), frame 1 = frame 0 shifted by the search offset
    semi, desc = image0["semi"], image0["desc"]
    g = desc.reshape(80, 24, 256)
    s = semi.reshape(80, 24, 65)
    d1 = np.roll(g, (-4, -4), axis=(0, 1)).reshape(1920, 256).copy()
    s1 = np.roll(s, (-4, -4), axis=(0, 1)).reshape(1920, 65).copy()
    res = reference.tracking_main(image0["semi_scale"], semi, desc, image0["semi_scale"], s1, d1)
    m = _t2_pair(oracle, image0["semi_scale"], semi, desc, image0["semi_scale"], s1, d1)
    assert res["n"] == m["n"] > 0
    assert (res["pts0"] == m["pts0"]).all() and (res["pts1"] == m["pts1"]).all()


def test_ransac_and_error_vs_reference(oracle, reference):
    rng = np.random.default_rng(3)
    p1 = rng.integers(0, 640, size=(150, 2)).astype(np.float32)
    p2 = p1 + rng.integers(-1, 2, size=(150, 2)).astype(np.float32)
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)
    E1, inl1, n1 = reference.ransac(p1, p2, K)
    E2, inl2, n2, wrote = oracle.ransac_identity(p1, p2)
    assert n1 == n2 > 0 and (inl1 == inl2).all() and (E1 == E2).all()
    for i in range(10):
        a = reference.lib.compute_reprojection_error(p1[i], p2[i], E1)
        b = oracle.lib.orc_reproj_error(p1[i], p2[i], E2)
        assert a == b


@pytest.mark.parametrize("n,frac_inl", [(1, 1.0), (7, 0.5), (150, 0.0), (999, 1.0), (1000, 1.0), (1001, 1.0),
                                        (1500, 0.8), (4096, 0.3), (4096, 1.0)])
def test_ransac_cap_and_empty_cases_vs_reference(oracle, reference, n, frac_inl):
    """pnp_solver.c:110-165 beyond tracking_main's 150 points: the 1000-entry inlier array (:146 stops
    counting there, :152 replaces the best on every iteration that reaches it), fewer points than the 8
    draws, and no inlier at all (the reference writes nothing: *num_inliers keeps the caller's value; the
    restatement reports wrote = 0 and the library's defined result is E = I, 0 inliers)."""
    rng = np.random.default_rng(n)
    p1 = (rng.random((n, 2)) * np.array([1241, 376])).astype(np.float32)
    off = np.where(rng.random((n, 1)) < frac_inl, rng.random((n, 2)) * 1.4 - 0.7, 3.0 + rng.random((n, 2)))
    p2 = (p1 + off.astype(np.float32)).astype(np.float32)
    K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]], np.float32)
    for iters, thr in ((10, 1.1), (1, 0.25), (3, 1e9)):
        E1, inl1, n1 = reference.ransac(p1, p2, K, iters, thr)
        E2, inl2, n2, wrote = oracle.ransac_identity(p1, p2, iters, thr)
        if n1 < 0:                      # the reference never wrote its outputs
            assert wrote == 0 and n2 == 0
            continue
        assert wrote == 1 and n1 == n2 and (inl1 == inl2).all() and (E1 == E2).all()
        assert n1 <= 1000 and (thr < 1e8 or n1 == min(n, 1000))


def test_svd3_vs_reference_random(oracle, reference):
    rng = np.random.default_rng(11)
    for i in range(50):
        A = (rng.normal(size=(3, 3)) * 10 ** rng.uniform(-2, 2)).astype(np.float32)
        for x, y in zip(oracle.svd3(A), reference.svd3(A)):
            assert (bits(x) == bits(y)).all()


@pytest.mark.parametrize("seed,permille", [(1, 140), (2, 400), (3, 700), (4, 1000), (5, 60)])
def test_nms_restatement_equals_reference_run_nms(oracle, reference, seed, permille):
    """orc_nms vs the reference's own run_nms.c main() (compiled unmodified into T1): the same
    suppression events in the same order and the same surviving keypoints."""
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import synth
    s, _, _ = synth.synth_frame(seed, 24, 80, 0, 0, 0, permille)
    ev, kp = reference.run_nms(float(synth.SEMI_SCALE), s)
    idx, pr, _ = oracle.softmax(float(synth.SEMI_SCALE), s)
    mi, pp, n, ev2 = oracle.nms(24, 80, idx, pr, want_events=True)
    cells = np.nonzero(mi != 64)[0]
    kp2 = np.stack([(cells // 24) * 8 + mi[cells] % 8, (cells % 24) * 8 + mi[cells] // 8], axis=1)
    assert n == len(ev) and np.array_equal(ev, ev2) and np.array_equal(kp, kp2)
    assert (pp[(mi == 64) & (idx != 64)] == 64.0).all()


def test_nms_on_reference_fixture(oracle, reference, image0):
    ev, kp = reference.run_nms(image0["semi_scale"], image0["semi"])
    idx, pr, _ = oracle.softmax(float(image0["semi_scale"]), image0["semi"])
    mi, pp, n, ev2 = oracle.nms(24, 80, idx, pr, want_events=True)
    assert n == len(ev) > 0 and np.array_equal(ev, ev2)
