"""Generates the committed golden fixtures from the reference tree (run in the build
container, where /root/reference exists; the GPU box only reads the .npz files).

    python tests/golden/make_golden.py

Writes
  ref_image0.npz   the reference's own quantized SuperPoint output for KITTI image 0
                   (include/data/quantized/quantized_image0.h: semi int8 [1920,65], desc int8
                   [1920,256], scales) and its float-softmax ground truth
                   (include/data/quantized/pair0_gt.h: image0_indices_gt, image0_probs_gt)
  ref_kat.npz      known answers produced by RUNNING the unmodified reference sources
                   (oracle/_ref/libmaveric_ref.so, built by oracle/Makefile) on that input and
                   on synthetic pairs: compute_softmax, compute_top_N, the tracking_main match
                   list, RANSAC inliers, pose-from-E, svd3 on random matrices, matmul/matmul2.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def parse_array(text, name, dtype):
    m = re.search(r"%s\s*(\[[^=]*\])\s*=\s*\{(.*?)\};" % re.escape(name), text, re.S)
    dims = [int(d) for d in re.findall(r"\[(\d+)\]", m.group(1))]
    vals = np.array(re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(2)), dtype=np.float64)
    return vals.astype(dtype).reshape(dims)


def parse_scalar(text, name):
    return float(re.search(r"%s\s*=\s*([-0-9.eE+]+)\s*;" % re.escape(name), text).group(1))


def main():
    from oracle import orc
    import maveric_slam_b200  # noqa: F401
    from maveric_slam_b200 import synth

    q = open(os.path.join(REF, "include/data/quantized/quantized_image0.h")).read()
    gt = open(os.path.join(REF, "include/data/quantized/pair0_gt.h")).read()
    semi = parse_array(q, "image0_semi", np.int8)
    desc = parse_array(q, "image0_desc", np.int8)
    semi_scale = np.float32(parse_scalar(q, "image0_semi_scale"))
    desc_scale = np.float32(parse_scalar(q, "image0_desc_scale"))
    idx_gt = parse_array(gt, "image0_indices_gt", np.int32)      # [80][24] = [col][row]
    probs_gt = parse_array(gt, "image0_probs_gt", np.float32)
    assert semi.shape == (1920, 65) and desc.shape == (1920, 256) and idx_gt.shape == (80, 24)
    np.savez_compressed(os.path.join(OUT, "ref_image0.npz"), semi=semi, desc=desc, semi_scale=semi_scale,
                        desc_scale=desc_scale, indices_gt=idx_gt.reshape(-1), probs_gt=probs_gt.reshape(-1))

    r = orc.Reference()
    kat = {}
    idx, pr, nv = r.softmax(semi_scale, semi)
    kat["img0_softmax_idx"], kat["img0_softmax_prob"], kat["img0_num_valid"] = idx, pr, np.int32(nv)
    pa, ix, pp = r.top_n(semi_scale, semi, 100)
    kat["img0_top100_patch"], kat["img0_top100_idx"], kat["img0_top100_prob"] = pa, ix, pp
    # self pair (image1 := image0), SURVEY §8c (iii)
    res = r.tracking_main(semi_scale, semi, desc, semi_scale, semi, desc)
    kat["self_pts0"], kat["self_pts1"] = res["pts0"], res["pts1"]
    kat["self_num_inliers"], kat["self_inliers"] = np.int32(res["num_inliers"]), res["inliers"]
    R1, R2, t = r.recover_pose(np.eye(3, dtype=np.float32))
    kat["pose_R1"], kat["pose_R2"], kat["pose_t"] = R1, R2, t

    # synthetic pairs at the reference's native shape, several seeds
    for seed in range(4):
        off = synth.default_offsets(2, seed)
        s0, d0, _ = synth.synth_frame(seed, 24, 80, 0, int(off[0, 0]), int(off[0, 1]))
        s1, d1, _ = synth.synth_frame(seed, 24, 80, 1, int(off[1, 0]), int(off[1, 1]))
        res = r.tracking_main(synth.SEMI_SCALE, s0, d0, synth.SEMI_SCALE, s1, d1)
        kat[f"syn{seed}_pts0"], kat[f"syn{seed}_pts1"] = res["pts0"], res["pts1"]
        kat[f"syn{seed}_num_inliers"] = np.int32(res["num_inliers"])

    rng = np.random.default_rng(7)
    mats = rng.normal(size=(16, 3, 3)).astype(np.float32)
    mats[0] = np.eye(3)
    mats[1] = np.diag([3.0, 2.0, 0.0])
    usv = np.zeros((16, 3, 3, 3), np.float32)
    for i in range(16):
        usv[i] = np.stack(r.svd3(mats[i]))
    kat["svd_in"], kat["svd_usv"] = mats, usv

    A = rng.normal(size=(7, 5)).astype(np.float32)
    B = rng.normal(size=(5, 6)).astype(np.float32)
    Cm = rng.normal(size=(7, 6)).astype(np.float32)
    Cout = Cm.copy()
    r.lib.matmul(7, 6, 5, A, B, Cout, 5, 6, 6, 0.5, 1.25, False, False)
    kat["mm_A"], kat["mm_B"], kat["mm_C0"], kat["mm_C1"] = A, B, Cm, Cout
    At = np.ascontiguousarray(A.T)
    Bt = np.ascontiguousarray(B.T)
    C2 = np.zeros((7, 6), np.float32)
    r.lib.matmul2(7, 6, 5, At, Bt, Cm.ctypes.data, C2, 7, 5, 6, 6, 1.5, -0.75, 2.0, True, True)
    kat["mm2_C"] = C2
    np.savez_compressed(os.path.join(OUT, "ref_kat.npz"), **kat)
    print("wrote", sorted(kat.keys()))
    print("self pair matches:", len(kat["self_pts0"]), "num_valid:", nv, "top100:", len(pa))


if __name__ == "__main__":
    main()
