"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libmods_ref.so, built in
place from /root/reference by oracle/build_ref.sh).  Run here (the reference is not on the GPU
box); the vectors are committed so that the oracle restatement stays pinned everywhere.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Reference  # noqa: E402
import synth  # noqa: E402


def main():
    R = Reference()
    out = {}
    # (1) synthetic 320x240 view: full per-view pipeline of the reference
    img = synth.blob_image(320, 240, seed=11)
    raw = R.hessaff_detect(img, raw=True)
    reg = R.hessaff_detect(img, raw=False)
    ori = R.detect_orientation(img, reg)
    det, rep, desc = R.view_pipeline(img)
    out.update(s_seed=11, s_raw=raw, s_reg=reg, s_ori=ori, s_det=det, s_rep=rep, s_desc=desc.astype(np.uint8))
    # (2) a 300x300 crop of the reference's own example image (build/examples/cat.png), gray = (B+G+R)/3
    try:
        import cv2
        im = cv2.imread("/root/reference/build/examples/cat.png").astype(np.float32)
        g = ((im[:, :, 0] + im[:, :, 1] + im[:, :, 2]) / 3.0).astype(np.float32)[250:550, 150:450].copy()
        d2, r2, de2 = R.view_pipeline(g)
        out.update(cat_gray=g, cat_det=d2, cat_rep=r2, cat_desc=de2.astype(np.uint8))
        m = R.mser_detect(g)
        out.update(cat_mser=m)
    except Exception as e:  # pragma: no cover
        print("cat crop skipped:", e)
    # (3) DEGENSAC scorers + LO-RANSAC-H with a fixed seed on a planted homography
    rng = np.random.default_rng(7)
    n = 400
    u = np.zeros((n, 6)); u[:, 0:2] = rng.random((n, 2)) * 800; u[:, 2] = 1; u[:, 5] = 1
    Hgt = synth.gt_homography(800, 800)
    p = (Hgt @ u[:, 0:3].T).T; u[:, 3:5] = p[:, :2] / p[:, 2:3] + rng.normal(size=(n, 2))
    u[260:, 3:5] = rng.random((n - 260, 2)) * 800
    M = np.linalg.inv(Hgt).T.ravel()
    out.update(r_u=u, r_M=M, r_scores=np.stack([R.score(w, u, M) for w in range(5)]))
    res = R.exp_ransacH(u, seed=12345)
    out.update(r_H=res["H"], r_inl=res["inl"], r_stats=np.array([res["I"], res["samples"], res["lo"], res["rejected"]]), r_J=res["J"])
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
