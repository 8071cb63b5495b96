"""TEST INFRASTRUCTURE: golden vectors for BASELINE config C1 -- the reference's own example pair build/examples/cat.png vs cat2.png
(598x1000 / 1000x563), the 11-view [HessianAffine4] tier of build/iters_mods_cviu.ini:56-63 (TiltSet 1,2,4,6,8, Phi 360, initSigma 0.2,
RootSIFT, FGINN 0.8), one iteration of mods.cpp:229-415.  Everything is computed by the reference's sources compiled in place
(oracle/_ref): GenerateSynthImageCorr -> DetectAffineKeypoints -> DetectOrientation -> ReprojectRegions -> DescribeRegions per view,
appended in view order (imagerepresentation.cpp:2044-2045), then matching.cpp's MatchFlannFGINN (exact linear k-NN) ->
DuplicateFiltering -> LORANSACFiltering with the RANSAC seed fixed.

Run in the build container (needs /root/reference and cv2):  python tests/golden/make_golden_cat.py
The images travel as the exact integer B+G+R sums (u16); gray = sum / 3 in float32 is what both sides are fed."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

VIEWS = [(1.0, 0.0, 1.0)] + [(float(t), k * np.pi / n, 1.0) for t, n in ((2, 1), (4, 2), (6, 3), (8, 4)) for k in range(n)]   # SetVSPars: n = floor(180 t / Phi)
INIT_SIGMA = 0.2
SEED = 1


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def gray_of(sum_u16):
    return (sum_u16.astype(np.float32) / np.float32(3.0)).astype(np.float32)


def main():
    import cv2
    from oracle.pyoracle import Reference
    R = Reference()
    out = {}
    per_img = []
    for k, name in enumerate(("cat.png", "cat2.png")):
        im = cv2.imread("/root/reference/build/examples/" + name)
        s = im[:, :, 0].astype(np.uint16) + im[:, :, 1].astype(np.uint16) + im[:, :, 2].astype(np.uint16)
        g = gray_of(s)
        ref_gray = ((im[:, :, 0].astype(np.float32) + im[:, :, 1].astype(np.float32) + im[:, :, 2].astype(np.float32)) / 3.0).astype(np.float32)
        assert np.array_equal(g, ref_gray)            # synth-detection.cpp:256-263
        dets, reps, descs, counts = [], [], [], []
        for (tilt, phi, zoom) in VIEWS:
            d, r, de = R.view_pipeline_synth(g, tilt, phi, zoom, detector=0, InitSigma=INIT_SIGMA)
            dets.append(d); reps.append(r); descs.append(de.astype(np.uint8)); counts.append(len(d))
        det, rep, desc = np.concatenate(dets), np.concatenate(reps), np.concatenate(descs)
        per_img.append((rep, desc))
        out["sum%d" % (k + 1)] = s
        out["view_counts%d" % (k + 1)] = np.array(counts, np.int32)
        out["view_digests%d" % (k + 1)] = np.array([digest(a, b, c) for a, b, c in zip(dets, reps, descs)])
        out["first_view_rep%d" % (k + 1)] = reps[0]; out["first_view_desc%d" % (k + 1)] = descs[0]
        print(name, im.shape, "regions per view", counts)
    back = R.pair_back([(per_img[0][0], per_img[0][1].astype(np.float32), per_img[1][0], per_img[1][1].astype(np.float32), 0.8)],
                       contradDist=30.0, duplicateDist=2.0, err_threshold=3.0, confidence=0.99, max_samples=100000, HLAFCoef=12.0, LAFCoef=2.0,
                       errorType=0, doSymmCheck=1, seed=SEED)
    print("tentatives, unique, inliers, verified:", back["counts"])
    rows = back["tent"][back["verified"]]
    ver_xy = np.c_[per_img[0][0][rows[:, 1].astype(int), :2], per_img[1][0][rows[:, 2].astype(int), :2]]
    out.update(views=np.array(VIEWS), init_sigma=INIT_SIGMA, seed=SEED, tent=back["tent"][:, 1:], kept=back["kept"], verified=back["verified"], verified_xy=ver_xy,
               H=back["H"], counts=np.array(back["counts"], np.int32))
    np.savez_compressed(os.path.join(HERE, "cat_pair_vectors.npz"), **out)


if __name__ == "__main__":
    main()
