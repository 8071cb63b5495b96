"""One 4096x3072 pair with the [HessianAffine4] + [MSER2] view tiers through the C++ pair driver (mb2_mods_pair): the single-GPU
time to compare tools/views_sharded.py against."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mods_b200 as mb
from mods_b200 import synth
w, h = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096x3072").split("x"))
A = synth.blob_image(w, h, seed=1, n_blobs=int(1.5e-3 * w * h)); B = synth.warp_image(A, synth.gt_homography(w, h), seed=2)
ctx = mb.Context(0)
cfg = mb.PairConfig.default(); cfg.use_mser = 1
hess = [(1, 0, 1, 0.2)] + [(t, k * np.pi / n, 1, 0.2) for t, n in ((2, 1), (4, 2), (6, 3), (8, 4)) for k in range(n)]
mser = [(1, 0, 1, 0.8), (1, 0, 0.25, 0.8), (1, 0, 0.125, 0.8)]
cfg.set_views(hess, mser)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
for it in range(3):
    t0 = time.perf_counter(); res, _ = ctx.mods_pair(dA, dB, cfg, shape1=(h, w), shape2=(h, w)); dt = time.perf_counter() - t0
    print("views %d+%d: %.1f ms  regions %d/%d tentatives %d verified %d  (detect %.1f match %.1f dup %.1f ransac %.1f)" % (
        len(hess), len(mser), 1e3 * dt, res.regions1, res.regions2, res.tentatives, res.verified, res.ms_detect_describe, res.ms_match, res.ms_duplicate, res.ms_ransac))
