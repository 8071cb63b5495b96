"""Workload for `ncu --set full` captures: ONE 4096x3072 pair through mb2_mods_pair (single stream).
Usage: ncu --set full --clock-control none --import-source on -k regex:<kernels> -c <n> -o gpurun_out/x python tools/ncu_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mods_b200 as mb
from mods_b200 import synth

W, H = 4096, 3072
cache = "/tmp/ncu_pair_%dx%d.npz" % (W, H)
if os.path.exists(cache):
    z = np.load(cache); A, B = z["A"], z["B"]
else:
    A = synth.blob_image(W, H, seed=1, n_blobs=int(1.5e-3 * W * H))
    B = synth.warp_image(A, synth.gt_homography(W, H), seed=2)
    np.savez(cache, A=A, B=B)
ctx = mb.Context(0)
ctx.profile_begin()          # keeps both images on one stream (the capture is serialised anyway)
cfg = mb.PairConfig.default()
cfg.use_mser = 0 if os.environ.get("NCU_NO_MSER") else 1   # the C3 workload: HessianAffine + MSER
res, _ = ctx.mods_pair(A, B, cfg)
ctx.profile_end()
print("regions %d %d tentatives %d verified %d" % (res.regions1, res.regions2, res.tentatives, res.verified))
