"""Runs MSER on the bench images of a given rank (seed0 = 1 + 1000 * rank) to look for data-dependent stalls."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mods_b200 as mb
from mods_b200 import synth
rank = int(sys.argv[1]); w, h = 4096, 3072
ctx = mb.Context(0)
for i in range(3):
    s0 = 1 + 1000 * rank + 16 * i
    A = synth.blob_image(w, h, seed=s0, n_blobs=int(1.5e-3 * w * h)); B = synth.warp_image(A, synth.gt_homography(w, h), seed=s0 + 1)
    for name, img in (("A", A), ("B", B)):
        t0 = time.perf_counter(); k = ctx.mser_detect(img, capacity=400000); print("seed %d %s: %d keys %.1f ms" % (s0, name, len(k), 1e3 * (time.perf_counter() - t0)), flush=True)
    t0 = time.perf_counter(); v = ctx.mser_pair_views(A, B); print("  pair views %.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
    cfg = mb.PairConfig.default(); cfg.use_mser = 1
    t0 = time.perf_counter(); r, _ = ctx.mods_pair(A, B, cfg); print("  mods_pair %.1f ms verified %d" % (1e3 * (time.perf_counter() - t0), r.verified), flush=True)
