"""Workload for `ncu` captures of the MSER kernels: ONE 4096x3072 image (both polarities) through mb2_mser_detect, twice (the first call
allocates).  Usage: ncu --set full --clock-control none --import-source on -k regex:k_mtree -s <skip> -c <n> -o x python tools/ncu_target_mser.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mods_b200 as mb
from mods_b200 import synth

W, H = 4096, 3072
cache = "/tmp/ncu_img_%dx%d.npy" % (W, H)
if os.path.exists(cache):
    A = np.load(cache)
else:
    A = synth.blob_image(W, H, seed=1, n_blobs=int(1.5e-3 * W * H)); np.save(cache, A)
ctx = mb.Context(0)
for _ in range(2):
    k = ctx.mser_detect(A)
print("regions", len(k))
