"""Workload for `ncu --set full` captures of the verification and DoG kernels: one batched scorer call of the SURVEY 8d micro-benchmark
shape (n = 8192 correspondences x K = 4096 hypotheses, HDs then FDs) and one DoG detection on a 4096x3072 image.
Usage: ncu --set full --clock-control none --import-source on -k regex:^(k_score|k_wide_rows|k_dog_cols) -c 6 -o gpurun_out/x python tools/ncu_target_verify.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mods_b200 as mb
from mods_b200 import synth

ctx = mb.Context(0)
rng = np.random.default_rng(1)
n, K = 8192, 4096
u = np.zeros((n, 6)); u[:, 0:2] = rng.random((n, 2)) * 1000; u[:, 2] = 1; u[:, 5] = 1
Hgt = synth.gt_homography(1000, 1000)
p = (Hgt @ u[:, 0:3].T).T; u[:, 3:5] = p[:, :2] / p[:, 2:3] + rng.normal(size=(n, 2))
u[int(0.6 * n):, 3:5] = rng.random((n - int(0.6 * n), 2)) * 1000
models = np.stack([np.linalg.inv(Hgt).T.ravel() * (1 + 1e-3 * rng.normal(size=9)) for _ in range(K)])
for which in (0, 3):
    I, J = ctx.score_models(which, u, models, 9.0)
    print("which", which, "mean inliers", I.mean())
W, H = 4096, 3072
A = synth.blob_image(W, H, seed=1, n_blobs=int(1.5e-3 * W * H))
k = ctx.hessaff_detect(A, mb.HessaffParams.dog(), capacity=2000000)
print("DoG keys", len(k))
