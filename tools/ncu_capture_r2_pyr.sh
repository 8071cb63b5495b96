#!/bin/bash
# `ncu --set full` capture of the round-2 scale-space kernels (TMA-staged blur + Hessian + in-level extremum test, k_nms_finish) on one
# 4096x3072 HessianAffine detection (second call: buffers allocated); raw page only travels back.
tag=${1:-r2}
mkdir -p gpurun_out /tmp/ncu
cat > /tmp/ncu_pyr.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, mods_b200 as mb
from mods_b200 import synth
W, H = 4096, 3072
cache = "/tmp/ncu_img_%dx%d.npy" % (W, H)
A = np.load(cache) if os.path.exists(cache) else synth.blob_image(W, H, seed=1, n_blobs=int(1.5e-3 * W * H))
if not os.path.exists(cache): np.save(cache, A)
ctx = mb.Context(0)
for _ in range(2): k = ctx.hessaff_detect(A)
print("keys", len(k))
PY
python /tmp/ncu_pyr.py > /dev/null 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k 'regex:^(k_blur_hess_tma|k_nms_finish|k_hessian|k_resize_half)' -s 52 -c 7 -o /tmp/ncu/full_${tag}_pyr python /tmp/ncu_pyr.py > gpurun_out/full_${tag}_pyr.log 2>&1
ncu -i /tmp/ncu/full_${tag}_pyr.ncu-rep --page raw --csv > gpurun_out/full_${tag}_pyr_raw.csv 2>/dev/null
tail -2 gpurun_out/full_${tag}_pyr.log
