#!/bin/bash
# `ncu --set full` capture of the matcher kernels (k_nn_tc<1>, k_nn_tc<2>, k_nn_resolve) on a 30k x 30k FGINN match of SIFT-like u8
# descriptors (SURVEY 8d NN micro-benchmark shape); second call of the same match, raw page only travels back.
tag=${1:-r2}
mkdir -p gpurun_out /tmp/ncu
cat > /tmp/ncu_nn.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests"))
import numpy as np, mods_b200 as mb
import synth
nt = nq = 30000
t_desc, _ = synth.random_descriptors(nt, 1)
q_desc, src = synth.random_descriptors(nq, 2, dup_of=t_desc, dup_frac=0.4, jitter=4)
txy = np.random.default_rng(3).uniform(0, 4096, size=(nt, 2))
ctx = mb.Context(0)
for _ in range(2): a = ctx.match_fginn(q_desc, t_desc, txy)
print("tentatives", len(a))
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k 'regex:^(k_nn_tc|k_nn_resolve)' -s 4 -c 4 -o /tmp/ncu/full_${tag}_nn python /tmp/ncu_nn.py > gpurun_out/full_${tag}_nn.log 2>&1
ncu -i /tmp/ncu/full_${tag}_nn.ncu-rep --page raw --csv > gpurun_out/full_${tag}_nn_raw.csv 2>/dev/null
tail -2 gpurun_out/full_${tag}_nn.log
