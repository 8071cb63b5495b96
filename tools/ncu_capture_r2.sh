#!/bin/bash
# `ncu --set full` capture of the round-2 component-tree kernels (second mb2_mser_detect call: buffers allocated), raw page only.
tag=${1:-r2}
mkdir -p gpurun_out /tmp/ncu
python tools/ncu_target_mser.py > /dev/null 2>&1   # image cache
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k 'regex:^(k_mtree_|k_mser_emulate|k_mser_sa|k_mser_runs)' -s 8 -c 8 -o /tmp/ncu/full_${tag}_tree python tools/ncu_target_mser.py > gpurun_out/full_${tag}_tree.log 2>&1
ncu -i /tmp/ncu/full_${tag}_tree.ncu-rep --page raw --csv > gpurun_out/full_${tag}_tree_raw.csv 2>/dev/null
tail -3 gpurun_out/full_${tag}_tree.log; ls -la /tmp/ncu
