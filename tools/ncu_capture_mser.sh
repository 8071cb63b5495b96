#!/bin/bash
# `ncu --set full` capture of the MSER / view-synthesis kernels added after the r1 a/b/c captures (run on the GPU box through gpurun).
# Only the raw CSV pages travel back (the .ncu-rep files exceed gpurun's 64 MiB return limit).
tag=${1:-r1f}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k 'regex:^(k_mser_tree|k_mser_down|k_mser_emulate|k_mser_regions_b|k_mser_runs|k_mser_prep|k_mser_moments)' -c 7 -o /tmp/ncu/full_${tag}_d python tools/ncu_target.py > gpurun_out/full_${tag}_d.log 2>&1
ncu -i /tmp/ncu/full_${tag}_d.ncu-rep --page raw --csv > gpurun_out/full_${tag}_d_raw.csv 2>/dev/null
timeout 600 $NCU -k 'regex:^(k_warp_affine|k_blur101_rows|k_blur101_cols)' -c 4 -o /tmp/ncu/full_${tag}_e python tools/ncu_target_synth.py > gpurun_out/full_${tag}_e.log 2>&1
ncu -i /tmp/ncu/full_${tag}_e.ncu-rep --page raw --csv > gpurun_out/full_${tag}_e_raw.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out/full_${tag}_*
