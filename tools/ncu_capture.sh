#!/bin/bash
# Round-tagged `ncu --set full` captures of the hot kernels (run on the GPU box through gpurun).
# usage: tools/ncu_capture.sh <tag>   -> gpurun_out/full_<tag>_{a,b,c}.ncu-rep + raw csv pages
tag=${1:-r1}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k 'regex:^(k_extract|k_orientation|k_affine_shape|k_sift_grad|k_sift_votes|k_photonorm_stats|k_sift_finish)' -c 7 -o gpurun_out/full_${tag}_a python tools/ncu_target.py > gpurun_out/full_${tag}_a.log 2>&1
timeout 600 $NCU -k 'regex:^(k_blur_hess|k_nms|k_hessian|k_resize_half|k_localize)' -c 9 -o gpurun_out/full_${tag}_b python tools/ncu_target.py > gpurun_out/full_${tag}_b.log 2>&1
timeout 600 $NCU -k 'regex:^(k_nn_tc|k_score|k_resid)' -c 4 -o gpurun_out/full_${tag}_c python tools/ncu_target.py > gpurun_out/full_${tag}_c.log 2>&1
timeout 600 $NCU -k 'regex:^(k_mser_tree|k_mser_down|k_mser_emulate|k_mser_regions_b|k_mser_runs|k_mser_prep)' -c 6 -o gpurun_out/full_${tag}_d python tools/ncu_target.py > gpurun_out/full_${tag}_d.log 2>&1
for s in a b c d; do ncu -i gpurun_out/full_${tag}_$s.ncu-rep --page raw --csv > gpurun_out/full_${tag}_${s}_raw.csv 2>/dev/null; done
ls -la gpurun_out/full_${tag}_*
