#!/bin/bash
# `ncu --set full` capture of the description / orientation / Baumberg kernels and the MSER kernels after the tree on ONE 4096x3072 pair
# through mb2_mods_pair (tools/ncu_target.py); raw page only travels back.
tag=${1:-r2}
mkdir -p gpurun_out /tmp/ncu
python tools/ncu_target.py > /dev/null 2>&1   # image cache
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 1200 $NCU -k 'regex:^(k_extract|k_sift_|k_photonorm|k_orientation|k_affine_shape|k_mser_emulate|k_mser_sa|k_mser_regions_b|k_mser_runs|k_mser_ownkeys|k_mtree_local)' -c 60 -o /tmp/ncu/full_${tag}_desc python tools/ncu_target.py > gpurun_out/full_${tag}_desc.log 2>&1
ncu -i /tmp/ncu/full_${tag}_desc.ncu-rep --page raw --csv > gpurun_out/full_${tag}_desc_raw.csv 2>/dev/null
tail -2 gpurun_out/full_${tag}_desc.log
