#!/bin/bash
# perm tables by episode parity: ahead seed beside k_wg_mat.  A/B against the previous build + parity.
mkdir -p gpurun_out
P=crafter_b200/_lib/variants/libcrafter_b200_prev.so
echo "== A/B default"
python tools/ab_knobs.py - CRAFTER_B200_LIB=$P - CRAFTER_B200_LIB=$P 2>&1 | tee gpurun_out/r02_ab_perm2.txt
echo "== A/B area256"
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_LIB=$P 2>&1 | tee -a gpurun_out/r02_ab_perm2.txt
echo "== kernel times"
python tools/kernel_times.py 2>&1 | tail -8 | tee gpurun_out/r02_kernel_times_perm2.txt
echo "== gpu tests"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests_perm2.txt
