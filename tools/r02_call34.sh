#!/bin/bash
mkdir -p gpurun_out
V=crafter_b200/_lib/variants/libcrafter_b200
echo "== A/B k_wg_obj 256 threads (default) vs 1024, side priority"
python tools/ab_knobs.py - CRAFTER_B200_LIB=${V}_obj1024.so CRAFTER_B200_SIDE_PRIO=1 - CRAFTER_B200_LIB=${V}_obj1024.so,CRAFTER_B200_SIDE_PRIO=1 2>&1 | tee gpurun_out/r02_ab_obj_threads.txt
AB_CONFIG=view15 python tools/ab_knobs.py - CRAFTER_B200_LIB=${V}_obj1024.so CRAFTER_B200_SIDE_PRIO=1 2>&1 | tee -a gpurun_out/r02_ab_obj_threads.txt
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_LIB=${V}_obj1024.so CRAFTER_B200_SIDE_PRIO=1 2>&1 | tee -a gpurun_out/r02_ab_obj_threads.txt
echo "== kernel times"
python tools/kernel_times.py 2>&1 | tail -1 | tee gpurun_out/r02_kernel_times_obj256.txt
CRAFTER_B200_SIDE_PRIO=1 python tools/kernel_times.py 2>&1 | tail -1 | tee -a gpurun_out/r02_kernel_times_obj256.txt
echo "== gpu tests (worldgen)"
timeout 900 python -m pytest tests -x -q -m gpu -k "world or reset or parity or golden" 2>&1 | tail -3
