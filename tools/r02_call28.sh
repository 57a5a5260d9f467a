#!/bin/bash
mkdir -p gpurun_out
echo "== A/B k_wg_mat early exit, grid cap (CTAs per SM)"
python tools/ab_knobs.py - CRAFTER_B200_WG_GRID=3 CRAFTER_B200_WG_GRID=4 CRAFTER_B200_WG_GRID=8 - 2>&1 | tee gpurun_out/r02_ab_wg_grid.txt
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_WG_GRID=3 CRAFTER_B200_WG_GRID=8 2>&1 | tee -a gpurun_out/r02_ab_wg_grid.txt
AB_CONFIG=view15 python tools/ab_knobs.py - CRAFTER_B200_WG_GRID=3 2>&1 | tee -a gpurun_out/r02_ab_wg_grid.txt
echo "== kernel times"
python tools/kernel_times.py 2>&1 | tail -5 | tee gpurun_out/r02_kernel_times_wg_grid.txt
CRAFTER_B200_WG_GRID=3 python tools/kernel_times.py 2>&1 | tail -1 | tee -a gpurun_out/r02_kernel_times_wg_grid.txt
python tools/kernel_times.py 1024 256 2>&1 | tail -1 | tee -a gpurun_out/r02_kernel_times_wg_grid.txt
