#!/bin/bash
mkdir -p gpurun_out
V=crafter_b200/_lib/variants/libcrafter_b200
echo "== A/B k_wg_mat tile shapes, default config"
python tools/ab_knobs.py - CRAFTER_B200_LIB=${V}_wgc128.so CRAFTER_B200_LIB=${V}_wgc64.so CRAFTER_B200_LIB=${V}_wgc64t128.so CRAFTER_B200_LIB=${V}_wgc32t128.so - 2>&1 | tee gpurun_out/r02_ab_wg_tiles.txt
echo "== area256"
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_LIB=${V}_wgc128.so CRAFTER_B200_LIB=${V}_wgc64.so CRAFTER_B200_LIB=${V}_wgc64t128.so CRAFTER_B200_LIB=${V}_wgc32t128.so 2>&1 | tee -a gpurun_out/r02_ab_wg_tiles.txt
echo "== kernel times"
for v in wgc128 wgc64 wgc64t128 wgc32t128; do echo $v; CRAFTER_B200_LIB=${V}_$v.so python tools/kernel_times.py 2>&1 | tail -1; done | tee gpurun_out/r02_kernel_times_wg_tiles.txt
