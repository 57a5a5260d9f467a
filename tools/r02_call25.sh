#!/bin/bash
mkdir -p gpurun_out
echo "== A/B priority"
python tools/ab_knobs.py - CRAFTER_B200_PRIO=1 - CRAFTER_B200_PRIO=1 2>&1 | tee gpurun_out/r02_ab_prio.txt
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_PRIO=1 2>&1 | tee -a gpurun_out/r02_ab_prio.txt
AB_CONFIG=view15 python tools/ab_knobs.py - CRAFTER_B200_PRIO=1 2>&1 | tee -a gpurun_out/r02_ab_prio.txt
echo "== kernel times prio"
CRAFTER_B200_PRIO=1 python tools/kernel_times.py 2>&1 | tail -2 | tee gpurun_out/r02_kernel_times_prio.txt
echo "== kernel times area256"
python tools/kernel_times.py 1024 256 2>&1 | tail -3 | tee gpurun_out/r02_kernel_times_area256.txt
echo "== kernel times view15"
python tools/kernel_times.py 4096 64 15 128 2>&1 | tail -3 | tee gpurun_out/r02_kernel_times_view15.txt
