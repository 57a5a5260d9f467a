#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== schedule tests"; timeout 1500 python -m pytest tests/test_zz_schedules_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== A/B default"; timeout 900 python tools/ab_knobs.py - CRAFTER_B200_FRAME_ORDER=1 - CRAFTER_B200_FRAME_ORDER=1 2>&1 | tee gpurun_out/r02_ab_frame_order.txt
echo "== A/B area256"; AB_CONFIG=area256 timeout 600 python tools/ab_knobs.py - CRAFTER_B200_FRAME_ORDER=1 2>&1 | tee -a gpurun_out/r02_ab_frame_order.txt
echo "== A/B view15"; AB_CONFIG=view15 timeout 600 python tools/ab_knobs.py - CRAFTER_B200_FRAME_ORDER=1 2>&1 | tee -a gpurun_out/r02_ab_frame_order.txt
echo "== kernel times"; CRAFTER_B200_FRAME_ORDER=1 timeout 300 python tools/kernel_times.py 2>&1 | tail -2
