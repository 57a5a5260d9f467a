#!/bin/bash
mkdir -p gpurun_out
echo "== A/B frame order (night frames first)"
python tools/ab_knobs.py - CRAFTER_B200_FRAME_ORDER=0 - CRAFTER_B200_FRAME_ORDER=0 2>&1 | tee gpurun_out/r02_ab_frame_order_night_first.txt
AB_CONFIG=view15 python tools/ab_knobs.py - CRAFTER_B200_FRAME_ORDER=0 2>&1 | tee -a gpurun_out/r02_ab_frame_order_night_first.txt
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_FRAME_ORDER=0 2>&1 | tee -a gpurun_out/r02_ab_frame_order_night_first.txt
echo "== kernel times"
python tools/kernel_times.py 2>&1 | tail -1 | tee gpurun_out/r02_kernel_times_night_first.txt
echo "== timeline"
python tools/render_trace.py 2>&1 | tee gpurun_out/r02_render_timeline_night_first.txt | grep -E "===|CTA start"
echo "== gpu tests"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests_night_first.txt
