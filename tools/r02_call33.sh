#!/bin/bash
mkdir -p gpurun_out
echo "== A/B view ahead / frame order"
python tools/ab_knobs.py - CRAFTER_B200_VIEW_AHEAD=0 CRAFTER_B200_FRAME_ORDER=0 - CRAFTER_B200_VIEW_AHEAD=0 2>&1 | tee gpurun_out/r02_ab_view_ahead.txt
AB_CONFIG=view15 python tools/ab_knobs.py - CRAFTER_B200_VIEW_AHEAD=0 CRAFTER_B200_FRAME_ORDER=0 2>&1 | tee -a gpurun_out/r02_ab_view_ahead.txt
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_VIEW_AHEAD=0 CRAFTER_B200_FRAME_ORDER=0 2>&1 | tee -a gpurun_out/r02_ab_view_ahead.txt
echo "== kernel times"
python tools/kernel_times.py 2>&1 | tail -1 | tee gpurun_out/r02_kernel_times_view_ahead.txt
echo "== timeline"
python tools/render_trace.py 2>&1 | tee gpurun_out/r02_render_timeline_view_ahead.txt | tail -36
echo "== gpu tests"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests_view_ahead.txt
