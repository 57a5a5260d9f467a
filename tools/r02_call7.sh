#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== schedule tests"; timeout 1500 python -m pytest tests/test_zz_schedules_gpu.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "== A/B"
timeout 1200 python tools/ab_knobs.py CRAFTER_B200_QUEUE=0 CRAFTER_B200_QUEUE=0,CRAFTER_B200_PERSIST=0 CRAFTER_B200_QUEUE=0,CRAFTER_B200_OBS_EVICT_FIRST=1 \
  - CRAFTER_B200_PERSIST=0 CRAFTER_B200_PDL=0 CRAFTER_B200_OBS_EVICT_FIRST=1 CRAFTER_B200_QUEUE=0 2>&1 | tee gpurun_out/r02_ab_persist.txt
echo "== timeline (queue)"; timeout 300 python tools/step_trace.py 2>&1 | tee gpurun_out/r02_consume_trace2.txt
echo "== kernel times"; CRAFTER_B200_QUEUE=0 timeout 300 python tools/kernel_times.py 2>&1 | tail -2 | tee gpurun_out/r02_kernel_times_chain2.txt
timeout 300 python tools/kernel_times.py 2>&1 | tail -2 | tee gpurun_out/r02_kernel_times_queue2.txt
