#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_scenarios_golden.py -m gpu -x -q 2>&1 | tail -3
echo "== A/B default"; timeout 600 python tools/ab_knobs.py - - 2>&1
echo "== kernel times default"; timeout 300 python tools/kernel_times.py 2>&1 | tail -1
echo "== kernel times view15"; timeout 300 python tools/kernel_times.py 4096 64 15 128 2>&1 | tail -1
AB_CONFIG=view15 timeout 600 python tools/ab_knobs.py - 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render -s 900 -c 1 -o gpurun_out/r02_k_render_blit python tools/profile_step.py --steps 903 > gpurun_out/p.log 2>&1; tail -1 gpurun_out/p.log
