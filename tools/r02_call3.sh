#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== k_step timeline"; timeout 300 python tools/step_trace.py 2>&1 | tee gpurun_out/r02_step_trace.txt
echo "== ncu k_step"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 600 -c 1 -o gpurun_out/r02_k_step python tools/profile_step.py --steps 603 > gpurun_out/p.log 2>&1; tail -2 gpurun_out/p.log
