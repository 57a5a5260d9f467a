#!/bin/bash
# ncu --set full with source of the chain's kernels at steady state (step 900: desynchronised)
set -u
mkdir -p gpurun_out
export CRAFTER_B200_QUEUE=0
for k in k_post k_update k_render k_wg_mat; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 900 -c 1 -o gpurun_out/r02_$k python tools/profile_step.py --steps 903 > gpurun_out/p.log 2>&1; tail -1 gpurun_out/p.log
done
