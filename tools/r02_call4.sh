#!/bin/bash
# Round 2, GPU call 4: the queue schedule (k_update ~PDL~> k_consume): schedule tests, A/B, timeline, bench.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_gpu_tests.txt
echo "== A/B"
timeout 900 python tools/ab_knobs.py - CRAFTER_B200_PDL=0 CRAFTER_B200_QUEUE=0 CRAFTER_B200_NO_GRAPH=1 CRAFTER_B200_NO_GRAPH=1,CRAFTER_B200_PDL=0 - 2>&1 | tee gpurun_out/r02_ab_queue.txt
echo "== timeline"; timeout 300 python tools/step_trace.py 2>&1 | tee gpurun_out/r02_consume_trace.txt
echo "== kernel times in graph"; timeout 300 python tools/kernel_times.py 2>&1 | tee gpurun_out/r02_kernel_times_queue.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 1200 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
