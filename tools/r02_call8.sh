#!/bin/bash
# Round 2, GPU call 8: split render over compact work lists; A/B against one render launch.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests.txt
echo "== A/B"
timeout 1200 python tools/ab_knobs.py - CRAFTER_B200_SPLIT=0 CRAFTER_B200_OBS_EVICT_FIRST=0 CRAFTER_B200_SPLIT=0,CRAFTER_B200_OBS_EVICT_FIRST=0 CRAFTER_B200_NO_GRAPH=1 - 2>&1 | tee gpurun_out/r02_ab_split.txt
echo "== kernel times"; timeout 300 python tools/kernel_times.py 2>&1 | tail -3 | tee gpurun_out/r02_kernel_times_split.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 600 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench.json'));print({k:d[k] for k in ('value','ms_per_step','regime','value_warm_l2')}, d['e2e']['value'], d['roofline']['frac'], d['kernels_ms_in_graph'])"
