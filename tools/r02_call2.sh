#!/bin/bash
# Round 2, GPU call 2: the one-launch tick (k_step) -- parity suite first, then A/B against the chain.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_gpu_tests.txt
V=crafter_b200/_lib/variants
echo "== A/B"
timeout 900 python tools/ab_knobs.py - CRAFTER_B200_STEP_KERNEL=0 CRAFTER_B200_LIB=$V/libcrafter_b200_step4.so CRAFTER_B200_LIB=$V/libcrafter_b200_step6.so \
  CRAFTER_B200_NO_GRAPH=1 CRAFTER_B200_NO_SPECIALIZE=1 - 2>&1 | tee gpurun_out/r02_ab_step_kernel.txt
echo "== kernel times in graph"; timeout 300 python tools/kernel_times.py 2>&1 | tee gpurun_out/r02_kernel_times_kstep.txt
CRAFTER_B200_STEP_KERNEL=0 timeout 300 python tools/kernel_times.py 2>&1 | tee gpurun_out/r02_kernel_times_chain.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 1500 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
echo "== bench, the driver's short run"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_short.json 2>> gpurun_out/r02_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_short.json'));print({k:d[k] for k in ('value','ms_per_step','regime')}, d['e2e']['value'])"
