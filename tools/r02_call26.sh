#!/bin/bash
mkdir -p gpurun_out
P=crafter_b200/_lib/variants/libcrafter_b200_prev.so
echo "== A/B default (prev = before perm tables by parity)"
python tools/ab_knobs.py - - 2>&1 | tee gpurun_out/r02_ab_lazy_octaves.txt
AB_CONFIG=area256 python tools/ab_knobs.py - 2>&1 | tee -a gpurun_out/r02_ab_lazy_octaves.txt
AB_CONFIG=view15 python tools/ab_knobs.py - 2>&1 | tee -a gpurun_out/r02_ab_lazy_octaves.txt
echo "== kernel times"
python tools/kernel_times.py 2>&1 | tail -2 | tee gpurun_out/r02_kernel_times_lazy.txt
python tools/kernel_times.py 1024 256 2>&1 | tail -1 | tee -a gpurun_out/r02_kernel_times_lazy.txt
echo "== gpu tests"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests_lazy.txt
