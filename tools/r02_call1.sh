#!/bin/bash
# Round 2, GPU call 1: A/B of every schedule / knob / build variant written blind at the end of round 1
# (steady state: 700 pre-roll steps, best of 3 x 1000), in-situ kernel times, launch list, and the
# attempt to obtain the real `opensimplex` package on the box.
set -u
mkdir -p gpurun_out
echo "== pip opensimplex"; (timeout 120 python -m pip install --target /tmp/osx opensimplex; echo "exit=$?"; timeout 60 python -m pip download -d /tmp/osx_dl opensimplex; echo "exit=$?") > gpurun_out/r02_pip_opensimplex.log 2>&1; tail -3 gpurun_out/r02_pip_opensimplex.log
V=crafter_b200/_lib/variants
echo "== A/B"
timeout 1200 python tools/ab_knobs.py - CRAFTER_B200_DRAW_PREFETCH=0 CRAFTER_B200_INCR_CENSUS=0 CRAFTER_B200_SPLIT=1 CRAFTER_B200_DEFER_WG=1 \
  CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_SPLIT=1 \
  CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_FUSED=1 CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_FUSED=2 \
  CRAFTER_B200_LIB=$V/libcrafter_b200_upd2.so CRAFTER_B200_LIB=$V/libcrafter_b200_upd8.so \
  CRAFTER_B200_LIB=$V/libcrafter_b200_bal256.so CRAFTER_B200_LIB=$V/libcrafter_b200_wg4.so \
  CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_LIB=$V/libcrafter_b200_upd2.so CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_LIB=$V/libcrafter_b200_upd8.so \
  CRAFTER_B200_NO_SPECIALIZE=1 - \
  2>&1 | tee gpurun_out/r02_ab_knobs.txt
echo "== kernel times (eager, in situ)"; timeout 300 python tools/kernel_times.py 2>&1 | tee gpurun_out/r02_kernel_times.txt
echo "== bench (old protocol, long)"; timeout 600 python bench.py --steps 2000 --warmup 700 --no-cpu-baseline > gpurun_out/r02_bench_pre.json 2> gpurun_out/r02_bench_pre.err; tail -c 900 gpurun_out/r02_bench_pre.json
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 450 --csv --log-file gpurun_out/r02_launches_head.csv python tools/profile_step.py --steps 760 > gpurun_out/l.log 2>&1; tail -1 gpurun_out/l.log
echo "== config sweep"; timeout 400 python tools/config_sweep.py 2>&1 | tee gpurun_out/r02_config_sweep_pre.jsonl
