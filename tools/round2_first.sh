#!/bin/bash
# First GPU call of round 2 (gpurun --timeout 1500 -- 'bash tools/round2_first.sh'): everything that
# was written after round 1's GPU budget ran out, cheapest checks first.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
echo "== experimental schedules, isolated"; timeout 900 python -m pytest tests/test_zz_deferred_gpu.py -m gpu -q -rxX 2>&1 | tail -8 | tee gpurun_out/r02_experimental_tests.txt
echo "== product GPU suite"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_deferred_gpu.py 2>&1 | tail -3 | tee gpurun_out/r02_gpu_tests.txt
echo "== A/B of the knobs (device-timed us/step, bench workload)"
timeout 900 python tools/ab_knobs.py - CRAFTER_B200_DRAW_PREFETCH=0 CRAFTER_B200_INCR_CENSUS=0 CRAFTER_B200_SPLIT=1 CRAFTER_B200_DEFER_WG=1 \
  CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_SPLIT=1 \
  CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_FUSED=1 CRAFTER_B200_DEFER_WG=1,CRAFTER_B200_FUSED=2 \
  2>&1 | tee gpurun_out/r02_ab_knobs.txt
echo "== bench"; timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 700 gpurun_out/r02_bench.json
echo "== area 256x256 (BASELINE configs[3]) with and without the incremental census"
for k in 0 1; do CRAFTER_B200_INCR_CENSUS=$k timeout 300 python tools/config_sweep.py 2>&1 | tail -4 | tee -a gpurun_out/r02_config_sweep_census$k.txt; done
