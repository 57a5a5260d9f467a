#!/bin/bash
mkdir -p gpurun_out
P=crafter_b200/_lib/variants/libcrafter_b200_prev.so
echo "== A/B day runs (prev = HEAD before)"
python tools/ab_knobs.py - CRAFTER_B200_LIB=$P - CRAFTER_B200_LIB=$P 2>&1 | tee gpurun_out/r02_ab_day_runs.txt
AB_CONFIG=view15 python tools/ab_knobs.py - CRAFTER_B200_LIB=$P 2>&1 | tee -a gpurun_out/r02_ab_day_runs.txt
AB_CONFIG=area256 python tools/ab_knobs.py - CRAFTER_B200_LIB=$P 2>&1 | tee -a gpurun_out/r02_ab_day_runs.txt
echo "== kernel times"
python tools/kernel_times.py 2>&1 | tail -1 | tee gpurun_out/r02_kernel_times_day_runs.txt
echo "== gpu tests"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests_day_runs.txt
echo "== render alone"
python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(d['value'], d['value_warm_l2'], d['roofline']['ms_per_launch_alone_warm_l2'], d['roofline']['ms_per_launch_in_graph'])"
