#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== kernel times default"; timeout 300 python tools/kernel_times.py 2>&1 | tail -2
echo "== kernel times area256"; timeout 300 python tools/kernel_times.py 1024 256 2>&1 | tail -2
echo "== kernel times view15"; timeout 300 python tools/kernel_times.py 4096 64 15 128 2>&1 | tail -2
echo "== A/B default"; timeout 600 python tools/ab_knobs.py - - 2>&1
for c in area256; do
  timeout 900 python bench.py --config $c --steps 500 --no-cpu-baseline > gpurun_out/r02_bench_$c.json 2>> gpurun_out/r02_bench.err
  python -c "
import json,sys;d=json.load(open('gpurun_out/r02_bench_$c.json'));print({k:d[k] for k in ('value','ms_per_step','value_warm_l2')}, 'e2e', d['e2e']['value'], {k:round(v*1e3,1) for k,v in d['kernels_ms_in_graph'].items()})"
done
