#!/bin/bash
# Round 2, GPU call 9: HEAD's test suite, the three bench configs, per-kernel times of every config, launch list + ncu captures.
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests.txt
for c in default area256 view15; do
  echo "== bench $c"; timeout 900 python bench.py --config $c > gpurun_out/r02_bench_$c.json 2>> gpurun_out/r02_bench.err
  python -c "
import json,sys;d=json.load(open('gpurun_out/r02_bench_$c.json'));print({k:d[k] for k in ('value','ms_per_step','value_warm_l2')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['regime']['resets_per_step'], {k:round(v*1e3,1) for k,v in d['kernels_ms_in_graph'].items()}, 'cpu', d['cpu_baseline']['value'])"
done
echo "== driver-style short run"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_short.json 2>> gpurun_out/r02_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_short.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench_reference.json
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6400 -c 450 --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py --steps 960 > gpurun_out/l.log 2>&1; tail -1 gpurun_out/l.log
for k in k_render k_update k_post k_wg_mat; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 900 -c 1 -o gpurun_out/r02_$k python tools/profile_step.py --steps 903 > gpurun_out/p.log 2>&1; tail -1 gpurun_out/p.log
done
