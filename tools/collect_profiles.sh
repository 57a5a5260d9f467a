#!/bin/bash
# Run on the GPU box (gpurun): tests, the bench lines of the three BASELINE configs, the reference arm, the
# driver's short run, per-kernel times inside the graph, the tick / balance timelines, a launch list and
# full-set captures of the step's kernels at steady state.  Outputs land in gpurun_out/; summaries are
# copied into profiles/ by tools/summarise_profiles.py in the build container.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_gpu_tests.txt
for c in default area256 view15; do
  timeout 900 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2>> gpurun_out/${TAG}_bench.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_$c.json'));print('$c', {k:d[k] for k in ('value','ms_per_step','value_warm_l2')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])"
done
cp gpurun_out/${TAG}_bench_default.json gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_short.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench_reference.json; echo
timeout 300 python tools/kernel_times.py > gpurun_out/${TAG}_kernel_times.txt 2>&1; tail -1 gpurun_out/${TAG}_kernel_times.txt
timeout 300 python tools/kernel_times.py 1024 256 > gpurun_out/${TAG}_kernel_times_area256.txt 2>&1; tail -1 gpurun_out/${TAG}_kernel_times_area256.txt
timeout 300 python tools/kernel_times.py 4096 64 15 128 > gpurun_out/${TAG}_kernel_times_view15.txt 2>&1; tail -1 gpurun_out/${TAG}_kernel_times_view15.txt
timeout 300 python tools/balance_trace.py > gpurun_out/${TAG}_tick_balance_timeline.txt 2>&1; tail -3 gpurun_out/${TAG}_tick_balance_timeline.txt
timeout 300 python tools/render_trace.py > gpurun_out/${TAG}_render_timeline.txt 2>&1; grep -E "===|CTA start" gpurun_out/${TAG}_render_timeline.txt | tail -4
timeout 300 python tools/branch_trace.py > gpurun_out/${TAG}_branch_timeline.txt 2>&1; tail -9 gpurun_out/${TAG}_branch_timeline.txt
timeout 300 python tools/config_sweep.py > gpurun_out/${TAG}_config_sweep.jsonl 2>&1; tail -1 gpurun_out/${TAG}_config_sweep.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6400 -c 450 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --steps 960 > gpurun_out/l.log 2>&1; tail -1 gpurun_out/l.log
for k in k_render k_update k_post k_wg_mat; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 900 -c 1 -o gpurun_out/${TAG}_$k python tools/profile_step.py --steps 903 > gpurun_out/p.log 2>&1; tail -1 gpurun_out/p.log
done
