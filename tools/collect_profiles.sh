#!/bin/bash
# Run on the GPU box (gpurun): tests, bench line, warm launch list, full-set captures of the step's
# kernels at steady state.  Outputs land in gpurun_out/; summaries are copied into profiles/ by
# tools/summarise_profiles.py in the build container.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench_reference.json
python tools/kernel_times.py > gpurun_out/${TAG}_kernel_times.txt 2>&1; tail -2 gpurun_out/${TAG}_kernel_times.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 450 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --steps 460 > gpurun_out/l.log 2>&1; tail -1 gpurun_out/l.log
for spec in "k_render 400" "k_update 400" "k_wg_mat 400" "k_post 400"; do set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/${TAG}_$1 python tools/profile_step.py --steps $(($2 + 2)) > gpurun_out/p.log 2>&1; tail -1 gpurun_out/p.log
done
