#!/bin/bash
# Round 2, 2 GPUs: the multi-GPU bench line (per-rank e2e timing, no barrier inside the window) and the gather's cost.
set -u
mkdir -p gpurun_out
P=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 500 --warmup 50 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_n2.json'));print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])"; tail -2 gpurun_out/r02_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) tools/gather_cost.py 2> gpurun_out/r02_gather.err | tee gpurun_out/r02_gather_cost_n2.json; tail -2 gpurun_out/r02_gather.err
echo "== single-GPU tests added this round"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
