#!/bin/bash
# compute-sanitizer passes over GPU parity tests (run on the GPU box): eager launches so that every
# kernel is instrumented individually.  Output: gpurun_out/<tag>_sanitizer.txt
TAG=${1:-r01}
OUT=gpurun_out/${TAG}_sanitizer.txt
mkdir -p gpurun_out
export CRAFTER_B200_NO_GRAPH=1
echo "compute-sanitizer ($(nvcc --version | grep -o 'release [0-9.]*')) on the B200 box, eager launches (CRAFTER_B200_NO_GRAPH=1)" > $OUT
echo "--- memcheck" >> $OUT
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "auto_reset or compaction or render_subset or episode_recorder or state_dict" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|\[100%\]" >> $OUT
echo "--- racecheck" >> $OUT
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "auto_reset and (default_short or big_area)" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard|\[100%\]" | head -40 >> $OUT
echo "--- synccheck" >> $OUT
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "auto_reset and (default_short or big_area)" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|\[100%\]" >> $OUT
cat $OUT
