#!/bin/bash
# See tools/simt_asan.py.  Output of the last run: profiles/r01_simt_asan.txt
set -e
cd "$(dirname "$0")/.."
cp tests/simt/simt.h /tmp/simt_asan.h
sed "s|#include \"simt.h\"|#include \"/tmp/simt_asan.h\"|; s|#include \"../../crafter_b200|#include \"$PWD/crafter_b200|" tests/simt/simt_env.cpp > /tmp/simt_env_asan.cpp
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared \
  -o /tmp/libsimt_asan.so /tmp/simt_env_asan.cpp -lm
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
  python tools/simt_asan.py
