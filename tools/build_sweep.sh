#!/bin/bash
# Builds the tuning sweep (developer tool): tools/sweep
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude \
  -Xptxas=-v tools/sweep.cu sbmc_b200/csrc/runtime.cu sbmc_b200/csrc/generic.cu -o tools/sweep 2> tools/sweep.ptxas.log
grep -E "spill|Used" tools/sweep.ptxas.log | paste - - | grep -v " 0 bytes spill stores, 0 bytes spill loads" || true
