#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r1d_pytest.txt
echo "== sweep"
for w in dda dws s2g splat; do timeout 300 tools/sweep $w 2>&1 | grep -v "^mem"; done | tee gpurun_out/r1d_sweep.txt
