#!/bin/bash
mkdir -p gpurun_out
echo "== sweep dws"
timeout 300 tools/sweep dws 2>&1 | grep -v "^mem" | tee gpurun_out/r1u_sweep.txt
echo "== tiled 4K"
for n in 1 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/tiled_inference.py --fast 2>&1 | grep "^{\|Error\|error" | tee gpurun_out/r1u_tiled$n.json | head -5
done
