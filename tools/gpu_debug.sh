#!/bin/bash
mkdir -p gpurun_out
echo "== pytest conv1x1"
timeout 900 python -m pytest tests/test_conv1x1.py -m gpu -q -x 2>&1 | tail -25
echo "== chains"
timeout 600 python benchmarks/model_bench.py chains --steps 5 2>&1 | grep "^{\|Error\|error" | tee gpurun_out/r1o_chains.json | cut -c1-500
echo "== cfg3 forward with bf16 chains"
timeout 900 python benchmarks/model_bench.py forward --bf16-chains --variants fused 2>&1 | grep "^{\|Error\|error" | tee gpurun_out/r1o_cfg3_chains.json | cut -c1-700
