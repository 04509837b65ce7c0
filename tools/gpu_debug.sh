#!/bin/bash
mkdir -p gpurun_out
echo "== pytest modules"
timeout 900 python -m pytest tests/test_modules.py tests/test_conv1x1.py -m gpu -q -x 2>&1 | tail -5
echo "== cfg3"
timeout 600 python benchmarks/model_bench.py forward --bf16-unet --bf16-chains --variants fused 2>&1 | grep "^{\|Error\|error" | tee gpurun_out/r1v_cfg3_bf16all.json | cut -c1-700
timeout 600 python tools/torch_profile.py unet_chains 2>&1 | grep -v "^$" | cut -c1-200 | tee gpurun_out/r1v_torch_profile.txt | head -24
