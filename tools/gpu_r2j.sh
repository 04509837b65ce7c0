#!/bin/bash
tag=${1:-r2j}
mkdir -p gpurun_out
echo "== conv tests (pair kernel for Cout=128)"; timeout 300 python -m pytest tests/test_conv3x3.py -m gpu -q -x 2>&1 | tail -4
echo "== layout / pool tests"; timeout 300 python -m pytest tests/test_conv1x1.py -m gpu -q -x -k "to_nhwc or pipelined" 2>&1 | tail -2
echo "== convs bench, CTA pairs"; timeout 300 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs_pair.jsonl | cut -c1-200
echo "== convs bench, single CTA"; SBMC_B200_CONV_PAIR=0 timeout 300 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | head -2 | tee gpurun_out/${tag}_convs_single.jsonl | cut -c1-200
echo "== cfg3 forward"
timeout 600 python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 5 --warmup 2 2>&1 | grep "^{" | tee gpurun_out/${tag}_cfg3.json | cut -c1-700
