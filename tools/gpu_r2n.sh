#!/bin/bash
tag=${1:-r2n}
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_conv1x1.py tests/test_conv3x3.py -m gpu -q -x 2>&1 | tail -3
echo "== chains v3 bench"; timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_chains_v3.jsonl | cut -c1-200
echo "== convs bench"; timeout 300 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs.jsonl | cut -c1-200
echo "== convs bench, pair"; SBMC_B200_CONV_PAIR=1 timeout 300 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | head -2 | tee gpurun_out/${tag}_convs_pair.jsonl | cut -c1-200
echo "== cfg3 forward"
timeout 600 python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 5 --warmup 2 2>&1 | grep "^{" | tee gpurun_out/${tag}_cfg3.json | cut -c1-400
