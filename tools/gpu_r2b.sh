#!/bin/bash
# round 2, session b: conv3x3 parity, chain / conv benches, config-3 forward
tag=${1:-r2b}
mkdir -p gpurun_out
echo "== conv3x3 tests"; timeout 600 python -m pytest tests/test_conv3x3.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest_conv3x3.txt
echo "== chains v3 bench"; timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_chains_v3.jsonl
echo "== chains v3 bench spp 8"; timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 --spp 8 2>&1 | grep "^{" | tee gpurun_out/${tag}_chains_v3_spp8.jsonl
echo "== convs bench"; timeout 600 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs.jsonl
echo "== cfg3 forward (pipelined chains + own convs)"
timeout 600 python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 5 --warmup 2 2>&1 | grep "^{" | tee gpurun_out/${tag}_cfg3.json | cut -c1-900
