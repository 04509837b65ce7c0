#!/bin/bash
tag=${1:-r2p}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
echo "== convs bench"; timeout 300 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs.jsonl | cut -c1-200
echo "== cfg3 forward"
timeout 600 python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 5 --warmup 2 2>&1 | grep "^{" | tee gpurun_out/${tag}_cfg3.json | cut -c1-600
echo "== trace build"
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" timeout 200 python tools/chain_trace.py 2>&1 | tail -75 | tee gpurun_out/${tag}_chain_trace.txt
python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
