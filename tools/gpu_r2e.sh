#!/bin/bash
tag=${1:-r2e}
mkdir -p gpurun_out
echo "== trace build"
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" timeout 200 python tools/chain_trace.py 2>&1 | tail -100 | tee gpurun_out/${tag}_chain_trace.txt
echo "== no-store build"
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_NOSTORE" python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 2>&1 | grep "^{" | cut -c1-260 | tee gpurun_out/${tag}_chains_v3_nostore.jsonl
python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
