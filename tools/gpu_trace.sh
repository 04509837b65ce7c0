#!/bin/bash
mkdir -p gpurun_out
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" timeout 200 python tools/chain_trace.py 2>&1 | tail -75 | tee gpurun_out/${1:-r2q}_chain_trace.txt
python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
