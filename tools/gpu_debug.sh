#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_timing.py 2>&1 | grep -v "^W\|^\[W\|OMP_NUM\|\*\*\*" | tee gpurun_out/r1j_shard_timing.txt
