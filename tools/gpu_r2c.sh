#!/bin/bash
# round 2, session c: ncu of the pipelined chain kernels and the conv3x3 kernel
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_v3 -c 3 -f \
  -o gpurun_out/${tag}_chain_prof python benchmarks/model_bench.py chains_v3 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu_chain.log 2>&1
tail -2 gpurun_out/${tag}_ncu_chain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -c 7 -f \
  -o gpurun_out/${tag}_conv_prof python benchmarks/model_bench.py convs --steps 1 --warmup 0 > gpurun_out/${tag}_ncu_conv.log 2>&1
tail -2 gpurun_out/${tag}_ncu_conv.log
