#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/multi_gpu_check.py 2>&1 | grep "MULTI_GPU\|False\|Error" 
for n in 8 4; do
echo "== bench $n gpus"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline 2> gpurun_out/r1l_bench$n.err | grep "^{" | tee gpurun_out/r1l_bench$n.json | cut -c1-200
done
