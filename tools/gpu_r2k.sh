#!/bin/bash
# round 2, session k: full GPU suite, sanitizer on the new kernels, final ncu captures and bench line
tag=${1:-r2k}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.txt
echo "== compute-sanitizer memcheck on the round-2 kernels"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_conv1x1.py tests/test_conv3x3.py -m gpu -q -x \
  -k "pipelined or conv3x3_matches or maxpool or cta_pair or rewrite" 2>&1 | tail -6 | tee gpurun_out/${tag}_sanitizer.txt
echo "== chains v3 bench"; timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_chains_v3.jsonl | cut -c1-330
echo "== chains v3 bench spp 8"; timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 --spp 8 2>&1 | grep "^{" | tee gpurun_out/${tag}_chains_v3_spp8.jsonl | cut -c1-330
echo "== convs bench"; timeout 300 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs.jsonl | cut -c1-200
echo "== bench.py"; timeout 900 python bench.py 2> gpurun_out/${tag}_bench.err | grep "^{" | tee gpurun_out/${tag}_bench.json | cut -c1-300
echo "== ncu chain"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_v3 -c 3 -f \
  -o gpurun_out/${tag}_chain_prof python benchmarks/model_bench.py chains_v3 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu_chain.log 2>&1
tail -1 gpurun_out/${tag}_ncu_chain.log
echo "== ncu conv"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -c 7 -f \
  -o gpurun_out/${tag}_conv_prof python benchmarks/model_bench.py convs --steps 1 --warmup 0 > gpurun_out/${tag}_ncu_conv.log 2>&1
tail -1 gpurun_out/${tag}_ncu_conv.log
echo "== ncu glue"
timeout 600 ncu --set full --clock-control none -k regex:"upsample_concat|nchw_to_nhwc|maxpool2x2" -c 6 -f \
  -o gpurun_out/${tag}_glue_prof python benchmarks/model_bench.py forward --bf16-chains --bf16-unet --variants fused --steps 1 --warmup 0 > gpurun_out/${tag}_ncu_glue.log 2>&1
tail -1 gpurun_out/${tag}_ncu_glue.log
