#!/bin/bash
# One gpurun call: parity tests, tuning sweep, bench, ncu launch list + full capture.
# usage: tools/gpu_round.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt
free -g | head -2 >> gpurun_out/${tag}_gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${tag}_pytest.txt
echo "== sweep"; timeout 600 tools/sweep all > gpurun_out/${tag}_sweep.txt 2>&1; cat gpurun_out/${tag}_sweep.txt
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kw_|s2g' -c 60 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -3 gpurun_out/${tag}_ncu_bench.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kw_' -c 3 -f -o gpurun_out/${tag}_prof \
  python bench.py --steps 1 --warmup 0 --spp 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out
