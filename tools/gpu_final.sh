#!/bin/bash
# Final-state verification: what the driver runs at round end, plus ncu of the final kernels.
tag=${1:-r1x}
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
echo "== bench"; timeout 900 python bench.py 2> gpurun_out/${tag}_bench.err | grep "^{" | tee gpurun_out/${tag}_bench.json | cut -c1-300
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 | tee gpurun_out/${tag}_bench_ref.json | cut -c1-200
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kw_|s2g' -c 60 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kw_' -c 3 -f -o gpurun_out/${tag}_prof \
  python bench.py --steps 1 --warmup 0 --spp 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log
echo "== cfg3 / cfg4"
timeout 600 python benchmarks/model_bench.py forward --bf16-unet --bf16-chains --variants fused 2>&1 | grep "^{" | tee gpurun_out/${tag}_cfg3_fast.json | cut -c1-400
timeout 600 python benchmarks/model_bench.py train 2>&1 | grep "^{" | tee gpurun_out/${tag}_cfg4.json | cut -c1-400
