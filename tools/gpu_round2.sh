#!/bin/bash
# First GPU session of the next round: everything added without a GPU in round 1
# (tile reader kernels, scripts) plus the standing checks, in ONE gpurun call:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh r2a'
tag=${1:-r2a}
mkdir -p gpurun_out
echo "== build + smoke"; timeout 900 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
echo "== tile reader tests (new kernels first, verbose)"
timeout 900 python -m pytest tests/test_tiles.py tests/test_scripts.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/${tag}_pytest_tiles.txt
echo "== sanitizer on the tile kernels"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_tiles.py -m gpu -q -x \
  -k "inflater or fixtures" 2>&1 | tail -8 | tee gpurun_out/${tag}_sanitizer_tiles.txt
echo "== optimizer tests"; timeout 600 python -m pytest tests/test_optim.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest_optim.txt
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.txt
echo "== tiles bench"; timeout 900 python bench.py --workload tiles 2>&1 | grep "^{" | tee gpurun_out/${tag}_tiles_bench.json | cut -c1-600
echo "== ncu tiles"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lz4_frames|tile_assemble' -c 4 -f \
  -o gpurun_out/${tag}_tiles_prof python benchmarks/tiles_bench.py --steps 1 > gpurun_out/${tag}_ncu_tiles.log 2>&1
tail -2 gpurun_out/${tag}_ncu_tiles.log
echo "== bench"; timeout 900 python bench.py 2> gpurun_out/${tag}_bench.err | grep "^{" | tee gpurun_out/${tag}_bench.json | cut -c1-300
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 | tee gpurun_out/${tag}_bench_ref.json | cut -c1-200
echo "== A/B: experimental 16-byte copies in the inflater (not the default build)"
SBMC_B200_NVCC_FLAGS="-DSBMC_LZ4_WIDE_COPY" python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
timeout 600 python -m pytest tests/test_tiles.py -m gpu -q -x -k "inflater or fixtures" 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_tiles_wide.txt
timeout 900 python benchmarks/tiles_bench.py 2>&1 | grep "^{" | tee gpurun_out/${tag}_tiles_bench_wide.json | cut -c1-600
python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1   # back to the default build
