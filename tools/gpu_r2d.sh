#!/bin/bash
# round 2, session d: re-check the two new kernels after their rewrites, timeline of the
# chain kernel (trace build), full bench.py line with the secondary block
tag=${1:-r2d}
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_conv1x1.py tests/test_conv3x3.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.txt
echo "== chains v3 bench"; timeout 300 python benchmarks/model_bench.py chains_v3 --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_chains_v3.jsonl
echo "== convs bench"; timeout 600 python benchmarks/model_bench.py convs --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/${tag}_convs.jsonl
echo "== bench.py"; timeout 900 python bench.py 2> gpurun_out/${tag}_bench.err | grep "^{" | tee gpurun_out/${tag}_bench.json | cut -c1-400
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench.json").read())
    print(json.dumps(d.get("secondary"), indent=1)[:3000])
except Exception as e:
    print("no bench line", e)
PY
tail -5 gpurun_out/${tag}_bench.err
echo "== trace build"
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
SBMC_B200_NVCC_FLAGS="-DSBMC_CHAIN_TRACE" timeout 200 python tools/chain_trace.py 2>&1 | tail -110 | tee gpurun_out/${tag}_chain_trace.txt
python -c "from sbmc_b200 import build; build.build(force=True)" > /dev/null 2>&1
