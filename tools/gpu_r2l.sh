#!/bin/bash
tag=${1:-r2l}
mkdir -p gpurun_out
echo "== training-path tests"; timeout 600 python -m pytest tests/test_conv3x3.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
echo "== config 4 variants"
timeout 900 python - <<'PY' 2>&1 | grep -v Warning | tail -12
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch as th
import bench
class A: pass
a = A()
sec = bench.run_secondary(a, th, None, th.device("cuda", 0), 0, 1)
print(json.dumps(sec["config4_train_step"], indent=1))
print(sec.get("error_config34"))
open("gpurun_out/r2l_secondary.json", "w").write(json.dumps(sec))
PY
