#!/bin/bash
tag=${1:-r2t}
mkdir -p gpurun_out
echo "== chain training tests"; timeout 600 python -m pytest tests/test_chain_train.py -m gpu -q -x 2>&1 | grep -E "passed|failed|Error|assert|^E " | head -20
echo "== config 4 variants"
timeout 900 python - <<'PY' 2>&1 | grep -v Warning | tail -40
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch as th
import bench
class A: pass
sec = bench.run_secondary(A(), th, None, th.device("cuda", 0), 0, 1)
c4 = sec["config4_train_step"]
print({k: round(v["ms"], 2) for k, v in c4.items() if isinstance(v, dict)})
print(sec.get("error_config34"))
open("gpurun_out/r2t_secondary.json", "w").write(json.dumps(sec))
PY
