#!/bin/bash
tag=${1:-r2w}
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_train_ops.py tests/test_train_pipeline.py tests/test_optim.py -m gpu -q 2>&1 | grep -E "passed|failed|Error|^E  |assert " | head -40
echo "== config 4 variants"
timeout 900 python - <<'PY' 2>&1 | grep -v Warning | tail -20
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch as th
import bench
class A: pass
sec = bench.run_secondary(A(), th, None, th.device("cuda", 0), 0, 1)
c4 = sec["config4_train_step"]
print({k: round(v["ms"], 2) for k, v in c4.items() if isinstance(v, dict)})
print(sec.get("error_config34"))
open("gpurun_out/%s_secondary.json" % os.environ.get("TAG", "r2w"), "w").write(json.dumps(sec))
PY
