#!/bin/bash
tag=${1:-r2v}
mkdir -p gpurun_out
echo "== training kernels + pipeline tests"
timeout 900 python -m pytest tests/test_train_ops.py tests/test_train_pipeline.py tests/test_chain_train.py -m gpu -q 2>&1 | grep -E "passed|failed|Error|^E  |assert " | head -40
echo "== config 4 bf16_train timing"
timeout 600 python - <<'PY' 2>&1 | grep -v Warning | tail -20
import json, sys, os, time
sys.path.insert(0, os.getcwd())
import torch as th
from sbmc_b200 import interfaces, models
dev = th.device("cuda", 0)
res = {}
for mode in ("bf16_train",):
    th.manual_seed(0)
    net = models.Multisteps(93, 3).to(dev).train()
    net.bf16_train = True
    iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True)
    batch = {"radiance": th.rand(8, 8, 3, 128, 128, device=dev),
             "features": th.randn(8, 8, 93, 128, 128, device=dev),
             "global_features": th.randn(8, 3, 1, 1, device=dev),
             "target_image": th.rand(8, 3, 128, 128, device=dev)}
    for _ in range(3):
        iface.backward(batch, iface.forward(batch))
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out = iface.backward(batch, iface.forward(batch))
    b.record(); th.cuda.synchronize()
    res[mode] = a.elapsed_time(b) / 10
    print(mode, res[mode], out)
open("gpurun_out/%s_cfg4.json" % os.environ.get("TAG", "r2v"), "w").write(json.dumps(res))
PY
echo "== profile"
timeout 600 python tools/profile_train.py bf16_train 2>&1 | grep -v Warn | cut -c1-70,150-215 | head -60 > gpurun_out/${tag}_profile.txt
head -50 gpurun_out/${tag}_profile.txt
