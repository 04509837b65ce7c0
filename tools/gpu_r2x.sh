#!/bin/bash
tag=${TAG:-r2x}
mkdir -p gpurun_out
echo "== train kernels"; timeout 600 python benchmarks/train_kernels_bench.py 2>&1 | grep -v Warn | tail -40
echo "== config 4 graph variants"
timeout 900 python - <<'PY' 2>&1 | grep -v Warning | tail -12
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch as th
from sbmc_b200 import interfaces, models, train_pipeline
import bench
dev = th.device("cuda", 0)
batch = {"radiance": th.rand(8, 8, 3, 128, 128, device=dev),
         "features": th.randn(8, 8, 93, 128, 128, device=dev),
         "global_features": th.randn(8, 3, 1, 1, device=dev),
         "target_image": th.rand(8, 3, 128, 128, device=dev)}
res = {}
for label, own, graph in (("pipeline_eager", True, False), ("pipeline_graph", True, True),
                          ("pipeline_graph_cudnn_wgrad3x3", False, True)):
    train_pipeline.OWN_WGRAD3X3 = own
    th.manual_seed(0)
    net = models.Multisteps(93, 3).to(dev).train()
    net.bf16_train = True
    iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True, cuda_graph=graph)
    res[label] = bench._timed_cuda(th, lambda: iface.train_step(batch), 2, 10)
    del net, iface
train_pipeline.OWN_WGRAD3X3 = True
print(res)
open("gpurun_out/%s_cfg4.json" % os.environ.get("TAG", "r2x"), "w").write(json.dumps(res))
PY
echo "== profile"
timeout 600 python tools/profile_train.py bf16_train 2>&1 | grep -v Warn | cut -c1-70,150-215 | head -60 > gpurun_out/${tag}_profile.txt
head -45 gpurun_out/${tag}_profile.txt
