#!/usr/bin/env python
"""Developer tool: kernel-time table of one config-4 training step (torch profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from torch.profiler import profile, ProfilerActivity
from sbmc_b200 import interfaces, models

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16_train"
dev = th.device("cuda", 0)
th.manual_seed(0)
net = models.Multisteps(93, 3).to(dev).train()
net.bf16_train = mode == "bf16_train"
net.bf16_unet_train = mode == "bf16_unet"
iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True,
                                                allow_tf32=(mode == "tf32"))
batch = {"radiance": th.rand(8, 8, 3, 128, 128, device=dev),
         "features": th.randn(8, 8, 93, 128, 128, device=dev),
         "global_features": th.randn(8, 3, 1, 1, device=dev),
         "target_image": th.rand(8, 3, 128, 128, device=dev)}
for _ in range(3):
    iface.backward(batch, iface.forward(batch))
th.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    iface.backward(batch, iface.forward(batch))
    th.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
