#!/usr/bin/env python
"""Top CUDA kernels of the config-3 forward (torch profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from torch.profiler import profile, ProfilerActivity
from sbmc_b200 import models

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = "cuda"
th.manual_seed(0)
net = models.Multisteps(93, 3).to(dev).eval()
if "unet" in mode:
    net.bf16_unet = True
    net = net.to(memory_format=th.channels_last)
if "chains" in mode:
    net.bf16_chains = True
bs, spp, h, w = 1, 4, 720, 1280
batch = {"radiance": th.rand(bs, spp, 3, h, w, device=dev),
         "features": th.randn(bs, spp, 93, h, w, device=dev),
         "global_features": th.randn(bs, 3, 1, 1, device=dev)}
with th.no_grad():
    net(batch); net(batch)
    th.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(batch)
        th.cuda.synchronize()
print("MODE", mode)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
