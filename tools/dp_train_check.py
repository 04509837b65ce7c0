#!/usr/bin/env python
"""Developer tool (run under torchrun): data-parallel config-4 training steps, every rank
its own batch, one CUDA graph per rank with the NCCL gradient all-reduce inside; prints the
step time (max over ranks) and whether all ranks hold the same weights, then tears down."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import torch.distributed as dist

from sbmc_b200 import interfaces, models

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
th.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = th.device("cuda", th.cuda.current_device())
dist.init_process_group("nccl", rank=rank, world_size=world)
graph = "--eager" not in sys.argv
th.manual_seed(0)
net = models.Multisteps(93, 3).to(dev).train()
net.bf16_train = True
iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True,
                                                cuda_graph=graph, distributed=True)
g = th.Generator(device=dev).manual_seed(11 + rank)
b4 = {"radiance": th.rand(8, 8, 3, 128, 128, device=dev, generator=g),
      "features": th.randn(8, 8, 93, 128, 128, device=dev, generator=g),
      "global_features": th.randn(8, 3, 1, 1, device=dev, generator=g),
      "target_image": th.rand(8, 3, 128, 128, device=dev, generator=g)}
for _ in range(2):
    iface.train_step(b4)
th.cuda.synchronize()
dist.barrier()
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    iface.train_step(b4)
e1.record()
th.cuda.synchronize()
t = th.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=th.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
chk = th.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"n_gpus": world, "cuda_graph": graph, "ms_per_step": float(t.item()),
                      "Msamples_per_s": world * 8 * 8 * 128 * 128 / float(t.item()) / 1e3,
                      "weights_identical_on_all_ranks": bool((lo == hi).item()),
                      "workload": "config 4 per rank (B=8, spp=8, 128x128, K=21), weak scaling"}), flush=True)
iface.close()
del iface, net
th.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
print("rank %d done" % rank, flush=True)
