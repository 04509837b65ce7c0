import sys, os
sys.path.insert(0, os.getcwd())
import torch as th
from sbmc_b200 import models
import bench
dev = th.device("cuda", 0)
th.manual_seed(0)
net = models.Multisteps(93, 3).to(dev).eval().to(memory_format=th.channels_last)
net.bf16_chains = net.bf16_unet = True
batch = {"radiance": th.rand(1, 4, 3, 720, 1280, device=dev),
         "features": th.randn(1, 4, 93, 720, 1280, device=dev),
         "global_features": th.randn(1, 3, 1, 1, device=dev)}
def fwd():
    with th.no_grad():
        return net(batch)["radiance"]
print("eager", bench._timed_cuda(th, fwd, 2, 8))
s = th.cuda.Stream(); s.wait_stream(th.cuda.current_stream())
with th.cuda.stream(s):
    fwd(); fwd()
th.cuda.current_stream().wait_stream(s)
g = th.cuda.CUDAGraph()
with th.cuda.graph(g):
    out = fwd()
print("graph", bench._timed_cuda(th, g.replay, 2, 8))
ref = fwd()
g.replay(); th.cuda.synchronize()
print("same", (out - ref).abs().max().item())
