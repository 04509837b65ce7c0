import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from sbmc_b200 import conv1x1, modules
dev = "cuda"; th.manual_seed(0)
H, W = 720, 1280
emb = modules.ConvChain(256, 128, width=128, depth=3, ksize=1, pad=False).to(dev).eval()
reg = modules.ConvChain(256, 441, depth=3, width=128, ksize=1, activation="leaky_relu", pad=False, output_type="linear").to(dev).eval()
xa = th.randn(1, H * W, 128, device=dev).to(th.bfloat16)
xb = th.randn(1, H * W, 128, device=dev).to(th.bfloat16)
with th.no_grad():
    for _ in range(2):
        conv1x1.chain_forward_nhwc(emb, xa, 128, xb=xb)
        conv1x1.chain_forward_nhwc(reg, xa, 128, xb=xb, nhwc_out=False)
th.cuda.synchronize(); print("done")
