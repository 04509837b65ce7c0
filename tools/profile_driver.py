#!/usr/bin/env python
"""Launches each hot kernel once at BASELINE config-2 call size (N=4, 720x1280,
K=21) for ncu captures: python tools/profile_driver.py [s2g|splat|conv|all]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from sbmc_b200 import conv1x1, modules, splat
import sbmc_b200.functions as funcs

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
th.manual_seed(0)
B, C, H, W, K = 4, 3, 720, 1280, 21
if which in ("s2g", "splat", "all"):
    logits = th.randn(B, K * K, H, W, device=dev)
    rad = th.rand(B, C, H, W, device=dev)
if which in ("s2g", "all"):
    funcs.Scatter2Gather.apply(logits.view(B, K, K, H, W))
if which in ("splat", "all"):
    with th.no_grad():
        st = splat.progressive_splat_update(rad, logits, None, None, None, True)
        st = splat.progressive_splat_update(rad, logits, *st, True)
    r = rad.clone().requires_grad_(True)
    l = logits.clone().requires_grad_(True)
    out = splat.ProgressiveSplat.apply(r, l, None, None, None)
    (out[0].sum() + out[1].sum()).backward()
if which in ("conv", "all"):
    reg = modules.ConvChain(256, 441, depth=3, width=128, ksize=1, activation="leaky_relu",
                            pad=False, output_type="linear").to(dev).eval()
    emb = modules.ConvChain(96, 128, width=128, depth=3, ksize=1, pad=False).to(dev).eval()
    with th.no_grad():
        conv1x1.chain_forward(reg, th.randn(1, 256, H, W, device=dev))
        conv1x1.chain_forward(emb, th.randn(1, 96, H, W, device=dev))
th.cuda.synchronize()
print("done")
