#!/usr/bin/env python
"""Developer tool: timeline of CTA 0 of the pipelined chain kernel.
Build with SBMC_B200_NVCC_FLAGS=-DSBMC_CHAIN_TRACE first (not the default build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from sbmc_b200 import _lib, conv1x1, modules

NAMES = {1: "prod: F slot free, loads issued", 10: "mma: waiting for F", 11: "mma: F landed",
         12: "mma: X_a free -> L1a", 13: "mma: X_b free -> L1b", 14: "mma: E1a done -> L2a",
         15: "mma: E1b done -> L2b", 16: "mma: E2a done -> L3a", 17: "mma: E2b done -> L3b",
         20: "eg a: L1 acc ready", 21: "eg a: E1 done", 22: "eg a: L2 acc ready", 23: "eg a: E2 done",
         24: "eg a: L3 acc ready", 25: "eg a: E3 done",
         30: "eg b: L1 acc ready", 31: "eg b: E1 done", 32: "eg b: L2 acc ready", 33: "eg b: E2 done",
         34: "eg b: L3 acc ready", 35: "eg b: E3 done"}
dev = th.device("cuda", 0)
lib = _lib.load()
trace = th.zeros(4 * 2000 * 2, dtype=th.int64, device=dev)
ctypes.c_void_p.in_dll(lib, "sbmc_b200_debug_pointer").value = trace.data_ptr()
h, w, spp = 720, 1280, 4
emb = modules.ConvChain(256, 128, width=128, depth=3, ksize=1, pad=False).to(dev).eval()
feats = th.randn(1, spp, h * w, 128, device=dev).to(th.bfloat16)
prop = th.randn(1, h * w, 128, device=dev).to(th.bfloat16)
with th.no_grad():
    for _ in range(2):
        conv1x1.chain_samples_nhwc(emb, feats, 128, prop=prop, want_mean=True)
    th.cuda.synchronize()
    trace.zero_()
    conv1x1.chain_samples_nhwc(emb, feats, 128, prop=prop, want_mean=True)
    th.cuda.synchronize()
t = trace.cpu().view(4, 2000, 2).tolist()
ev = sorted((c, code) for role in t for code, c in role if c)
t0 = ev[0][0]
# print items 10..13 (steady state)
start = [i for i, (c, code) in enumerate(ev) if code == 10]
lo, hi = start[20], start[24]
prev = ev[lo][0]
for c, code in ev[lo:hi]:
    print("%8d  (+%5d)  %s" % (c - ev[lo][0], c - prev, NAMES.get(code, code)))
    prev = c
print("items traced: %d, cycles per pair-item (steady): %.0f" % (len(start), (ev[start[-2]][0] - ev[start[10]][0]) / (len(start) - 12)))
