#!/usr/bin/env python
"""Per-kernel times of the training pipeline at BASELINE config 4 shapes (B=8, spp=8,
128x128): the 1x1 GEMM layer, the 1x1 / 3x3 weight-gradient kernels (against cuBLAS /
cuDNN bf16 on the same operands) and the memory-bound passes, with the HBM or tensor
bound of each.  One JSON line per kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

from sbmc_b200 import train_ops as T

BF = th.bfloat16
HBM = 6551e9
TENSOR = 1644e12


NCU = "--ncu" in sys.argv          # one launch per kernel, no library references


def timed(fn, warm=3, reps=20):
    if NCU:
        warm, reps = 0, 1
    for _ in range(warm):
        fn()
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    th.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = th.device("cuda", 0)
    bs, spp, h, w = 8, 8, 128, 128
    S, P_ = bs * spp * h * w, bs * h * w
    th.manual_seed(0)
    x = th.randn(S, 128, device=dev).to(BF)
    ctx = th.randn(P_, 128, device=dev).to(BF)
    w1 = th.randn(128, 256, device=dev).to(BF)
    w2 = th.randn(128, 128, device=dev).to(BF)
    b = th.randn(128, device=dev)
    out = []

    def line(name, ms, bytes_=None, flops=None, ref_ms=None, ref=None):
        d = {"kernel": name, "ms": round(ms, 4)}
        if bytes_ is not None:
            d["hbm_bound_ms"] = round(bytes_ / HBM * 1e3, 4)
            d["frac_of_hbm_bound"] = round(bytes_ / HBM * 1e3 / ms, 3)
        if flops is not None:
            d["TFLOP/s"] = round(flops / ms / 1e9, 1)
            d["frac_of_bf16_peak"] = round(flops / ms / 1e9 / (TENSOR / 1e12), 3)
        if ref_ms is not None:
            d[ref] = round(ref_ms, 4)
        out.append(d)
        print(json.dumps(d), flush=True)

    line("linear 128->128 + bias + relu (S rows)", timed(lambda: T.linear(x, w2, b, 1)),
         bytes_=S * 256 * 2)
    line("linear [x | ctx] 256->128, second source per pixel", timed(
        lambda: T.linear(x, w1, b, 1, xb=ctx, hw=h * w, spp=spp)), bytes_=S * 256 * 2 + P_ * 256)
    line("linear 128->128 with activation-derivative mask (data gradient)", timed(
        lambda: T.linear(x, w2, None, 0, mask=x, mask_act=1)), bytes_=S * 256 * 3)
    logits = th.empty(spp, bs, 441, h * w, device=dev)
    w3 = th.randn(512, 128, device=dev).to(BF)
    b3 = th.randn(512, device=dev)
    line("linear 128->441 fp32 logits planes", timed(lambda: T.linear(
        x, w3, b3, 0, hw=h * w, spp=spp, out_mode=2, out=logits, out_img_stride=441 * h * w,
        out_smp_stride=bs * 441 * h * w, cout_valid=441)), bytes_=S * (256 + 441 * 4))
    dy = th.randn(S, 128, device=dev).to(BF)
    t_ref = None if NCU else timed(lambda: th.matmul(dy.t(), x))
    line("wgrad 1x1 128x128 over S rows (+ bias gradient)", timed(lambda: T.wgrad(dy, x)),
         bytes_=S * 512, ref_ms=t_ref, ref="cublas_bf16_matmul_ms")
    dy5 = th.randn(S, 512, device=dev).to(BF)
    t_ref = None if NCU else timed(lambda: th.matmul(dy5.t(), x))
    line("wgrad 1x1 512x128 over S rows (regressor)", timed(lambda: T.wgrad(dy5, x, cout_valid=441)),
         bytes_=S * (1024 + 256), ref_ms=t_ref, ref="cublas_bf16_matmul_ms")
    del dy5
    g = th.randn(bs, 441, h * w, device=dev)
    rows = th.empty(bs, spp, h * w, 512, device=dev, dtype=BF)
    line("planes_to_rows 441 fp32 planes -> 512 bf16 channels (one sample)", timed(
        lambda: T.planes_to_rows(g, 512, out=rows[:, 0], out_img_stride=spp * h * w * 512)),
         bytes_=bs * h * w * (441 * 4 + 1024))
    del rows, g, logits
    line("spp_reduce (mean over 8 samples)", timed(lambda: T.spp_reduce(x, bs, spp, 1.0 / spp)),
         bytes_=S * 256 + P_ * 256)
    line("bcast_add", timed(lambda: T.bcast_add(x, ctx, bs, spp, 0.125)), bytes_=S * 512 + P_ * 256)
    for hh, cin, cout in ((128, 128, 128), (128, 384, 128), (64, 128, 256), (64, 256, 256),
                          (64, 768, 256), (32, 256, 512), (32, 512, 512)):
        xx = th.randn(bs, hh, hh, cin, device=dev).to(BF)
        dp = th.randn(bs, hh, hh, cout, device=dev).to(BF)
        flops = 2.0 * 9 * cin * cout * bs * hh * hh
        t_ref = None if NCU else timed(lambda: th.nn.grad.conv2d_weight(
            xx.permute(0, 3, 1, 2), (cout, cin, 3, 3), dp.permute(0, 3, 1, 2), padding=1))
        line("wgrad3x3 %d->%d @ %dx%d x8 (+ bias gradient)" % (cin, cout, hh, hh),
             timed(lambda: T.wgrad3x3(dp, xx, want_bias=True)),
             flops=flops, ref_ms=t_ref, ref="cudnn_bf16_wgrad_ms")
        w9 = th.randn(9, cout, cin, device=dev).to(BF)
        zb = th.zeros(cout, device=dev)
        line("conv3x3 fwd %d->%d @ %dx%d x8" % (cin, cout, hh, hh),
             timed(lambda: T.conv3x3(xx, w9, zb, 1)), flops=flops)
    with open(os.path.join("gpurun_out", os.environ.get("TAG", "r2x") + "_train_kernels.jsonl"), "w") as f:
        for d in out:
            f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
