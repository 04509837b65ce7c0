#!/usr/bin/env python
"""Secondary benchmarks for the callers of the hot path (BASELINE.json configs 3, 4).

  python benchmarks/model_bench.py forward   # config 3: Multisteps(93,3) eval forward,
                                             #   spp=4, 1280x720, 1 GPU
  python benchmarks/model_bench.py train     # config 4: train step B=8 spp=8 128x128 K=21,
                                             #   fwd + TonemappedRelativeMSE + bwd + clip + Adam

Each mode prints one JSON line per variant:
  "fused"    -- ProgressiveKernelApply through the single-pass sm_100a kernels
  "composed" -- the reference's op chain (Scatter2Gather -> max -> sub_ -> exp_ ->
                KernelWeighting) on the same CUDA ops, i.e. what the reference's
                modules.py does per sample
Random-init weights, synthetic inputs of the reference's shapes (93 sample
features + 3 global features, sbmc/datasets.py:312-354).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th

from sbmc_b200 import _lib, interfaces, models


def timed(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    _lib.timing_collect(); _lib.timing_enable(True)
    l0 = _lib.launch_count()
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    th.cuda.synchronize()
    _lib.timing_enable(False)
    kern = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps}
            for k, v in _lib.timing_collect().items()}
    return a.elapsed_time(b) / steps, (_lib.launch_count() - l0) / steps, kern


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["forward", "train"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--spp", type=int)
    ap.add_argument("--h", type=int)
    ap.add_argument("--w", type=int)
    ap.add_argument("--bs", type=int)
    ap.add_argument("--autocast", action="store_true", help="bf16 autocast for the convs")
    ap.add_argument("--variants", default="fused,composed")
    a = ap.parse_args()
    dev = th.device("cuda", 0)
    th.manual_seed(0)
    if a.mode == "forward":
        bs, spp, h, w = a.bs or 1, a.spp or 4, a.h or 720, a.w or 1280
    else:
        bs, spp, h, w = a.bs or 8, a.spp or 8, a.h or 128, a.w or 128
    net = models.Multisteps(93, 3).to(dev)
    batch = {"radiance": th.rand(bs, spp, 3, h, w, device=dev),
             "features": th.randn(bs, spp, 93, h, w, device=dev),
             "global_features": th.randn(bs, 3, 1, 1, device=dev),
             "target_image": th.rand(bs, 3, h, w, device=dev)}
    samples = bs * spp * h * w
    for variant in a.variants.split(","):
        net.kernel_update.fused = variant == "fused"
        if a.mode == "forward":
            net.eval()

            def step():
                with th.no_grad(), th.autocast("cuda", dtype=th.bfloat16, enabled=a.autocast):
                    return net(batch)["radiance"]
        else:
            net.train()
            iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True)

            def step():
                with th.autocast("cuda", dtype=th.bfloat16, enabled=a.autocast):
                    fwd = iface.forward(batch)
                fwd["radiance"] = fwd["radiance"].float()
                return iface.backward(batch, fwd)
        th.cuda.reset_peak_memory_stats()
        ms, launches, kern = timed(step, a.warmup, a.steps)
        print(json.dumps({
            "bench": "Multisteps(93,3) %s" % ("eval forward (config 3)" if a.mode == "forward"
                                             else "train step (config 4)"),
            "variant": variant, "convs": "bf16 autocast (cuDNN)" if a.autocast else "fp32 (cuDNN)",
            "bs": bs, "spp": spp, "H": h, "W": w, "K": 21,
            "ms_per_step": ms, "Msamples_per_s": samples / ms / 1e3,
            "sbmc_b200_launches_per_step": launches, "sbmc_b200_kernels": kern,
            "peak_mem_GB": th.cuda.max_memory_allocated() / 1e9}), flush=True)


if __name__ == "__main__":
    main()
