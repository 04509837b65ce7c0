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


def bench_chains(a, dev):
    """The per-sample 1x1 chains alone: fused tcgen05 kernel vs the torch ConvChain
    (cuDNN convs + activation kernels) on one [1, C, 720, 1280] sample plane."""
    from sbmc_b200 import conv1x1, modules
    h, w = a.h or 720, a.w or 1280
    chains = {
        "embedding_00 96->128->128->128": modules.ConvChain(96, 128, width=128, depth=3, ksize=1, pad=False),
        "embedding_01 256->128->128->128": modules.ConvChain(256, 128, width=128, depth=3, ksize=1, pad=False),
        "kernel_regressor 256->128->128->441": modules.ConvChain(
            256, 441, depth=3, width=128, ksize=1, activation="leaky_relu", pad=False,
            output_type="linear"),
    }
    for name, chain in chains.items():
        chain = chain.to(dev).eval()
        cin = conv1x1._convs(chain)[0].in_channels
        cout = conv1x1._convs(chain)[2].out_channels
        x = th.randn(1, cin, h, w, device=dev)
        out = th.empty(1, cout, h, w, device=dev)
        macs = h * w * (cin * 128 + 128 * 128 + 128 * cout)
        with th.no_grad():
            ms_f, _, kern = timed(lambda: conv1x1.chain_forward(chain, x, out=out), a.warmup, a.steps)
            ms_t, _, _ = timed(lambda: chain(x), a.warmup, a.steps)
            rel = ((out - chain(x)).norm() / chain(x).norm()).item()
        print(json.dumps({
            "bench": "1x1 chain " + name, "H": h, "W": w,
            "fused_tcgen05_ms": ms_f, "torch_cudnn_fp32_ms": ms_t, "speedup": ms_t / ms_f,
            "fused_TFLOPs": 2 * macs / ms_f / 1e9,
            "fused_GBs_algorithmic": 4.0 * h * w * (cin + cout) / ms_f / 1e6,
            "rel_err_vs_fp32": rel}), flush=True)


def bench_chains_v3(a, dev):
    """The pipelined all-samples chain kernel (csrc/chain_v3.cu) on one 720p image with
    `spp` samples: time per sample plane against the kernel's own HBM roofline and
    against the round-1 serial per-sample kernel on the same bf16 NHWC tensors."""
    from sbmc_b200 import conv1x1, modules
    h, w, spp = a.h or 720, a.w or 1280, a.spp or 4
    hw = h * w
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(
            os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, ValueError, KeyError):
        peak = 6650.0
    emb = modules.ConvChain(256, 128, width=128, depth=3, ksize=1, pad=False).to(dev).eval()
    emb0 = modules.ConvChain(96, 128, width=128, depth=3, ksize=1, pad=False).to(dev).eval()
    reg = modules.ConvChain(256, 441, depth=3, width=128, ksize=1, activation="leaky_relu",
                            pad=False, output_type="linear").to(dev).eval()
    feats = th.randn(1, spp, hw, 128, device=dev).to(th.bfloat16)
    prop = th.randn(1, hw, 128, device=dev).to(th.bfloat16)
    gf = th.randn(1, 3, device=dev)
    out = th.empty_like(feats)
    logits = th.empty(1, spp, 441, hw, device=dev)
    one = th.empty(1, 441, hw, device=dev)
    cases = [
        # name, callable (all samples), serial callable (one sample), bytes per sample plane
        ("embedding_00 96->128^3 (+mean)",
         lambda: conv1x1.chain_samples_nhwc(emb0, feats, 93, gf=gf, out=out, want_mean=True),
         lambda: conv1x1.chain_forward_nhwc(emb0, feats[:, 0], 93, gf=gf, out=out[:, 0]),
         hw * (256 + 256 + 256 / spp)),
        ("embedding_01 256->128^3 (+mean)",
         lambda: conv1x1.chain_samples_nhwc(emb, feats, 128, prop=prop, out=out, want_mean=True),
         lambda: conv1x1.chain_forward_nhwc(emb, feats[:, 0], 128, xb=prop, out=out[:, 0]),
         hw * (256 + 256 + 2 * 256 / spp)),
        ("kernel_regressor 256->128->128->441 (fp32 logits)",
         lambda: conv1x1.chain_samples_nhwc(reg, feats, 128, prop=prop, regress=True, out=logits),
         lambda: conv1x1.chain_forward_nhwc(reg, feats[:, 0], 128, xb=prop, out=one,
                                            nhwc_out=False),
         hw * (256 + 1764 + 256 / spp)),
    ]
    with th.no_grad():
        for name, fused, serial, nbytes in cases:
            ms_f, _, _ = timed(fused, a.warmup, a.steps)
            ms_s, _, _ = timed(serial, a.warmup, a.steps)
            per_plane = ms_f / spp
            print(json.dumps({
                "bench": "pipelined 1x1 chain " + name, "H": h, "W": w, "spp": spp,
                "pipelined_ms_per_sample_plane": per_plane, "serial_round1_ms_per_sample_plane": ms_s,
                "algorithmic_bytes_per_sample_plane": nbytes,
                "hbm_bound_ms": nbytes / peak / 1e6, "frac_of_hbm_bound": nbytes / peak / 1e6 / per_plane,
                "GBps": nbytes / per_plane / 1e6, "peak_GBps": peak}), flush=True)


def bench_convs(a, dev):
    """Every 3x3 convolution shape of the U-net (sbmc/modules.py:278-305) at the three
    resolutions of a 720p image: csrc/conv3x3.cu (bias + activation fused) against
    cuDNN bf16 channels_last + the separate bias / activation pass."""
    from sbmc_b200 import conv3x3, unet_fast
    import torch.nn.functional as F
    h, w = a.h or 720, a.w or 1280
    shapes = [(128, 128, 1), (384, 128, 1), (128, 256, 2), (256, 256, 2), (768, 256, 2),
              (256, 512, 4), (512, 512, 4)]
    total = {"own": 0.0, "cudnn": 0.0}
    counts = {(128, 128, 1): 5, (384, 128, 1): 1, (128, 256, 2): 1, (256, 256, 2): 4,
              (768, 256, 2): 1, (256, 512, 4): 1, (512, 512, 4): 2}
    for cin, cout, div in shapes:
        hh, ww = h // div, w // div
        x = th.randn(1, hh, ww, cin, device=dev).to(th.bfloat16)
        wt = (th.randn(cout, cin, 3, 3, device=dev) / (3 * cin ** 0.5)).to(th.bfloat16)
        bias = th.randn(cout, device=dev)
        w9 = conv3x3.prepare_weight(wt)
        y = th.empty(1, hh, ww, cout, device=dev, dtype=th.bfloat16)
        xc = x.permute(0, 3, 1, 2)
        wc = wt.contiguous(memory_format=th.channels_last)

        def lib():
            z = F.conv2d(xc, wc, None, 1, 1)
            unet_fast._bias_act_(z, bias, 2)
            return z
        ms_o, _, _ = timed(lambda: conv3x3.conv3x3_nhwc(x, w9, bias, 2, out=y), a.warmup, a.steps)
        ms_c, _, _ = timed(lib, a.warmup, a.steps)
        ms_cc, _, _ = timed(lambda: F.conv2d(xc, wc, None, 1, 1), a.warmup, a.steps)
        flops = 2.0 * hh * ww * cin * cout * 9
        rel = ((y.permute(0, 3, 1, 2).float() - lib().float()).norm() / lib().float().norm()).item()
        total["own"] += counts[(cin, cout, div)] * ms_o
        total["cudnn"] += counts[(cin, cout, div)] * ms_c
        print(json.dumps({
            "bench": "conv3x3 %d->%d @ %dx%d" % (cin, cout, ww, hh), "own_ms": ms_o,
            "cudnn_conv_plus_bias_act_ms": ms_c, "cudnn_conv_only_ms": ms_cc,
            "own_TFLOPs": flops / ms_o / 1e9, "cudnn_conv_only_TFLOPs": flops / ms_cc / 1e9,
            "convs_per_unet": counts[(cin, cout, div)], "rel_diff_vs_cudnn": rel}), flush=True)
    print(json.dumps({"bench": "conv3x3: one U-net's 15 convolutions (sum of the above x counts)",
                      "own_ms": total["own"], "cudnn_ms": total["cudnn"]}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["forward", "train", "chains", "chains_v3", "convs"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--spp", type=int)
    ap.add_argument("--h", type=int)
    ap.add_argument("--w", type=int)
    ap.add_argument("--bs", type=int)
    ap.add_argument("--autocast", action="store_true", help="bf16 autocast for the convs")
    ap.add_argument("--variants", default="fused,composed")
    ap.add_argument("--bf16-unet", action="store_true",
                    help="U-nets through cuDNN in bf16 / channels_last")
    ap.add_argument("--bf16-chains", action="store_true",
                    help="fused tcgen05 1x1 chains (embeddings, kernel regressor)")
    a = ap.parse_args()
    dev = th.device("cuda", 0)
    th.manual_seed(0)
    if a.mode == "chains":
        return bench_chains(a, dev)
    if a.mode == "chains_v3":
        return bench_chains_v3(a, dev)
    if a.mode == "convs":
        return bench_convs(a, dev)
    if a.mode == "forward":
        bs, spp, h, w = a.bs or 1, a.spp or 4, a.h or 720, a.w or 1280
    else:
        bs, spp, h, w = a.bs or 8, a.spp or 8, a.h or 128, a.w or 128
    net = models.Multisteps(93, 3).to(dev)
    net.bf16_chains = a.bf16_chains
    net.bf16_unet = a.bf16_unet
    if a.bf16_unet:
        net = net.to(memory_format=th.channels_last)
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
            "chains_1x1": "fused tcgen05 bf16" if a.bf16_chains else "cuDNN",
            "unet": "cuDNN bf16 channels_last" if a.bf16_unet else "cuDNN (as convs)",
            "bs": bs, "spp": spp, "H": h, "W": w, "K": 21,
            "ms_per_step": ms, "Msamples_per_s": samples / ms / 1e3,
            "sbmc_b200_launches_per_step": launches, "sbmc_b200_kernels": kern,
            "peak_mem_GB": th.cuda.max_memory_allocated() / 1e9}), flush=True)


if __name__ == "__main__":
    main()
