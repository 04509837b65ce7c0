#!/usr/bin/env python
"""Developer tool: torch-profiler timeline of the end-to-end loop in steady state -- when do
the loader's kernels (inflate, assembly) run relative to the training step's kernels?"""
import os, sys, time, tempfile, shutil, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
from torch.profiler import profile, ProfilerActivity
from sbmc_b200 import datasets, interfaces, models
from tests import tile_io

root = tempfile.mkdtemp(prefix="sbmc_probe_")
try:
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    tile_io.write_scene(root, "scene", np.random.default_rng(0), 128, 6, 6, 8, quantize=1.0 / 256, compress=compress)
    files = sorted(os.listdir(os.path.join(root, "scene")))
    with open(os.path.join(root, "list.txt"), "w") as fid:
        fid.write("\n".join(os.path.join("scene", f) for f in files * 40) + "\n")
    data = datasets.TilesDataset(os.path.join(root, "list.txt"), spp=8)
    dev = th.device("cuda", 0)
    net = models.Multisteps(data.num_features, data.num_global_features).to(dev).train()
    net.bf16_train = True
    iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True, cuda_graph=True)
    it = iter(datasets.PrefetchLoader(data, batch_size=8, shuffle=True, drop_last=True, device_prefetch=16))
    for i in range(80):
        iface.train_step(next(it))
    th.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(40):
            iface.train_step(next(it))
        th.cuda.synchronize()
    it.close()
    ev = [e for e in prof.events() if e.device_type == th.autograd.DeviceType.CUDA]
    t0 = min(e.time_range.start for e in ev)
    infl = [(e.time_range.start - t0, e.time_range.end - t0) for e in ev if "lz4_frames" in e.name]
    asm = [(e.time_range.start - t0, e.time_range.end - t0) for e in ev if "tile_assemble" in e.name]
    memcpy = [(e.time_range.start - t0, e.time_range.end - t0, e.name) for e in ev if "Memcpy HtoD" in e.name and e.time_range.end - e.time_range.start > 5000]
    train = sorted((e.time_range.start - t0, e.time_range.end - t0) for e in ev
                   if any(k in e.name for k in ("conv3x3", "linear_kernel", "wgrad", "splat")))
    print("profiled span %.1f ms, %d training kernels, busy %.1f ms" % (
        (max(e.time_range.end for e in ev) - t0) / 1e3, len(train), sum(b - a for a, b in train) / 1e3))
    for a, b in infl:
        inside = sum(max(0, min(b, d) - max(a, c)) for c, d in train)
        n_in = sum(1 for c, d in train if c >= a and d <= b)
        print("inflate kernel %.1f .. %.1f ms (%.1f ms): %d training kernels inside it, %.1f ms of training-kernel time"
              % (a / 1e3, b / 1e3, (b - a) / 1e3, n_in, inside / 1e3))
    for a, b in asm:
        print("assemble kernel %.1f .. %.1f ms" % (a / 1e3, b / 1e3))
    for a, b, n in memcpy[:6]:
        print("H2D copy %.1f .. %.1f ms (%.1f ms)" % (a / 1e3, b / 1e3, (b - a) / 1e3))
    # gaps of the training stream
    gaps = [(train[i + 1][0] - train[i][1], train[i][1]) for i in range(len(train) - 1)]
    big = sorted(gaps, reverse=True)[:5]
    print("largest gaps between training kernels (ms, at ms):", [(round(g / 1e3, 1), round(t / 1e3, 1)) for g, t in big])
    iface.close()
finally:
    shutil.rmtree(root, ignore_errors=True)
