#!/usr/bin/env python
"""Developer tool: per-batch timeline of the end-to-end loop (wait for the loader vs train step)."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
from sbmc_b200 import datasets, interfaces, models
from tests import tile_io

root = tempfile.mkdtemp(prefix="sbmc_probe_")
try:
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    tile_io.write_scene(root, "scene", np.random.default_rng(0), 128, 6, 6, 8, quantize=1.0 / 256, compress=compress)
    files = sorted(os.listdir(os.path.join(root, "scene")))
    with open(os.path.join(root, "list.txt"), "w") as fid:
        fid.write("\n".join(os.path.join("scene", f) for f in files * 32) + "\n")
    data = datasets.TilesDataset(os.path.join(root, "list.txt"), spp=8)
    dev = th.device("cuda", 0)
    net = models.Multisteps(data.num_features, data.num_global_features).to(dev).train()
    net.bf16_train = True
    iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True, cuda_graph=True)
    ld = datasets.PrefetchLoader(data, batch_size=8, shuffle=True, drop_last=True, device_prefetch=16)
    ld.background_decode = False
    it = iter(ld)
    rows = []
    for i in range(80):
        t0 = time.perf_counter()
        batch = next(it)
        t1 = time.perf_counter()
        iface.train_step(batch)
        t2 = time.perf_counter()
        rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    it.close()
    for g in range(0, 80, 16):
        w = [r[0] for r in rows[g:g + 16]]; s = [r[1] for r in rows[g:g + 16]]
        print("batches %2d-%2d: wait first %.1f ms, other waits %.2f ms avg; train step %.2f ms avg (max %.1f)"
              % (g, g + 15, w[0], sum(w[1:]) / 15, sum(s) / 16, max(s)), flush=True)
    iface.close()
finally:
    shutil.rmtree(root, ignore_errors=True)
