#!/usr/bin/env python
"""Developer tool: host-side timeline of the prefetching loader's worker against the consumer."""
import os, sys, time, tempfile, shutil, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
from sbmc_b200 import datasets, interfaces, models
from tests import tile_io

T0 = time.perf_counter()
LOG = []
def log(what):
    LOG.append((1e3 * (time.perf_counter() - T0), threading.current_thread().name, what))

orig_dev, orig_host = datasets.PrefetchLoader._device_half, datasets.PrefetchLoader._host_half
def dev_half(self, batch, planned):
    log("device half start (%d tiles)" % len(batch))
    out = orig_dev(self, batch, planned)
    log("device half end")
    return out
def host_half(self, batch, slot, cuda_device=None):
    log("host half start")
    out = orig_host(self, batch, slot, cuda_device)
    log("host half end")
    return out
datasets.PrefetchLoader._device_half = dev_half
datasets.PrefetchLoader._host_half = host_half

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
    for i in range(112):
        if i % 16 == 0:
            log("consumer asks for batch %d" % i)
        b = next(it)
        if i % 16 == 0:
            log("consumer got batch %d" % i)
        iface.train_step(b)
    it.close()
    for t, name, what in LOG:
        if t > LOG[-1][0] - 1500:
            print("%9.1f ms  %-22s %s" % (t, name, what))
    iface.close()
finally:
    shutil.rmtree(root, ignore_errors=True)
