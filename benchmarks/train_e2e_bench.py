#!/usr/bin/env python
"""End-to-end training throughput at BASELINE config 4 through the callers' own loop:
.bin + LZ4 tiles on disk -> PrefetchLoader (GPU inflate + assembly) -> Multisteps ->
loss -> backward -> clip + Adam (scripts/train.py's path; reference: scripts/train.py +
sbmc/datasets.py + sbmc/interfaces.py).

Writes a synthetic scene of 128 x 128 tiles at 8 spp (the reference renderer's format,
real LZ4 frames), then times, on the same batches: the loader alone, the training step
alone on a resident batch, and the two together as `_compat.Trainer` runs them.
Prints one JSON line.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th

from sbmc_b200 import datasets, interfaces, models
from tests import tile_io


def main(argv=None):
    """Prints (and returns) the JSON line; `argv` as on the command line."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=6, help="tiles per side of the synthetic scene")
    ap.add_argument("--ts", type=int, default=128)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--bs", type=int, default=8)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--repeat", type=int, default=56, help="times every tile appears in an epoch")
    ap.add_argument("--prefetch", type=int, default=None,
                    help="PrefetchLoader device_prefetch (batches decoded ahead on a side stream)")
    ap.add_argument("--own_stream", action="store_true",
                    help="run the training loop on a non-default CUDA stream")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph")
    ap.add_argument("--fp32", action="store_true", help="the reference-arithmetic fp32 path")
    ap.add_argument("--quantize", type=float, default=1.0 / 256)
    a = ap.parse_args(argv)
    dev = th.device("cuda", th.cuda.current_device())
    line = None
    root = tempfile.mkdtemp(prefix="sbmc_train_e2e_")
    try:
        rng = np.random.default_rng(0)
        compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
        tile_io.write_scene(root, "scene", rng, a.ts, a.tiles, a.tiles, a.spp, quantize=a.quantize,
                            compress=compress)
        files = sorted(os.listdir(os.path.join(root, "scene")))
        file_bytes = sum(os.path.getsize(os.path.join(root, "scene", f)) for f in files)
        # an epoch of `repeat` passes over the tiles (a file list naming every tile that often),
        # so that the loader reaches its steady state inside one epoch
        with open(os.path.join(root, "list.txt"), "w") as fid:
            fid.write("\n".join(os.path.join("scene", f) for f in files * a.repeat) + "\n")
        data = datasets.TilesDataset(os.path.join(root, "list.txt"), spp=a.spp,
                                     mode=datasets.TilesDataset.SBMC_MODE)
        kw = {} if a.prefetch is None else {"device_prefetch": a.prefetch}

        def loader():
            return datasets.PrefetchLoader(data, batch_size=a.bs, shuffle=True, drop_last=True, **kw)

        if a.own_stream:
            th.cuda.set_stream(th.cuda.Stream(device=dev))
        th.manual_seed(0)
        net = models.Multisteps(data.num_features, data.num_global_features).to(dev).train()
        net.bf16_train = not a.fp32
        iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-4, cuda=True, fused_optimizer=True,
                                                        cuda_graph=not a.eager)

        def epochs(fn, steps, skip=0):
            """Runs fn(batch) over skip + steps batches of ONE pass of the loader (its steady
            state: the first `skip` batches, while the prefetch pipeline fills, are not
            timed) -> seconds for the `steps` batches."""
            done = 0
            t0 = None
            for batch in loader():
                if done == skip:
                    th.cuda.synchronize()
                    t0 = time.perf_counter()
                fn(batch)
                done += 1
                if done >= skip + steps:
                    break
            th.cuda.synchronize()
            assert done == skip + steps, "epoch too short: raise --repeat"
            return time.perf_counter() - t0

        keep = {}

        def remember(batch):
            keep["batch"] = batch
        skip = max(4 * (a.prefetch or 1), 16)   # the pipeline is ahead after ~3 groups (first-touch
        #                                          allocations of the staging / decode buffers)
        epochs(remember, 8)                                   # page cache, allocator, staging
        t_load = epochs(remember, a.steps, skip) / a.steps
        resident = {k: (v.clone() if isinstance(v, th.Tensor) else v) for k, v in keep["batch"].items()}
        for _ in range(3):
            iface.train_step(dict(resident))                  # capture / warm-up
        th.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            iface.train_step(dict(resident))
        th.cuda.synchronize()
        t_step = (time.perf_counter() - t0) / a.steps
        t_e2e = epochs(lambda b: iface.train_step(b), a.steps, skip) / a.steps
        samples = a.bs * a.spp * a.ts * a.ts
        line = {
            "metric": "Msamples/s (B*spp*H*W) training end to end: tiles on disk -> optimizer step",
            "config": {"workload": "BASELINE config 4 through the loader: B=%d, spp=%d, %dx%d tiles, K=21, "
                                   "%d tiles on disk (%.1f MB, LZ4 frames)" % (
                                       a.bs, a.spp, a.ts, a.ts, len(files), file_bytes / 1e6),
                       "path": ("fp32 reference arithmetic" if a.fp32 else "bf16 pipeline")
                       + (", eager" if a.eager else ", one CUDA graph"),
                       "device_prefetch": a.prefetch, "own_stream": a.own_stream},
            "loader_only_ms_per_batch": 1e3 * t_load,
            "step_only_ms": 1e3 * t_step,
            "end_to_end_ms_per_step": 1e3 * t_e2e,
            "end_to_end_Msamples_per_s": samples / t_e2e / 1e6,
            "step_only_Msamples_per_s": samples / t_step / 1e6,
            "loader_only_Msamples_per_s": samples / t_load / 1e6}
        print(json.dumps(line), flush=True)
        iface.close()
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return line


if __name__ == "__main__":
    main()
