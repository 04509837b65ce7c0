#!/usr/bin/env python
"""Developer tool: where a PrefetchLoader batch spends its time (host file reads vs
H2D + inflate + assemble + collate) for groups of 8 and 64 tiles of 128 x 128 x 8 spp."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
from torch.utils.data import default_collate
from sbmc_b200 import _lib, datasets
from tests import tile_io

root = tempfile.mkdtemp(prefix="sbmc_probe_")
try:
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    tile_io.write_scene(root, "scene", np.random.default_rng(0), 128, 8, 8, 8, quantize=1.0 / 256, compress=compress)
    data = datasets.TilesDataset(root, spp=8)
    loader = datasets.PrefetchLoader(data, batch_size=8)
    print("cpus", os.cpu_count(), "io threads", datasets._IO_THREADS)
    for ntiles in (8, 64, 8, 64):
        idx = list(range(ntiles))
        t0 = time.perf_counter()
        planned = loader._host_half(idx, 0, th.cuda.current_device())
        t1 = time.perf_counter()
        _lib.timing_collect(); _lib.timing_enable(True)
        items = loader._device_half(idx, planned)
        th.cuda.synchronize()
        t2 = time.perf_counter()
        _lib.timing_enable(False)
        kern = _lib.timing_collect()
        out = [default_collate(items[k:k + 8]) for k in range(0, ntiles, 8)]
        th.cuda.synchronize()
        t3 = time.perf_counter()
        nbytes = planned[0].numel()
        print("tiles %d: host half %.1f ms (%.2f GB/s of %.0f MB), device half %.1f ms (kernels %s), collate %.1f ms"
              % (ntiles, 1e3 * (t1 - t0), nbytes / (t1 - t0) / 1e9, nbytes / 1e6, 1e3 * (t2 - t1),
                 {k: (round(v[0], 2), v[1]) for k, v in kern.items() if v[1]}, 1e3 * (t3 - t2)))
finally:
    shutil.rmtree(root, ignore_errors=True)
