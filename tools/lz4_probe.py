#!/usr/bin/env python
"""Developer tool: one inflate + assemble of 8 tiles (128 x 128 x 8 spp) for ncu."""
import os, sys, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as th
from sbmc_b200 import datasets
from tests import tile_io
root = tempfile.mkdtemp(prefix="sbmc_probe_")
try:
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    tile_io.write_scene(root, "scene", np.random.default_rng(0), 128, 4, 2, 8, quantize=1.0 / 256, compress=compress)
    data = datasets.TilesDataset(root, spp=8)
    items = data.__getitems__(list(range(8)))
    th.cuda.synchronize()
    print("ok", items[0]["features"].shape)
finally:
    shutil.rmtree(root, ignore_errors=True)
