#!/usr/bin/env python
"""Benchmark of the tile reader (sbmc_b200.datasets).

  python bench.py --workload tiles [--w 1280 --h 720 --tile-size 80 --spp 8 --steps 5]
  python benchmarks/tiles_bench.py [...]        # the same without the CPU leg

Writes one synthetic scene (tiles in the reference renderer's format, real LZ4
frames) to a scratch folder and times `FullImagesDataset[0]`:
  * "gpu": file bytes (page cache) -> pinned staging -> H2D of the COMPRESSED
    chunks -> warp-per-frame inflate -> assembly kernel -> tensors in HBM; wall
    clock around the item plus the two kernels' device times and their
    algorithmic bytes (inflate: compressed in + inflated out; assembly: inflated
    in + tensors out);
  * "cpu_baseline" (only through bench.py, the one benchmark allowed to execute
    oracle/: the reference algorithm restated, oracle/tiles_ref.py + lz4_oracle.c,
    single thread like the reference's DataLoader with num_workers=0) on a bounded
    number of tiles, scaled to the image.
Prints one JSON line.  Numbers are only meaningful on the GPU box.
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

from sbmc_b200 import _lib, datasets
from tests import tile_io


def main(argv=None, cpu_baseline=None):
    """cpu_baseline(tile file contents) -> seconds per tile (supplied by bench.py)."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=1280)
    ap.add_argument("--h", type=int, default=720)
    ap.add_argument("--ts", type=int, default=80)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu_tiles", type=int, default=4)
    ap.add_argument("--quantize", type=float, default=1.0 / 256,
                    help="value grid of the synthetic floats (controls the compression ratio)")
    a = ap.parse_args(argv)
    assert a.w % a.ts == 0 and a.h % a.ts == 0
    root = tempfile.mkdtemp(prefix="sbmc_tiles_")
    try:
        rng = np.random.default_rng(0)
        compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
        tile_io.write_scene(root, "scene", rng, a.ts, a.w // a.ts, a.h // a.ts, a.spp,
                            quantize=a.quantize, compress=compress)
        files = sorted(os.listdir(os.path.join(root, "scene")))
        file_bytes = sum(os.path.getsize(os.path.join(root, "scene", f)) for f in files)
        samples = a.spp * a.h * a.w
        inflated = a.h * a.w * (30 * 4 + a.spp * (63 * 4 + 12))
        out_bytes = a.h * a.w * 4 * (33 + a.spp * 96 + 3)
        line = {"metric": "Msamples/s (spp*H*W) tile reader: .bin+lz4 -> model input tensors",
                "unit": "Msamples/s",
                "config": {"workload": "FullImagesDataset item, %dx%d, %d spp, %d tiles of %d px"
                           % (a.w, a.h, a.spp, len(files), a.ts), "file_bytes": file_bytes,
                           "inflated_bytes": inflated, "tensor_bytes": out_bytes}}

        if th.cuda.is_available():
            d = datasets.FullImagesDataset(root)
            for _ in range(2):
                item = d[0]
            th.cuda.synchronize()
            _lib.timing_collect()
            _lib.timing_enable(True)
            l0 = _lib.launch_count()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                item = d[0]
            th.cuda.synchronize()
            wall = (time.perf_counter() - t0) / a.steps
            _lib.timing_enable(False)
            kern = _lib.timing_collect().get("tiles", (0.0, 0))
            del item
            line["gpu"] = {"value": samples / wall / 1e6, "ms_per_item": wall * 1e3,
                           "kernel_ms_per_item": kern[0] / a.steps,
                           "gpu_launches": (_lib.launch_count() - l0) / a.steps,
                           "h2d_bytes_per_item": file_bytes,
                           "algorithmic_bytes_per_item": file_bytes + 2 * inflated + out_bytes}
            if kern[0] > 0:
                line["gpu"]["kernels_GBps"] = (
                    (file_bytes + 2 * inflated + out_bytes) / (kern[0] / a.steps * 1e-3) / 1e9)

        if cpu_baseline is not None:
            n = min(a.cpu_tiles, len(files))
            bufs = [open(os.path.join(root, "scene", f), "rb").read() for f in files[:n]]
            cpu = cpu_baseline(bufs) * len(files)
            line["cpu_baseline"] = {"value": samples / cpu / 1e6, "unit": "Msamples/s", "cores": 1,
                                    "kind": "port",
                                    "sample": "%d of %d tiles, scaled" % (n, len(files))}
        print(json.dumps(line))
    finally:
        shutil.rmtree(root)


if __name__ == "__main__":
    main()
