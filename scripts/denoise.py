#!/usr/bin/env python
"""Denoise a folder of sample tiles with a trained model: the caller of the hot
path at inference time.  Same command line and flow as the reference's
scripts/denoise.py (:96-194): FullImagesDataset -> Multisteps / KPCN in eval mode
-> overlapping tiles when the image exceeds --tile_size -> .exr + .png.

Differences, all on purpose:
  * the tile loop forwards `global_features` to every tile and emits each tile
    once (the reference's `_split_tiles`, denoise.py:54-93, drops the global
    features and appends the same tile once per key, so its tiled path cannot run);
  * `model_params` stored in the checkpoint's meta (ksize / gather / pixel,
    train.py:62-66,84) are honoured instead of always building the default model;
  * `--fast` runs the per-sample chains on the tcgen05 kernels and the U-nets in
    bf16 channels-last (sbmc_b200.models.Multisteps.bf16_chains / bf16_unet);
  * samples are inflated and assembled on the GPU (sbmc_b200.datasets).
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch as th
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sbmc_b200 import _compat, datasets, imageio, models  # noqa: E402

LOG = _compat.get_logger(__name__)

TILED_KEYS = ("radiance", "features", "kpcn_diffuse_in", "kpcn_specular_in",
              "kpcn_diffuse_buffer", "kpcn_specular_buffer", "kpcn_albedo")
WHOLE_KEYS = ("global_features",)


def pad_to_input(part, out, kpcn_mode):
    """Zero-pads the network output back to its input's size (denoise.py:42-51)."""
    ref = part["kpcn_diffuse_in" if kpcn_mode else "features"]
    pad_h = (ref.shape[-2] - out.shape[-2]) // 2
    pad_w = (ref.shape[-1] - out.shape[-1]) // 2
    return th.nn.functional.pad(out, (pad_w, pad_w, pad_h, pad_h))


def split_tiles(batch, max_sz=1024, pad=256):
    """[(tile batch, y0, y1, x0, x1, (crop top, bottom, left, right))]: tiles of at
    most max_sz pixels stepping by max_sz - 2 pad; (y0, y1, x0, x1) is the region of
    the output each tile is responsible for."""
    h, w = batch["low_spp"].shape[-2:]
    if h <= max_sz and w <= max_sz:
        return [(batch, 0, h, 0, w, (0, 0, 0, 0))]
    step = max_sz - 2 * pad
    if step <= 0:
        raise ValueError("tile_size must exceed twice tile_pad")
    tiles = []
    for start_y in range(0, h, step):
        end_y = min(start_y + max_sz, h)
        pad_y = 0 if start_y == 0 else pad
        pad_y2 = 0 if end_y == h else pad
        for start_x in range(0, w, step):
            end_x = min(start_x + max_sz, w)
            pad_x = 0 if start_x == 0 else pad
            pad_x2 = 0 if end_x == w else pad
            part = {k: batch[k] for k in WHOLE_KEYS if k in batch}
            for k in TILED_KEYS:
                if k in batch:
                    part[k] = batch[k][..., start_y:end_y, start_x:end_x]
            tiles.append((part, start_y + pad_y, end_y - pad_y2, start_x + pad_x,
                          end_x - pad_x2, (pad_y, pad_y2, pad_x, pad_x2)))
            if end_x == w:
                break
        if end_y == h:
            break
    return tiles


def denoise_batch(model, batch, kpcn_mode, tile_size, tile_pad):
    out_radiance = th.zeros_like(batch["low_spp"])
    for part, y0, y1, x0, x1, crop in split_tiles(batch, tile_size, tile_pad):
        with th.no_grad():
            out = pad_to_input(part, model(part)["radiance"], kpcn_mode)
        out = out[..., crop[0]:out.shape[-2] - crop[1], crop[2]:out.shape[-1] - crop[3]]
        out_radiance[..., y0:y1, x0:x1] = out
    return out_radiance


def build_model(data, meta, fast=False):
    params = dict((meta or {}).get("model_params") or {})
    if (meta or {}).get("kpcn_mode"):
        LOG.info("Using [Bako2017] denoiser.")
        return models.KPCN(data.num_features, ksize=params.get("ksize", 21))
    model = models.Multisteps(data.num_features, data.num_global_features,
                              ksize=params.get("ksize", 21),
                              splat=not params.get("gather", False),
                              pixel=params.get("pixel", False))
    if fast:
        model.bf16_chains = True
        model.bf16_unet = True
    return model


def _device():
    if not th.cuda.is_available():
        raise RuntimeError("sbmc_b200 runs on a CUDA device only (no CPU path)")
    return "cuda"


def _sync(device):
    if device == "cuda":
        th.cuda.synchronize()


def main(args):
    start = time.time()
    if not os.path.exists(args.input):
        raise ValueError("input {} does not exist".format(args.input))
    device = _device()
    # the dataset wants a root of scene folders; `input` is one scene (denoise.py:101-104)
    data_root = os.path.abspath(args.input)
    tmpdir = tempfile.mkdtemp()
    os.symlink(data_root, os.path.join(tmpdir, os.path.basename(data_root)))
    LOG.info("Loading model %s", args.checkpoint)
    meta = _compat.Checkpointer.load_meta(args.checkpoint) or {}
    data_params = dict(meta.get("data_params") or {})
    if args.spp:
        data_params["spp"] = args.spp
    data = datasets.FullImagesDataset(tmpdir, **data_params)
    LOG.info("Denoising input with %s spp", data.spp)

    kpcn_mode = bool(meta.get("kpcn_mode"))
    model = build_model(data, meta, fast=args.fast)
    model.train(False)
    model.to(device)
    _, loaded = _compat.Checkpointer(args.checkpoint, model, None).load_latest()
    LOG.info("Loading latest checkpoint %s", "failed" if loaded is None else "success")
    LOG.info("setup time %.1f ms", (time.time() - start) * 1000)

    loader = DataLoader(data, batch_size=1, shuffle=False, num_workers=0)
    for batch in loader:
        _sync(device)
        start = time.time()
        out = denoise_batch(model, batch, kpcn_mode, int(args.tile_size), int(args.tile_pad))
        _sync(device)
        LOG.info("    denoising time %.1f ms", (time.time() - start) * 1000)
        image = out[0].float().cpu().numpy().transpose([1, 2, 0])
        outdir = os.path.dirname(os.path.abspath(args.output))
        os.makedirs(outdir, exist_ok=True)
        imageio.write_exr(args.output, image)
        png = os.path.splitext(args.output)[0] + ".png"
        imageio.write_png(png, (np.clip(image, 0, 1) * 255).astype(np.uint8))
    shutil.rmtree(tmpdir)


def parser():
    p = argparse.ArgumentParser()
    p.add_argument("--input", type=str, required=True,
                   help="folder containing the sample .bin files.")
    p.add_argument("--checkpoint", type=str, required=True,
                   help="folder containing the model checkpoint.")
    p.add_argument("--output", type=str, required=True, help="output destination.")
    p.add_argument("--spp", type=int, help="number of samples to use as input.")
    p.add_argument("--tile_size", default=1024, help="We process in tiles to limit GPU "
                   "memory usage. This is the tile size.")
    p.add_argument("--tile_pad", default=256, help="We process in tiles to limit GPU memory "
                   "usage. This is the padding around tiles, for overlapping tiles.")
    p.add_argument("--fast", action="store_true", help="bf16 tensor-core inference pipeline.")
    return p


if __name__ == "__main__":
    main(parser().parse_args())
