#!/usr/bin/env python
"""Train a denoiser: the caller of the hot path at training time.  Same options
and flow as the reference's scripts/train.py (:32-152): TilesDataset /
MultiSampleCountDataset -> Multisteps / KPCN -> SampleBasedDenoiserInterface
(Adam, tonemapped relative MSE, gradient clipping) -> checkpoints whose meta holds
model_params / kpcn_mode / data_params for scripts/denoise.py.

The `ttools` trainer, argument parser and callbacks are external to the reference
tree and absent here; sbmc_b200._compat carries small equivalents (no Visdom
display).  Samples are inflated and assembled on the GPU, so the DataLoader runs
without worker processes.
"""
import argparse
import os
import sys

import numpy as np
import torch as th
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sbmc_b200 import _compat, callbacks, datasets, interfaces, models  # noqa: E402

LOG = _compat.get_logger(__name__)


def _device():
    if not th.cuda.is_available():
        raise RuntimeError("sbmc_b200 runs on a CUDA device only (no CPU path)")
    return "cuda"


def _init_distributed():
    """(rank, world) under torchrun (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* in the
    environment): one process per GPU over NCCL."""
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    th.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world)
    return rank, world


def main(args):
    np.random.seed(0)
    th.manual_seed(0)
    device = _device()
    rank, world = _init_distributed() if args.distributed else (0, 1)
    data_args = dict(spp=args.spp,
                     mode=datasets.TilesDataset.KPCN_MODE if args.kpcn_mode
                     else datasets.TilesDataset.SBMC_MODE,
                     load_coords=args.load_coords, load_gbuffer=args.load_gbuffer,
                     load_p=args.load_p, load_ld=args.load_ld, load_bt=args.load_bt)
    if args.randomize_spp:
        if args.bs != 1:
            LOG.error("Training with randomized spp is only valid for batch_size=1, got %d",
                      args.bs)
            raise RuntimeError("Incorrect batch size")
        data = datasets.MultiSampleCountDataset(args.data, **data_args)
        LOG.info("Training with randomized sample count in [%d, %d]", 2, args.spp)
    else:
        data = datasets.TilesDataset(args.data, **data_args)
        LOG.info("Training with a single sample count: %dspp", args.spp)

    if args.kpcn_mode:
        model = models.KPCN(data.num_features, ksize=args.ksize)
        model_params = dict(ksize=args.ksize)
    else:
        model = models.Multisteps(data.num_features, data.num_global_features, ksize=args.ksize,
                                  splat=not args.gather, pixel=args.pixel)
        model_params = dict(ksize=args.ksize, gather=args.gather, pixel=args.pixel)

    # the next batch's tile files are read while the current step runs on the GPU
    loader = datasets.PrefetchLoader(data, batch_size=args.bs, shuffle=True)
    val_loader = None
    if args.val_data is not None:
        val = datasets.TilesDataset(args.val_data, **data_args)
        val_loader = DataLoader(val, batch_size=args.bs, num_workers=0, shuffle=False)

    meta = dict(model_params=model_params, kpcn_mode=args.kpcn_mode, data_params=data_args)
    LOG.info("Model configuration: %s", model_params)
    if args.bf16_train:
        if not hasattr(model, "bf16_train"):
            raise SystemExit("--bf16_train serves the sample-based model (not --kpcn_mode)")
        model.bf16_train = True        # mixed-precision pipeline on the repo's tcgen05 kernels
    interface = interfaces.SampleBasedDenoiserInterface(
        model, lr=args.lr, cuda=device == "cuda",
        fused_optimizer=args.fused_optimizer or args.cuda_graph, cuda_graph=args.cuda_graph,
        distributed=world > 1)
    checkpointer = _compat.Checkpointer(args.checkpoint_dir, model, meta=meta,
                                        optimizers=interface.optimizer)
    checkpointer.load_latest()

    trainer = _compat.Trainer(interface)
    if rank == 0:                   # one rank logs and writes checkpoints (all hold the same weights)
        trainer.add_callback(_compat.LoggingCallback(["loss", "rmse"], frequency=args.log_every))
        trainer.add_callback(_compat.CheckpointingCallback(checkpointer))
    if args.display_every > 0 and rank == 0:      # the reference shows this gallery in Visdom (train.py:116-118)
        trainer.add_callback(callbacks.DenoisingDisplayCallback(
            frequency=args.display_every, out_dir=os.path.join(args.checkpoint_dir, "display")))
    LOG.info("Training started, 'Ctrl+C' to abort.")
    trainer.train(loader, num_epochs=args.num_epochs, val_dataloader=val_loader,
                  max_steps=args.max_steps)
    interface.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def parser():
    p = argparse.ArgumentParser()
    # ttools.BasicArgumentParser's options used by the reference script
    p.add_argument("--data", required=True, help="path to the training data.")
    p.add_argument("--val_data", help="path to the validation data.")
    p.add_argument("--checkpoint_dir", required=True, help="output directory.")
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--bs", type=int, default=1)
    p.add_argument("--num_epochs", type=int)
    p.add_argument("--max_steps", type=int, help="stop after this many steps (extra).")
    # accepted so that the reference's command lines keep working; without effect here:
    # samples are decoded on the GPU (no loader workers), training is CUDA-only, no Visdom
    p.add_argument("--num_worker_threads", type=int, default=0, help=argparse.SUPPRESS)
    p.add_argument("--cuda", action="store_true", default=True, help=argparse.SUPPRESS)
    p.add_argument("--env", default=None, help=argparse.SUPPRESS)
    p.add_argument("--port", type=int, default=None, help=argparse.SUPPRESS)
    p.add_argument("--debug", action="store_true", default=False,
                   help="verbose logging (ttools.set_logger(debug)).")
    p.add_argument("--log_every", type=int, default=50)
    p.add_argument("--display_every", type=int, default=0,
                   help="write a low-spp / output / target / difference gallery every N steps.")
    p.add_argument("--fused_optimizer", action="store_true",
                   help="clip + Adam over all tensors in three launches (extra).")
    p.add_argument("--bf16_train", action="store_true",
                   help="mixed-precision training pipeline: bf16 activations, fp32 accumulation "
                        "and master weights, every GEMM on the tcgen05 kernels (extra).")
    p.add_argument("--cuda_graph", action="store_true",
                   help="capture the training step in a CUDA graph per batch shape (extra; "
                        "implies --fused_optimizer; use with --constant_spp).")
    p.add_argument("--device_prefetch", type=int, default=0,
                   help="decode this many batches at a time on a side stream while the previous "
                        "ones train (extra; LZ4 frames are serial chains, one per warp: wide groups "
                        "keep the inflater off the critical path).")
    p.add_argument("--distributed", action="store_true",
                   help="data-parallel training under torchrun: one process per GPU, gradients "
                        "averaged with one NCCL all-reduce per step (extra).")
    p.add_argument("--spp", type=int, default=8, help="Max number of samples per pixel.")
    p.add_argument("--kpcn_mode", dest="kpcn_mode", action="store_true", default=False)
    p.add_argument("--gather", dest="gather", action="store_true", default=False)
    p.add_argument("--pixel", dest="pixel", action="store_true", default=False)
    p.add_argument("--ksize", type=int, default=21, help="Size of the kernels")
    p.add_argument("--constant_spp", dest="randomize_spp", action="store_false", default=True)
    p.add_argument("--dont_use_coords", dest="load_coords", action="store_false", default=True)
    p.add_argument("--dont_use_gbuffer", dest="load_gbuffer", action="store_false", default=True)
    p.add_argument("--dont_use_p", dest="load_p", action="store_false", default=True)
    p.add_argument("--dont_use_ld", dest="load_ld", action="store_false", default=True)
    p.add_argument("--dont_use_bt", dest="load_bt", action="store_false", default=True)
    return p


if __name__ == "__main__":
    _args = parser().parse_args()
    if _args.debug:
        import logging
        logging.getLogger().setLevel(logging.DEBUG)
    main(_args)
