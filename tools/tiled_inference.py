#!/usr/bin/env python
"""BASELINE.json config 5: tiled multi-GPU inference of Multisteps (torchrun, one
rank per GPU).  --check compares against the unsharded forward on rank 0 (small
image); otherwise times spp=8, 3840x2160 (weak scaling is not meaningful here:
the image is fixed, so this is STRONG scaling of one 4K frame)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
import torch.distributed as dist
from sbmc_b200 import models, sharding


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--h", type=int, default=2160)
    ap.add_argument("--w", type=int, default=3840)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--fast", action="store_true", help="bf16 tcgen05 chains + bf16 U-net")
    ap.add_argument("--halo", action="store_true",
                    help="exchange U-net halos per step instead of recomputing a 144-row overlap")
    ap.add_argument("--unet-pad", type=int, default=64)
    ap.add_argument("--data", help="root folder of scene folders with .bin tiles: every rank reads "
                    "only the tiles of its row band (halo mode) instead of synthetic inputs")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    th.manual_seed(0)
    th.backends.cudnn.allow_tf32 = False
    th.backends.cuda.matmul.allow_tf32 = False
    if a.check:
        h, w, spp = 96 * max(world, 2), 160, 2
        net = models.Multisteps(20, 3, ksize=21).to(dev).eval()
        nf = 20
        if a.halo:
            net = net.to(memory_format=th.channels_last)
    else:
        h, w, spp = a.h, a.w, a.spp
        net = models.Multisteps(93, 3).to(dev).eval()
        nf = 93
        if a.fast:
            net.bf16_chains = net.bf16_unet = True
            net = net.to(memory_format=th.channels_last)
    if world > 1:   # same weights everywhere
        for prm in net.parameters():
            dist.broadcast(prm.data, 0)
    # --check: host tensors (the band is uploaded inside the call); timing runs: the
    # full frame is resident on every GPU so that only model + gather are timed
    gdev = "cpu" if a.check else dev
    g = th.Generator(device=gdev).manual_seed(1)
    samples = None if a.data else {
        "radiance": th.rand(1, spp, 3, h, w, generator=g, device=gdev),
        "features": th.randn(1, spp, nf, h, w, generator=g, device=gdev),
        "global_features": th.randn(1, 3, 1, 1, generator=g, device=gdev)}
    row_kw = {}
    if a.data:      # files -> each rank's band, read / inflated / assembled on its own GPU
        from sbmc_b200 import datasets
        assert a.halo, "--data feeds the halo mode"
        dset = datasets.FullImagesDataset(a.data, spp=a.spp if not a.check else None)
        h, w, spp, nf = dset.tiles_dset.image_height, dset.tiles_dset.image_width, dset.spp, \
            dset.num_features
        net = models.Multisteps(nf, 3).to(dev).eval().to(memory_format=th.channels_last)
        net.bf16_chains = net.bf16_unet = True
        if world > 1:
            for prm in net.parameters():
                dist.broadcast(prm.data, 0)
        lo, hi = sharding.halo_mode_rows(h, world, net.ksize, rank)
        t0 = time.perf_counter()
        band = dset.read_rows(0, lo, hi)
        th.cuda.synchronize()
        print("rank %d read rows [%d, %d) in %.1f ms" % (rank, lo, hi, 1e3 * (time.perf_counter() - t0)),
              flush=True)
        samples = {k: band[k].unsqueeze(0) for k in ("radiance", "features", "global_features")}
        row_kw = dict(image_height=h, row0=lo)

    def run():
        if a.halo:
            return sharding.multisteps_forward_halo(net, samples, rank, world,
                                                    unet_pad=a.unet_pad, **row_kw)["radiance"]
        return sharding.multisteps_forward_sharded(net, samples, rank, world)["radiance"]

    with th.no_grad():
        if a.check and a.halo:      # the halo path is the bf16 pipeline: compare like with like
            net.bf16_chains = net.bf16_unet = True
        out = run()
        if a.check:
            if rank == 0:
                if a.data:
                    whole = dset[0]
                    samples = {k: whole[k].unsqueeze(0)
                               for k in ("radiance", "features", "global_features")}
                ref = net({k: v.to(dev) for k, v in samples.items()})["radiance"]
                err = ((out - ref).norm() / ref.norm()).item()
                mx = (out - ref).abs().max().item()
                tol = 2e-2 if a.halo else 1e-4   # bf16 pipeline: tile-dependent rounding
                print("TILED_CHECK halo=%s world=%d shape=%s rel=%.2e max=%.2e %s" % (
                    a.halo, world, tuple(out.shape), err, mx, "OK" if err < tol else "FAILED"),
                    flush=True)
        else:
            th.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                out = run()
            th.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = (time.perf_counter() - t0) / a.steps
            if rank == 0:
                print(json.dumps({"bench": "Multisteps tiled inference (config 5)", "n_gpus": world,
                                  "spp": spp, "H": h, "W": w, "fast": a.fast, "mode": "halo exchange" if a.halo
                                  else "overlap recompute", "s_per_frame": dt,
                                  "Msamples_per_s": spp * h * w / dt / 1e6,
                                  "note": "frame resident on each GPU; strong scaling of one frame"}),
                      flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
