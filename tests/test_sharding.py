"""CPU suite: the multi-GPU row-band plan and its halo exchange / halo reduce,
driven with world_size 2 and 3 over gloo on CPU tensors.  The band compute is
stood in for by the oracle on the extended band (weights / d_output zero in the
halo rows), so the test checks exactly what the sharding layer is responsible
for: that exchanged halos reproduce the unsharded result."""
import os
import socket

import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from sbmc_b200 import sharding


def test_band_plan():
    plan = sharding.BandPlan(2160, 8, 21)
    assert plan.pad == 10
    assert [plan.rows(r) for r in range(8)] == [270] * 8
    assert plan.halo_top(0) == 0 and plan.halo_bot(7) == 0
    assert plan.halo_top(3) == 10 and plan.halo_bot(3) == 10
    plan = sharding.BandPlan(103, 4, 4)            # even kernel: c0=1, K-1-c0=2
    assert plan.pad == 2
    assert [plan.rows(r) for r in range(4)] == [26, 26, 26, 25]
    assert plan.y0 == [0, 26, 52, 78] and plan.y1[-1] == 103
    uplan = sharding.BandPlan(2160, 8, 21, pad=64)   # U-net halo of the tiled model
    assert uplan.pad == 64 and uplan.halo_top(0) == 0 and uplan.halo_bot(0) == 64
    assert uplan.halo_top(7) == 64 and uplan.halo_bot(7) == 0
    aplan = sharding.BandPlan(2160, 8, 21, align=4)  # 270 rows / rank is not a multiple of 4
    assert [aplan.rows(r) for r in range(8)] == [272] * 4 + [268] * 4
    assert all(y % 4 == 0 for y in aplan.y0) and aplan.y1[-1] == 2160
    assert sharding.BandPlan(103, 4, 4, align=4).y1 == [28, 56, 80, 103]
    with pytest.raises(ValueError):
        sharding.BandPlan(3, 4, 3)
    with pytest.raises(ValueError):
        sharding.BandPlan(16, 4, 21)               # bands shorter than the halo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, q):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, c, h, w, kh, kw = shape
        g = th.Generator().manual_seed(0)           # same tensors on every rank
        data = 2 * th.randn(n, c, h, w, generator=g)
        weights = th.randn(n, kh, kw, h, w, generator=g)
        d_output = th.randn(n, c, h, w, generator=g)
        d_sum_w = th.randn(n, h, w, generator=g)
        plan = sharding.BandPlan(h, world, kh, kw)
        top, bot, rows = plan.halo_top(rank), plan.halo_bot(rank), plan.rows(rank)

        band = plan.band(rank, data, 2).contiguous()
        ext = sharding.exchange_halo(plan, rank, band)
        want = data[:, :, plan.y0[rank] - top:plan.y1[rank] + bot]
        ok_ext = th.equal(ext, want)

        def pad_rows(t, dim):                        # zeros in the halo rows
            shp = list(t.shape)
            shp[dim] = top + rows + bot
            z = t.new_zeros(shp)
            z.narrow(dim, top, rows).copy_(t)
            return z

        w_ext = pad_rows(plan.band(rank, weights, 3), 3)
        do_ext = pad_rows(plan.band(rank, d_output, 2), 2)
        ds_ext = pad_rows(plan.band(rank, d_sum_w, 1), 1)
        out_ext, sw_ext = oracle.kernel_weighting(ext, w_ext)
        dd_ext, dw_ext = oracle.kernel_weighting_grad(ext, w_ext, do_ext, ds_ext)
        d_data = sharding.reduce_halo(plan, rank, dd_ext)

        # the pre-allocated / side-stream variant of the same two exchanges
        pipe = sharding.HaloPipeline(plan, rank, band.shape, "cpu")
        pext = pipe.new_ext().fill_(float("nan"))
        pipe.band(pext).copy_(band)
        pipe.wait(pipe.exchange_async(pext))
        dd2 = dd_ext.clone()
        pipe.wait(pipe.reduce_async(dd2))
        ok_pipe = th.equal(pext, want) and th.equal(pipe.band(dd2), d_data)

        out = sharding.gather_bands(plan, rank, out_ext[:, :, top:top + rows], 2)
        sum_w = sharding.gather_bands(plan, rank, sw_ext[:, top:top + rows], 1)
        d_data = sharding.gather_bands(plan, rank, d_data, 2)
        d_weights = sharding.gather_bands(plan, rank, dw_ext[:, :, :, top:top + rows], 3)

        ro, rs = oracle.kernel_weighting(data, weights)
        rdd, rdw = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
        res = {
            "ext": ok_ext, "pipeline": ok_pipe,
            # forward / d_weights use the same per-pixel arithmetic: bit-identical
            "output": th.equal(out, ro), "sum_w": th.equal(sum_w, rs),
            "d_weights": th.equal(d_weights, rdw),
            # d_data sums the seam rows in a different order
            "d_data": (d_data - rdd).abs().max().item() <= 1e-5 * rdd.abs().max().item(),
        }
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [
    (2, (2, 3, 24, 20, 5, 5)),
    (2, (1, 3, 45, 16, 21, 21)),
    (3, (1, 2, 31, 12, 4, 6)),
])
def test_halo_exchange_gloo(world, shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        assert all(res.values()), (rank, res)


def _model_worker(rank, world, port, q):
    """Tiled model inference (BASELINE config 5 logic) over gloo on CPU: the custom
    ops are stood in for by the oracle, the convs are torch's CPU kernels."""
    from sbmc_b200 import functions as funcs
    from sbmc_b200 import models
    from tests import kats
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        KW, S2G = kats.oracle_functions()
        funcs.KernelWeighting, funcs.Scatter2Gather = KW, S2G
        th.manual_seed(0)
        net = models.Multisteps(5, 2, width=8, embedding_width=8, ksize=5, nsteps=1).eval()
        g = th.Generator().manual_seed(1)
        h, w, spp = 48 * world, 24, 2
        samples = {"radiance": th.rand(1, spp, 3, h, w, generator=g),
                   "features": th.randn(1, spp, 5, h, w, generator=g),
                   "global_features": th.randn(1, 2, 1, 1, generator=g)}
        with th.no_grad():
            out = sharding.multisteps_forward_sharded(net, samples, rank, world, overlap=44)
            ref = net(samples)
        err = (out["radiance"] - ref["radiance"]).abs().max().item()
        q.put((rank, {"shape": out["radiance"].shape == ref["radiance"].shape,
                      "close": err < 1e-5}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tiled_model_inference_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_model_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        assert all(res.values()), (rank, res)


def _reader_worker(rank, world, port, q):
    """Row-sharded reading of a scene (tile files -> each rank's band, no exchange)
    followed by the final gather: every rank ends with the whole image.  The two
    launches of the reader run as the device code built for the host."""
    from sbmc_b200 import datasets
    from tests.test_tiles import DATA, EmulBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        datasets._backend = EmulBackend
        d = datasets.FullImagesDataset(DATA, spp=2)
        plan = sharding.BandPlan(d.tiles_dset.image_height, world, 3)
        res = {}
        lo, hi = plan.y0[rank], plan.y1[rank]
        band = d.read_rows(1, lo, hi)
        whole = d[1]
        for key, dim in (("features", 2), ("radiance", 2), ("low_spp", 1), ("target_image", 1)):
            full = sharding.gather_bands(plan, rank, band[key], dim)
            res[key] = th.equal(full, whole[key])
        # with the K x K halo the band is what the sharded model consumes
        top, bot = plan.halo_top(rank), plan.halo_bot(rank)
        ext = d.read_rows(1, lo - top, hi + bot)["features"]
        res["halo"] = th.equal(ext, whole["features"][..., lo - top:hi + bot, :])
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_reader_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reader_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        assert all(res.values()), (rank, res)
