"""The callers around the hot path: scripts/denoise.py
tiling, checkpoints, image writers on CPU; the two scripts end to end on the GPU."""
import importlib.util
import os
import struct
import sys
import zlib

import numpy as np
import pytest
import torch as th

from sbmc_b200 import _compat, imageio
from tests import tile_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_script(name):
    spec = importlib.util.spec_from_file_location("sbmc_b200_script_" + name,
                                                  os.path.join(ROOT, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Blur(th.nn.Module):
    """Stand-in denoiser: 5x5 box filter of the mean radiance, cropped by 2 like the
    real model crops by ksize // 2; adds the global feature so its plumbing shows."""

    def forward(self, batch):
        img = batch["radiance"].mean(1)
        k = th.ones(3, 1, 5, 5) / 25.0
        out = th.nn.functional.conv2d(img, k, groups=3)
        return {"radiance": out + batch["global_features"][:, :1]}


@pytest.mark.parametrize("h,w,tile,pad", [(40, 56, 24, 4), (30, 30, 64, 8), (50, 23, 20, 6),
                                          (64, 64, 32, 8)])
def test_tiled_denoising_equals_whole_image(h, w, tile, pad):
    denoise = load_script("denoise")
    g = th.Generator().manual_seed(h * w)
    batch = {"radiance": th.rand(1, 2, 3, h, w, generator=g),
             "features": th.rand(1, 2, 4, h, w, generator=g),
             "low_spp": th.rand(1, 3, h, w, generator=g),
             "global_features": th.rand(1, 3, 1, 1, generator=g)}
    model = _Blur()
    whole = denoise.denoise_batch(model, batch, False, 10 ** 6, pad)
    tiled = denoise.denoise_batch(model, batch, False, tile, pad)
    parts = denoise.split_tiles(batch, tile, pad)
    if h > tile or w > tile:
        assert len(parts) > 1
        covered = th.zeros(h, w)
        for part, y0, y1, x0, x1, crop in parts:
            assert "global_features" in part                  # the reference drops it
            assert part["features"].shape[-2] <= tile and part["features"].shape[-1] <= tile
            covered[y0:y1, x0:x1] += 1
        assert bool((covered == 1).all())                     # every pixel exactly once
    # away from the 2-pixel ring the blur cannot see, tiles and whole image agree
    assert th.allclose(tiled[..., 2:-2, 2:-2], whole[..., 2:-2, 2:-2], atol=1e-6)
    with pytest.raises(ValueError):          # tiles must advance
        denoise.split_tiles(batch, 8, 4)


def test_checkpointer_round_trip_and_latest(tmp_path):
    model = th.nn.Linear(3, 2)
    opt = th.optim.Adam(model.parameters(), lr=1e-3)
    model(th.ones(1, 3)).sum().backward()
    opt.step()
    meta = dict(model_params=dict(ksize=7, gather=False, pixel=False), kpcn_mode=False,
                data_params=dict(spp=4, mode="sbmc"))
    ck = _compat.Checkpointer(str(tmp_path), model, meta=meta, optimizers=opt)
    assert ck.load_latest() == (None, None)
    assert _compat.Checkpointer.load_meta(str(tmp_path)) is None
    first = ck.save("epoch_0", extras={"epoch": 0})
    os.utime(first, (1, 1))
    with th.no_grad():
        model.weight.add_(1.0)
    ck.save("epoch_1", extras={"epoch": 1})
    want = model.weight.clone()
    other = th.nn.Linear(3, 2)
    opt2 = th.optim.Adam(other.parameters(), lr=1e-3)
    extras, got_meta = _compat.Checkpointer(str(tmp_path), other, optimizers=opt2).load_latest()
    assert extras == {"epoch": 1} and got_meta == meta
    assert th.equal(other.weight, want)
    assert opt2.state_dict()["state"][0]["step"] == opt.state_dict()["state"][0]["step"]
    assert _compat.Checkpointer.load_meta(str(tmp_path)) == meta
    # a corrupt newest file falls back to the previous one
    (tmp_path / "zz_broken.pth").write_bytes(b"not a checkpoint")
    extras, _ = _compat.Checkpointer(str(tmp_path), other).load_latest()
    assert extras == {"epoch": 1}


def test_trainer_loop_and_callbacks(tmp_path):
    calls = []

    class Interface(object):
        def forward(self, batch):
            return {"y": batch * 2}

        def backward(self, batch, fwd):
            calls.append(("bwd", int(batch.sum())))
            return {"loss": 1.0, "rmse": 2.0}

        def init_validation(self):
            return {"loss": 0.0, "rmse": 0.0, "n": 0}

        def update_validation(self, batch, fwd, running):
            running["n"] += 1
            return running

    class Spy(object):
        def validation_end(self, val):
            calls.append(("val", val["n"]))

    model = th.nn.Linear(1, 1)
    trainer = _compat.Trainer(Interface())
    trainer.add_callback(_compat.LoggingCallback(["loss", "rmse"], frequency=1))
    trainer.add_callback(_compat.CheckpointingCallback(_compat.Checkpointer(str(tmp_path), model)))
    trainer.add_callback(Spy())
    data = [th.tensor([1]), th.tensor([2]), th.tensor([3])]
    assert trainer.train(data, num_epochs=2, val_dataloader=data[:2]) == 6
    assert calls.count(("val", 2)) == 2 and len([c for c in calls if c[0] == "bwd"]) == 6
    assert sorted(os.listdir(tmp_path)) == ["epoch_0.pth", "epoch_1.pth", "training_end.pth"]
    assert _compat.Trainer(Interface()).train(data, max_steps=2) == 2


def test_exr_and_png_writers(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.standard_normal((7, 5, 3)).astype(np.float32) * 100
    img[0, 0] = [np.inf, -0.0, 1e-30]
    path = str(tmp_path / "out.exr")
    imageio.write_exr(path, img)
    back = imageio.read_exr(path)
    assert back.dtype == np.float32 and np.array_equal(back.view(np.int32), img.view(np.int32))
    raw = open(path, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == 20000630 and b"channels\0chlist\0" in raw
    gray = rng.random((4, 6)).astype(np.float32)
    imageio.write_exr(path, gray)
    assert np.array_equal(imageio.read_exr(path)[..., 0], gray)
    with pytest.raises(ValueError):
        imageio.write_exr(path, np.zeros((2, 2, 2)))

    png = str(tmp_path / "out.png")
    rgb = rng.integers(0, 256, (9, 4, 3), dtype=np.uint8)
    imageio.write_png(png, rgb)
    buf = open(png, "rb").read()
    assert buf[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, dims = 8, b"", None
    while pos < len(buf):
        (n,) = struct.unpack_from(">I", buf, pos)
        tag, body = buf[pos + 4:pos + 8], buf[pos + 8:pos + 8 + n]
        assert struct.unpack_from(">I", buf, pos + 8 + n)[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
        if tag == b"IHDR":
            dims = struct.unpack(">IIBBBBB", body)
        if tag == b"IDAT":
            idat += body
        pos += 12 + n
    assert dims == (4, 9, 8, 2, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(9, 1 + 12)
    assert not rows[:, 0].any() and np.array_equal(rows[:, 1:].reshape(9, 4, 3), rgb)


def test_display_callback_gallery_follows_the_reference(tmp_path):
    from sbmc_b200 import callbacks
    g = th.Generator().manual_seed(0)
    batch = {"low_spp": 3 * th.rand(2, 3, 10, 12, generator=g) - 0.5,
             "target_image": 3 * th.rand(2, 3, 10, 12, generator=g),
             "spp": th.full((2, 1, 1, 1), 4, dtype=th.int32)}
    fwd = {"radiance": 3 * th.rand(2, 3, 8, 10, generator=g)}
    cb = callbacks.DenoisingDisplayCallback(frequency=2, out_dir=str(tmp_path))
    img = cb.visualized_image(batch, fwd)
    assert img.shape == (2, 3, 32, 10) and img.min() >= 0 and img.max() <= 1
    out, tgt = fwd["radiance"], batch["target_image"][..., 1:-1, 1:-1]
    want = th.cat([batch["low_spp"][..., 1:-1, 1:-1], out, tgt, (out - tgt).abs()], -2).clamp(min=0)
    want = (want / (1 + want)).pow(1 / 2.2).clamp(0, 1)      # sbmc/callbacks.py:49-57
    assert th.allclose(img, want)
    assert cb.caption(batch, fwd) == "vertically: 4spp, ours, target, difference"
    assert cb.batch_end(batch, fwd, {}) is None
    path = cb.batch_end(batch, fwd, {})
    assert path.endswith("images_000002.png") and os.path.getsize(path) > 100


def test_script_parsers_keep_the_reference_options():
    d = load_script("denoise").parser().parse_args(
        ["--input", "i", "--checkpoint", "c", "--output", "o.exr", "--spp", "4"])
    assert (d.tile_size, d.tile_pad, d.spp) == (1024, 256, 4)
    t = load_script("train").parser().parse_args(
        ["--data", "d", "--checkpoint_dir", "c", "--constant_spp", "--dont_use_bt", "--gather",
         "--num_worker_threads", "4", "--cuda", "--env", "sbmc", "--port", "8097", "--debug"])
    assert (t.spp, t.ksize, t.randomize_spp, t.load_bt, t.load_p, t.gather, t.lr) == (
        8, 21, False, False, True, True, 1e-4)


# ------------------------------------------------------------------ end to end, two backends
def _scene(tmp_path, ts=16, nx=3, ny=3, spp=2):
    rng = np.random.default_rng(5)
    compress = tile_io.compress_frame if tile_io.liblz4() else tile_io.stored_frame
    tile_io.write_scene(str(tmp_path / "data"), "scene", rng, ts, nx, ny, spp, quantize=1.0 / 32,
                        compress=compress)
    return str(tmp_path / "data")


@pytest.fixture(params=["host-emulation", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    """"gpu": the scripts as shipped.  "host-emulation": the same Python with the
    tile kernels built for the host (tests/test_tiles.py), the optimizer kernels
    likewise (tests/test_optim.py) and the two custom ops stood in for by the
    oracle, on CPU tensors -- checks the scripts' own logic without a GPU."""
    if request.param == "host-emulation":
        import sbmc_b200.functions as funcs
        from sbmc_b200 import datasets, optim
        from tests import kats
        from tests.test_optim import EmulBackend as OptimEmul
        from tests.test_tiles import EmulBackend as TilesEmul
        KW, S2G = kats.oracle_functions()
        monkeypatch.setattr(funcs, "KernelWeighting", KW)
        monkeypatch.setattr(funcs, "Scatter2Gather", S2G)
        monkeypatch.setattr(datasets, "_backend", TilesEmul)
        monkeypatch.setattr(optim, "_backend", OptimEmul)
    return request.param


@pytest.mark.parametrize("fused_optimizer", [False, True], ids=["adam", "fused-adam"])
def test_train_then_denoise_scripts(tmp_path, backend, fused_optimizer, monkeypatch):
    root = _scene(tmp_path)
    train = load_script("train")
    denoise = load_script("denoise")
    if backend == "host-emulation":
        monkeypatch.setattr(train, "_device", lambda: "cpu")
        monkeypatch.setattr(denoise, "_device", lambda: "cpu")
    ckpt = str(tmp_path / "ckpt")
    args = train.parser().parse_args(
        ["--data", root, "--checkpoint_dir", ckpt, "--constant_spp", "--spp", "2", "--bs", "3",
         "--ksize", "3", "--num_epochs", "1", "--max_steps", "2", "--log_every", "1",
         "--display_every", "2"]
        + (["--fused_optimizer"] if fused_optimizer else []))
    train.main(args)
    assert "training_end.pth" in os.listdir(ckpt)
    assert os.listdir(os.path.join(ckpt, "display")) == ["images_000002.png"]
    meta = _compat.Checkpointer.load_meta(ckpt)
    assert meta["model_params"]["ksize"] == 3 and meta["data_params"]["spp"] == 2

    outs = {}
    for name, extra in (("whole", []), ("tiled", ["--tile_size", "32", "--tile_pad", "8"])):
        out = str(tmp_path / ("%s.exr" % name))
        denoise.main(denoise.parser().parse_args(
            ["--input", os.path.join(root, "scene"), "--checkpoint", ckpt, "--output", out] + extra))
        outs[name] = imageio.read_exr(out)
        assert outs[name].shape == (48, 48, 3) and np.isfinite(outs[name]).all()
        assert os.path.exists(out.replace(".exr", ".png"))
    # kernel 3 crops one pixel: the ring is zero padding, the inside is denoised
    assert not outs["whole"][0].any() and outs["whole"][1:-1, 1:-1].any()
    # tiles see 8 pixels of context where the U-nets would want ~120: the two runs
    # agree only roughly -- what is checked is the stitching (no seams of zeros)
    assert np.count_nonzero(outs["tiled"][1:-1, 1:-1]) > 0.99 * 46 * 46 * 3


def test_kpcn_mode_scripts(tmp_path, backend):
    """The [Bako2017] comparison path of both scripts (train.py --kpcn_mode,
    denoise.py reading kpcn_mode from the checkpoint's meta; Makefile:167-173,112-116
    of the reference)."""
    root = _scene(tmp_path, ts=48, nx=1, ny=1, spp=2)
    train = load_script("train")
    denoise = load_script("denoise")
    if backend == "host-emulation":
        import unittest.mock as mock
        patches = [mock.patch.object(train, "_device", lambda: "cpu"),
                   mock.patch.object(denoise, "_device", lambda: "cpu")]
    else:
        patches = []
    for p_ in patches:
        p_.start()
    try:
        ckpt = str(tmp_path / "ckpt")
        train.main(train.parser().parse_args(
            ["--data", root, "--checkpoint_dir", ckpt, "--constant_spp", "--spp", "2", "--bs", "1",
             "--kpcn_mode", "--ksize", "3", "--num_epochs", "1", "--max_steps", "1"]))
        meta = _compat.Checkpointer.load_meta(ckpt)
        assert meta["kpcn_mode"] is True and meta["data_params"]["mode"] == "kpcn"
        out = str(tmp_path / "kpcn.exr")
        denoise.main(denoise.parser().parse_args(
            ["--input", os.path.join(root, "scene"), "--checkpoint", ckpt, "--output", out]))
        img = imageio.read_exr(out)
        # nine unpadded 5x5 convolutions crop 18 pixels per side (the gather kernel none)
        assert img.shape == (48, 48, 3) and np.isfinite(img).all()
        assert not img[:18].any() and not img[:, :18].any() and img[18:-18, 18:-18].all()
    finally:
        for p_ in patches:
            p_.stop()


@pytest.mark.gpu
def test_train_script_mixed_precision_pipeline_under_a_cuda_graph(tmp_path):
    """scripts/train.py --bf16_train --cuda_graph --device_prefetch 2: the tcgen05 training
    pipeline (sbmc_b200/train_pipeline.py) replayed from one CUDA graph per batch shape while
    the loader's worker thread decodes the next groups of batches on its side stream, from
    the .bin tiles to the checkpoint; the denoise script then loads what it wrote."""
    import torch as th
    root = _scene(tmp_path)
    train = load_script("train")
    denoise = load_script("denoise")
    ckpt = str(tmp_path / "ckpt")
    train.main(train.parser().parse_args(
        ["--data", root, "--checkpoint_dir", ckpt, "--constant_spp", "--spp", "2", "--bs", "3",
         "--ksize", "3", "--num_epochs", "2", "--max_steps", "5", "--log_every", "1",
         "--bf16_train", "--cuda_graph", "--device_prefetch", "2"]))
    state = th.load(os.path.join(ckpt, "training_end.pth"), map_location="cpu", weights_only=False)
    tensors = [v for v in state["model"].values() if isinstance(v, th.Tensor)]
    assert tensors and all(th.isfinite(v).all() for v in tensors)
    out = str(tmp_path / "out.exr")
    denoise.main(denoise.parser().parse_args(
        ["--input", os.path.join(root, "scene"), "--checkpoint", ckpt, "--output", out]))
    img = imageio.read_exr(out)
    assert img.shape == (48, 48, 3) and np.isfinite(img).all()
