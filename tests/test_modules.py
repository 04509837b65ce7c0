"""Module / model tests.  CPU suite: structure, error behaviour and state-dict
names (reference tests/test_modules.py:18-60 and SURVEY.md appendix A), and the
module logic with the two custom ops stood in for by the oracle.  GPU suite:
the same known-answer tests and a float64 torch restatement, on the real ops."""
import math

import pytest
import torch as th
import torch.nn.functional as F

from sbmc_b200 import functions as funcs
from sbmc_b200 import models, modules
from tests import kats


@pytest.fixture
def oracle_ops(monkeypatch):
    """Run the module logic on CPU: the custom ops are replaced by the oracle
    (test infrastructure; the product has no CPU path)."""
    KW, S2G = kats.oracle_functions()
    monkeypatch.setattr(funcs, "KernelWeighting", KW)
    monkeypatch.setattr(funcs, "Scatter2Gather", S2G)


def test_convchain_basic():
    """reference tests/test_modules.py:18-60."""
    with pytest.raises(ValueError):
        modules.ConvChain(3, 3, depth=0)
    with pytest.raises(ValueError):
        modules.ConvChain(3, 3, depth=-1)
    with pytest.raises(ValueError):
        modules.ConvChain(3, 3, output_type="randomstring")
    with pytest.raises(ValueError):
        modules.ConvChain(3, 3, activation="randomstring")
    with pytest.raises(ValueError):
        modules.ConvChain(3, 3, normalize=True, normalization_type="randomstring")
    for nrm in [False, True]:
        net = modules.ConvChain(3, 3, depth=3, width=32, normalize=nrm)
        idx = 1 if nrm else 0
        assert isinstance(net.layer_0, modules.ConvChain._ConvBNRelu)
        assert isinstance(net.layer_1, modules.ConvChain._ConvBNRelu)
        assert isinstance(net.prediction, th.nn.Conv2d)
        for layer, cin in ((net.layer_0, 3), (net.layer_1, 32)):
            l = list(layer.layer.children())
            assert isinstance(l[0], th.nn.Conv2d) and isinstance(l[1 + idx], th.nn.ReLU)
            assert l[0].kernel_size == (3, 3) and l[0].stride == (1, 1)
            assert l[0].in_channels == cin and l[0].out_channels == 32
            if nrm:
                assert isinstance(l[1], th.nn.BatchNorm2d)
        assert net.prediction.in_channels == 32 and net.prediction.out_channels == 3
        assert net.prediction.kernel_size == (3, 3)
        y = net(th.randn(1, 3, 8, 8))
        assert y.shape == (1, 3, 8, 8)
    assert modules.ConvChain(3, 5, ksize=5, pad=False, depth=2)(th.randn(1, 3, 12, 12)).shape \
        == (1, 5, 4, 4)


def test_multisteps_structure_and_state_dict_names():
    """layer shapes and parameter names of Multisteps(93, 3)."""
    with pytest.raises(ValueError):
        models.Multisteps(93, 3, ksize=4)
    with pytest.raises(ValueError):
        models.Multisteps(93, 3, ksize=1)
    with pytest.raises(ValueError):
        models.Multisteps(93, 3, nsteps=0)
    net = models.Multisteps(93, 3)
    sd = net.state_dict()
    shapes = {
        "embedding_00.layer_0.layer.0.weight_v": (128, 96, 1, 1),
        "embedding_00.layer_0.layer.0.weight_g": (128, 1, 1, 1),
        "embedding_00.layer_0.layer.0.bias": (128,),
        "embedding_00.prediction.weight_v": (128, 128, 1, 1),
        "embedding_01.layer_0.layer.0.weight_v": (128, 256, 1, 1),
        "embedding_02.layer_1.layer.0.weight_v": (128, 128, 1, 1),
        "propagation_00.net.left.layer_0.layer.0.weight_v": (128, 128, 3, 3),
        "propagation_00.net.next_level.left.layer_0.layer.0.weight_v": (256, 128, 3, 3),
        "propagation_00.net.next_level.next_level.left.layer_0.layer.0.weight_v": (512, 256, 3, 3),
        "propagation_00.net.next_level.next_level.left.prediction.weight_v": (512, 512, 3, 3),
        "propagation_00.net.next_level.right.layer_0.layer.0.weight_v": (256, 768, 3, 3),
        "propagation_02.net.right.layer_0.layer.0.weight_v": (128, 384, 3, 3),
        "propagation_02.net.right.prediction.weight_v": (128, 128, 3, 3),
        "kernel_regressor.layer_0.layer.0.weight_v": (128, 256, 1, 1),
        "kernel_regressor.prediction.weight_v": (441, 128, 1, 1),
        "kernel_regressor.prediction.bias": (441,),
    }
    for name, shape in shapes.items():
        assert tuple(sd[name].shape) == shape, name
    nparams = sum(p.numel() for p in net.parameters())
    assert 34.5e6 < nparams < 35.2e6          # ~34.8 M conv parameters
    assert isinstance(net.kernel_update, modules.ProgressiveKernelApply)
    assert net.kernel_update.splat
    assert isinstance(net.propagation_00.net.right.output_activation, th.nn.LeakyReLU)
    assert isinstance(net.kernel_regressor.layer_0.layer[1], th.nn.LeakyReLU)
    assert not hasattr(net.kernel_regressor, "output_activation")
    k = models.KPCN(27)
    assert tuple(k.state_dict()["diffuse.prediction.weight"].shape) == (441, 100, 5, 5)
    assert len([n for n in k.diffuse.state_dict() if n.endswith("weight")]) == 9


# -- KernelApply / ProgressiveKernelApply known answers (test_modules.py:63-140) --
def _kernel_apply_kat(device):
    bs, c, h, w, k = 4, 5, 16, 16, 3
    y, x, val = h // 2, w // 2, 1.43
    data = th.zeros(bs, c, h, w)
    data[0, 0, y, x] = val
    weights = th.zeros(bs, k * k, h, w)
    weights[0, :, y, x] = 1.0
    for splat in [True, False]:
        out, sum_w = modules.KernelApply(softmax=False, splat=splat)(
            data.to(device), weights.clone().to(device))
        out, sum_w = out.cpu(), sum_w.cpu()
        assert sum_w.shape == (bs, 1, h, w)
        assert abs(out[0, 0, y, x].item() - val) < 1e-4
        if splat:
            for dy in range(-(k // 2), k // 2 + 1):
                for dx in range(-(k // 2), k // 2 + 1):
                    assert abs(out[0, 0, y + dy, x + dx].item() - val) < 1e-4
                    assert abs(sum_w[0, 0, y + dy, x + dx].item() - 1) < 1e-4
        else:
            assert abs(sum_w[0, 0, y, x].item() - k * k) < 1e-4
    # softmax mode (KPCN): the weights of every pixel sum to one
    out, sum_w = modules.KernelApply(softmax=True, splat=False)(
        data.to(device), th.randn(bs, k * k, h, w).to(device))
    assert (sum_w.cpu() - 1).abs().max().item() < 1e-5


def _progressive_kat(device):
    bs, c, h, w, k = 4, 5, 16, 16, 3
    y, x, val = h // 2, w // 2, 1.43
    data = th.zeros(bs, c, h, w)
    data[0, 0, y, x] = val
    weights = th.zeros(bs, k * k, h, w)
    weights[0, :, y, x] = 1.0
    for splat in [True, False]:
        func = modules.ProgressiveKernelApply(splat=splat)
        out, sum_w, max_w = func(data.to(device), weights.clone().to(device), None, None, None)
        out, sum_w, max_w = out.cpu(), sum_w.cpu(), max_w.cpu()
        assert sum_w.shape == (bs, 1, h, w) and max_w.shape == (bs, 1, h, w)
        assert abs(out[0, 0, y, x].item() - val) < 1e-4
        if splat:
            for dy in range(-(k // 2), k // 2 + 1):
                for dx in range(-(k // 2), k // 2 + 1):
                    assert abs(out[0, 0, y + dy, x + dx].item() - val) < 1e-4
        else:
            assert abs(sum_w[0, 0, y, x].item() - k * k) < 1e-4
        with pytest.raises(RuntimeError):
            func(data.to(device), weights.clone().to(device), None, sum_w.to(device), None)


def _softmax_splat_reference(radiance, logits, k):
    """float64 restatement of `spp` progressive splat updates followed by the
    normalization: every sample's scatter kernel is softmax-normalized jointly
    with all other contributions landing on the same pixel (out-of-image sources
    contribute a zero logit to the denominator only)."""
    bs, spp, c, h, w = radiance.shape
    c0 = (k - 1) // 2
    L = logits.double().view(bs, spp, k, k, h, w)
    R = radiance.double()
    pad = (c0, c0, c0, c0)
    Lp = F.pad(L, pad)
    Rp = F.pad(R, pad)
    # gather form: pixel q receives from p = q + (dy - c0, dx - c0) the logit
    # S[K-1-dy, K-1-dx, p]
    G = th.zeros(bs, spp, k, k, h, w, dtype=th.float64)
    for dy in range(k):
        for dx in range(k):
            G[:, :, dy, dx] = Lp[:, :, k - 1 - dy, k - 1 - dx, dy:dy + h, dx:dx + w]
    m = G.amax(dim=(1, 2, 3), keepdim=True)
    E = th.exp(G - m)
    sum_w = E.sum(dim=(1, 2, 3))
    sum_r = th.zeros(bs, c, h, w, dtype=th.float64)
    for dy in range(k):
        for dx in range(k):
            sum_r += (E[:, :, dy, dx].unsqueeze(2) * Rp[:, :, :, dy:dy + h, dx:dx + w]).sum(1)
    return sum_r, sum_w.unsqueeze(1), m.view(bs, 1, h, w)


def _progressive_equals_joint_softmax(device):
    th.manual_seed(3)
    bs, spp, c, h, w, k = 2, 3, 3, 12, 16, 5
    radiance = th.rand(bs, spp, c, h, w)
    logits = 3 * th.randn(bs, spp, k * k, h, w)
    func = modules.ProgressiveKernelApply(splat=True)
    sum_r = sum_w = max_w = None
    for sp in range(spp):
        sum_r, sum_w, max_w = func(radiance[:, sp].to(device),
                                   logits[:, sp].contiguous().to(device), sum_r, sum_w, max_w)
    rr, rw, rm = _softmax_splat_reference(radiance, logits, k)
    assert th.allclose(max_w.cpu().double(), rm, atol=0, rtol=0)
    assert th.allclose(sum_w.cpu().double(), rw, rtol=2e-5)
    assert th.allclose(sum_r.cpu().double(), rr, rtol=2e-5, atol=1e-6)


def test_kernel_apply_cpu_logic(oracle_ops):
    _kernel_apply_kat("cpu")


def test_progressive_kernel_apply_cpu_logic(oracle_ops):
    _progressive_kat("cpu")
    _progressive_equals_joint_softmax("cpu")


def _tiny_multisteps(device, train):
    th.manual_seed(0)
    net = models.Multisteps(6, 2, width=8, embedding_width=8, ksize=5, nsteps=2).to(device)
    net.train(train)
    bs, spp, h, w = 1, 3, 24, 20
    samples = {"radiance": th.rand(bs, spp, 3, h, w, device=device),
               "features": th.randn(bs, spp, 6, h, w, device=device),
               "global_features": th.randn(bs, 2, 1, 1, device=device)}
    return net, samples


def test_multisteps_forward_cpu_logic(oracle_ops):
    net, samples = _tiny_multisteps("cpu", train=True)
    out = net(samples)["radiance"]
    assert out.shape == (1, 3, 20, 16)
    assert th.isfinite(out).all()
    out.mean().backward()
    assert all(p.grad is not None for p in net.kernel_regressor.parameters())
    # bs == 1: the eval path (one sample at a time) gives the same image
    net.eval()
    with th.no_grad():
        out_eval = net(samples)["radiance"]
    assert th.allclose(out, out_eval, rtol=1e-4, atol=1e-5)
    # the output is a convex combination of the sample radiances
    assert out_eval.min() >= 0 and out_eval.max() <= 1 + 1e-5


def test_kpcn_forward_cpu_logic(oracle_ops):
    th.manual_seed(0)
    net = models.KPCN(4, ksize=3, depth=2, width=8)
    h = w = 20
    data = {"kpcn_diffuse_in": th.randn(1, 4, h, w), "kpcn_specular_in": th.randn(1, 4, h, w),
            "kpcn_diffuse_buffer": th.rand(1, 3, h, w), "kpcn_specular_buffer": th.rand(1, 3, h, w),
            "kpcn_albedo": th.rand(1, 3, h, w)}
    out = net(data)
    assert out["radiance"].shape == (1, 3, h - 8, w - 8)
    assert th.allclose(out["radiance"],
                       data["kpcn_albedo"][..., 4:-4, 4:-4] * out["diffuse"]
                       + th.exp(out["specular"]) - 1)


# -- GPU: the same checks on the real ops ----------------------------------------
@pytest.mark.gpu
def test_kernel_apply_gpu():
    _kernel_apply_kat("cuda")


@pytest.mark.gpu
def test_progressive_kernel_apply_gpu():
    _progressive_kat("cuda")
    _progressive_equals_joint_softmax("cuda")


@pytest.mark.gpu
def test_multisteps_forward_gpu_matches_oracle_backed_cpu(monkeypatch):
    # fp32 parity: keep cuDNN / cuBLAS off TF32 for this comparison
    monkeypatch.setattr(th.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(th.backends.cuda.matmul, "allow_tf32", False)
    net, samples = _tiny_multisteps("cuda", train=True)
    out = net(samples)["radiance"]
    out.mean().backward()
    grads = [p.grad.detach().cpu().clone() for p in net.kernel_regressor.parameters()]
    net.eval()
    with th.no_grad():
        out_eval = net(samples)["radiance"]
    assert th.allclose(out, out_eval, rtol=1e-4, atol=1e-5)
    # same model on CPU with the oracle standing in for the custom ops
    KW, S2G = kats.oracle_functions()
    monkeypatch.setattr(funcs, "KernelWeighting", KW)
    monkeypatch.setattr(funcs, "Scatter2Gather", S2G)
    cpu_net = net.cpu().train()
    cpu_net.zero_grad()
    ref = cpu_net({k: v.cpu() for k, v in samples.items()})["radiance"]
    assert th.allclose(out.detach().cpu(), ref, rtol=1e-4, atol=1e-5)
    ref.mean().backward()
    for g, p in zip(grads, cpu_net.kernel_regressor.parameters()):
        assert th.allclose(g, p.grad, rtol=1e-3, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("splat", [True, False])
@pytest.mark.parametrize("shape", [(2, 3, 24, 132, 5), (1, 3, 30, 256, 21), (1, 3, 9, 18, 7),
                                   (2, 5, 12, 16, 3), (1, 2, 10, 12, 5)])
def test_fused_progressive_matches_composed(shape, splat):
    """The single-pass kernel against the reference chain built from the
    individual ops (Scatter2Gather -> max -> exp -> KernelWeighting)."""
    from sbmc_b200 import _lib
    bs, c, h, w, k = shape
    th.manual_seed(sum(shape))
    spp = 3
    radiance = th.rand(bs, spp, c, h, w, device="cuda")
    logits = 4 * th.randn(bs, spp, k * k, h, w, device="cuda")
    fused = modules.ProgressiveKernelApply(splat=splat)
    composed = modules.ProgressiveKernelApply(splat=splat)
    composed.fused = False
    a = (None, None, None)
    b = (None, None, None)
    with th.no_grad():
        for sp in range(spp):
            before = _lib.launch_count()
            a = fused(radiance[:, sp], logits[:, sp].clone(), *a)
            assert _lib.launch_count() == before + 1          # one kernel per update
            b = composed(radiance[:, sp], logits[:, sp].clone(), *b)
            assert th.equal(a[2], b[2])                        # max_w: exact
            # two fp32 evaluations of a K*K-term sum in different orders: at the
            # image border the composed chain adds hundreds of identical tiny
            # terms (zero logits of out-of-image taps) one by one, which costs it
            # up to ~K*K*eps; the fused kernel sums hierarchically
            assert th.allclose(a[1], b[1], rtol=1e-4, atol=0)
            assert th.allclose(a[0], b[0], rtol=1e-4, atol=1e-5)
    if splat:
        # against exact math (float64): the 1e-5 bar, for both paths
        rr, rw, rm = _softmax_splat_reference(radiance.cpu(), logits.cpu(), k)
        for name, got in (("fused", a), ("composed", b)):
            ew = ((got[1].cpu().double() - rw).abs() / rw).max().item()
            er = ((got[0].cpu().double() - rr).abs() / (rr.abs() + rw)).max().item()
            print("%s: max rel err sum_w %.2e sum_r %.2e" % (name, ew, er))
            bar = 1e-5 if name == "fused" else 1e-4
            assert ew <= bar and er <= bar, (name, ew, er)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 3, 24, 132, 5), (1, 3, 26, 256, 21), (1, 3, 12, 16, 3)])
@pytest.mark.parametrize("loss_kind", ["normalized", "raw"])
def test_fused_progressive_backward_matches_composed(shape, loss_kind):
    """Gradients of a 3-sample progressive splat through the fused autograd node
    against the reference chain of individual ops (both in fp32 on the GPU)."""
    bs, c, h, w, k = shape
    th.manual_seed(sum(shape))
    spp = 3
    radiance = th.rand(bs, spp, c, h, w, device="cuda")
    logits = 3 * th.randn(bs, spp, k * k, h, w, device="cuda")
    proj_r = th.randn(bs, c, h, w, device="cuda")
    proj_w = th.randn(bs, 1, h, w, device="cuda")

    def run(fused):
        mod = modules.ProgressiveKernelApply(splat=True)
        mod.fused = fused
        r = radiance.clone().requires_grad_(True)
        l = logits.clone().requires_grad_(True)
        state = (None, None, None)
        for sp in range(spp):
            # (the composed chain modifies its kernels in place: hand it a copy)
            state = mod(r[:, sp], l[:, sp] * 1.0, *state)
        sum_r, sum_w, max_w = state
        if loss_kind == "normalized":     # what Multisteps does (models.py:212)
            loss = ((sum_r / (sum_w + 1e-8)) * proj_r).sum()
        else:                              # the three outputs used independently
            loss = (sum_r * proj_r).sum() + (sum_w * proj_w).sum() + (max_w * proj_w).sum()
        loss.backward()
        return loss.item(), r.grad, l.grad

    la, ra, ka = run(True)
    lb, rb, kb = run(False)
    assert abs(la - lb) <= 1e-5 * abs(lb) + 1e-5
    scale_r, scale_k = rb.abs().max().item(), kb.abs().max().item()
    assert (ra - rb).abs().max().item() <= 1e-4 * scale_r, (ra - rb).abs().max().item() / scale_r
    assert (ka - kb).abs().max().item() <= 1e-4 * scale_k, (ka - kb).abs().max().item() / scale_k
    assert ((ka - kb).norm() / kb.norm()).item() <= 1e-5
    assert ((ra - rb).norm() / rb.norm()).item() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 16, 8, 5, 7, 10, 14), (1, 256, 128, 45, 80, 90, 160),
                                   (1, 8, 24, 3, 3, 7, 5)])
def test_fused_upsample_concat_matches_torch(shape):
    """cat([F.interpolate(coarse, size, bilinear), left], 1) in one pass (bf16 channels_last)."""
    n, cu, cs, hl, wl, h, w = shape
    th.manual_seed(0)
    coarse = th.randn(n, cu, hl, wl, device="cuda").to(th.bfloat16).contiguous(
        memory_format=th.channels_last)
    left = th.randn(n, cs, h, w, device="cuda").to(th.bfloat16).contiguous(
        memory_format=th.channels_last)
    assert modules._fused_upsample_ok(coarse, left)
    got = modules._upsample_concat(coarse, left)
    up = F.interpolate(coarse.float(), size=(h, w), mode="bilinear", align_corners=False)
    assert got.shape == (n, cu + cs, h, w)
    assert th.equal(got[:, cu:], left)
    # bf16 rounding of the interpolated value: half an ulp
    assert th.allclose(got[:, :cu].float(), up, rtol=2 ** -8, atol=1e-6)
    assert not modules._fused_upsample_ok(coarse.float(), left)


@pytest.mark.gpu
def test_unet_fast_path_matches_fp32_module():
    """The bf16 channels_last U-net path (cuDNN convs on cached folded weights +
    our bias/activation and upsample/concat kernels) against the fp32 module."""
    from sbmc_b200 import unet_fast
    th.manual_seed(0)
    net = modules.Autoencoder(16, 16, num_levels=3, increase_factor=2.0, num_convs=3, width=16,
                              ksize=3, output_type="leaky_relu", pooling="max").cuda().eval()
    with th.no_grad():
        for prm in net.parameters():
            prm.add_(0.05 * th.randn_like(prm))
    assert unet_fast.supports(net)
    x = th.randn(2, 16, 36, 52, device="cuda")
    with th.no_grad():
        ref = net(x)
        got = unet_fast.autoencoder_forward(net, x)
    assert got.dtype == th.bfloat16 and got.shape == ref.shape
    assert ((got.float() - ref).norm() / ref.norm()).item() < 3e-2
    assert not unet_fast.supports(modules.Autoencoder(16, 16, normalize=True))


@pytest.mark.gpu
def test_kpcn_forward_gpu_matches_oracle_backed_cpu(monkeypatch):
    """KPCN (sbmc/models.py:221-291): the second caller of KernelWeighting (gather
    kernels + softmax), GPU ops against the oracle-backed CPU run."""
    monkeypatch.setattr(th.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(th.backends.cuda.matmul, "allow_tf32", False)
    th.manual_seed(0)
    net = models.KPCN(4, ksize=5, depth=3, width=8).cuda().eval()
    h = w = 36
    data = {"kpcn_diffuse_in": th.randn(2, 4, h, w), "kpcn_specular_in": th.randn(2, 4, h, w),
            "kpcn_diffuse_buffer": th.rand(2, 3, h, w), "kpcn_specular_buffer": th.rand(2, 3, h, w),
            "kpcn_albedo": th.rand(2, 3, h, w)}
    with th.no_grad():
        got = net({k: v.cuda() for k, v in data.items()})
    KW, S2G = kats.oracle_functions()
    monkeypatch.setattr(funcs, "KernelWeighting", KW)
    monkeypatch.setattr(funcs, "Scatter2Gather", S2G)
    with th.no_grad():
        ref = net.cpu()(data)
    for key in ("radiance", "diffuse", "specular"):
        assert th.allclose(got[key].cpu(), ref[key], rtol=1e-4, atol=1e-5), key


@pytest.mark.gpu
def test_kernel_apply_softmax_gather_fused_matches_composed():
    """KPCN's KernelApply(softmax=True, splat=False): one fused pass without grad
    against softmax + KernelWeighting with grad enabled."""
    th.manual_seed(5)
    bs, c, h, w, k = 2, 3, 20, 36, 5
    data = th.rand(bs, c, h, w, device="cuda")
    logits = 3 * th.randn(bs, k * k, h, w, device="cuda")
    mod = modules.KernelApply(softmax=True, splat=False)
    with th.no_grad():
        out_f, sw_f = mod(data, logits)
    out_c, sw_c = mod(data, logits.clone().requires_grad_(True))
    assert th.allclose(out_f, out_c.detach(), rtol=2e-5, atol=1e-6)
    assert th.allclose(sw_f, sw_c.detach(), rtol=1e-5)


@pytest.mark.gpu
def test_random_small_shapes_fused_splat():
    """Seeded random shapes (ragged widths, K in 3..9, C in 1..5): the fused splat
    (tuned or generic kernel, whichever serves the shape) against the composed chain."""
    import random
    rng = random.Random(0)
    for _ in range(12):
        k = rng.choice([3, 5, 7, 9])
        c = rng.choice([1, 3, 3, 5])
        h, w = rng.randint(k, 30), rng.randint(k, 70)
        bs = rng.randint(1, 2)
        splat = rng.random() < 0.7
        th.manual_seed(h * w)
        rad = th.rand(bs, 2, c, h, w, device="cuda")
        logits = 3 * th.randn(bs, 2, k * k, h, w, device="cuda")
        fused = modules.ProgressiveKernelApply(splat=splat)
        comp = modules.ProgressiveKernelApply(splat=splat)
        comp.fused = False
        a = b = (None, None, None)
        with th.no_grad():
            for sp in range(2):
                a = fused(rad[:, sp], logits[:, sp].clone(), *a)
                b = comp(rad[:, sp], logits[:, sp].clone(), *b)
        assert th.equal(a[2], b[2]), (k, c, h, w, splat)
        assert th.allclose(a[1], b[1], rtol=1e-4), (k, c, h, w, splat)
        assert th.allclose(a[0], b[0], rtol=1e-4, atol=1e-5), (k, c, h, w, splat)
