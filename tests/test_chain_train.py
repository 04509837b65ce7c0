"""Opt-in mixed-precision training path: the tcgen05 GEMM layer (csrc/linear.cu), the
chain autograd Function built on it, and Multisteps.bf16_train against the fp32 model."""
import pytest
import torch as th
import torch.nn.functional as F

from sbmc_b200 import chain_train, models, modules


def test_chain_weights_padding_and_order():
    th.manual_seed(0)
    reg = modules.ConvChain(256, 441, depth=3, width=128, ksize=1, activation="leaky_relu",
                            pad=False, output_type="linear")
    w1, b1, w2, b2, w3, b3, act, cout = chain_train.chain_weights(reg, 256)
    assert w1.shape == (128, 256) and w2.shape == (128, 128) and w3.shape == (512, 128)
    assert b3.shape == (512,) and cout == 441 and act == 2
    assert (w3[441:] == 0).all() and (b3[441:] == 0).all()
    emb = modules.ConvChain(96, 128, width=128, depth=3, ksize=1, pad=False)
    w1 = chain_train.chain_weights(emb, 128)[0]
    assert w1.shape == (128, 128) and (w1[:, 96:] == 0).all()
    assert chain_train.chain_weights(emb, 128)[6] == 1
    # differentiable w.r.t. the module's parameters (weight normalization included)
    w1.float().sum().backward()
    assert emb.layer_0.layer[0].weight_g.grad is not None


@pytest.mark.gpu
@pytest.mark.parametrize("p,cin,cout", [(256, 128, 128), (1000, 256, 128), (77, 128, 512),
                                        (513, 64, 256), (300, 448, 128)])
@pytest.mark.parametrize("act,f32", [(0, True), (2, False), (1, False)])
def test_linear_matches_torch(p, cin, cout, act, f32):
    th.manual_seed(p + cin + cout)
    x = th.randn(p, cin, device="cuda").to(th.bfloat16)
    w = (th.randn(cout, cin, device="cuda") / cin ** 0.5).to(th.bfloat16)
    b = th.randn(cout, device="cuda")
    got = chain_train.linear_nhwc(x, w, b, act, th.float32 if f32 else th.bfloat16)
    ref = x.float() @ w.float().t() + b
    ref = F.relu(ref) if act == 1 else (F.leaky_relu(ref, 0.01) if act == 2 else ref)
    assert got.shape == ref.shape
    tol = 2e-5 if f32 else 4e-3
    assert ((got.float() - ref).norm() / ref.norm()).item() < tol
    nob = chain_train.linear_nhwc(x, w, None, 0, th.float32)
    assert ((nob - x.float() @ w.float().t()).norm() / nob.norm()).item() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["embedding", "regressor"])
def test_chain_fn_gradients(name):
    th.manual_seed(4)
    if name == "embedding":
        chain = modules.ConvChain(256, 128, width=128, depth=3, ksize=1, pad=False).cuda()
        f32 = False
    else:
        chain = modules.ConvChain(256, 441, depth=3, width=128, ksize=1, activation="leaky_relu",
                                  pad=False, output_type="linear").cuda()
        f32 = True
    p = 700
    x = th.randn(p, 256, device="cuda").to(th.bfloat16).requires_grad_(True)
    w1, b1, w2, b2, w3, b3, act, cout = chain_train.chain_weights(chain, 256)
    y = chain_train.ChainFn.apply(x, w1, b1, w2, b2, w3, b3, act, f32)[:, :cout]
    gy = th.randn(p, cout, device="cuda")
    y.float().backward(gy)
    got = {k: prm.grad.clone() for k, prm in chain.named_parameters()}
    gx = x.grad.clone()
    chain.zero_grad()
    xr = x.detach().float().requires_grad_(True)
    yr = chain(xr.t().reshape(1, 256, p, 1)).reshape(cout, p).t()
    yr.backward(gy)
    assert ((y.float() - yr).norm() / yr.norm()).item() < 2e-2
    # bf16 gradients between the layers: a few percent after three layers and two masks
    assert ((gx.float() - xr.grad).norm() / xr.grad.norm()).item() < 8e-2
    for k, prm in chain.named_parameters():
        err = ((got[k] - prm.grad).norm() / prm.grad.norm().clamp_min(1e-12)).item()
        assert err < 8e-2, (k, err)


@pytest.mark.gpu
def test_multisteps_bf16_train_close_to_fp32_training():
    th.manual_seed(0)
    net = models.Multisteps(12, 3, ksize=5, nsteps=2).cuda().train()
    bs, spp, h, w = 2, 2, 32, 48
    batch = {"radiance": th.rand(bs, spp, 3, h, w, device="cuda"),
             "features": th.randn(bs, spp, 12, h, w, device="cuda"),
             "global_features": th.randn(bs, 3, 1, 1, device="cuda")}
    prev = th.backends.cudnn.allow_tf32
    th.backends.cudnn.allow_tf32 = False
    try:
        ref = net(batch)["radiance"]
        ref.square().mean().backward()
        want = {k: p.grad.clone() for k, p in net.named_parameters()}
        net.zero_grad()
        net.bf16_train = True
        got = net(batch)["radiance"]
        got.square().mean().backward()
    finally:
        th.backends.cudnn.allow_tf32 = prev
    assert got.shape == ref.shape
    assert ((got - ref).norm() / ref.norm()).item() < 3e-2
    num = sum(((p.grad - want[k]) ** 2).sum() for k, p in net.named_parameters())
    den = sum((want[k] ** 2).sum() for k in want)
    assert (num / den).sqrt().item() < 0.2
    assert all(p.grad is not None and th.isfinite(p.grad).all() for p in net.parameters())
