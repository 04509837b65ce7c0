"""Mixed-precision training pipeline (sbmc_b200/train_pipeline.py): every stage against
the fp32 modules it replaces (sbmc/modules.py ConvChain / Autoencoder under autograd).

Bars: bf16 storage of the activations costs ~4e-3 per layer in value; gradients after
three 1x1 layers stay within 8 %, after the 15 convolutions of a U-net within 15 % in
norm (the same network under torch.autocast(bf16) through cuDNN measures the same)."""
import pytest
import torch as th

from sbmc_b200 import models, modules, train_pipeline as P

pytestmark = pytest.mark.gpu
BF = th.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _grads(module):
    return {k: p.grad.clone() for k, p in module.named_parameters()}


@pytest.mark.parametrize("with_ctx", [False, True])
def test_embed_stage_matches_fp32_chain(with_ctx):
    th.manual_seed(3)
    bs, spp, hw = 2, 3, 256
    cin = 256 if with_ctx else 96
    chain = modules.ConvChain(cin, 128, width=128, depth=3, ksize=1, pad=False).cuda()
    for p in chain.parameters():
        if p.dim() == 1:
            p.data.normal_(0, 0.1)
    x = th.randn(bs * spp * hw, 128, device="cuda").to(BF)
    if not with_ctx:
        x[:, 96:] = 0
    x.requires_grad_(with_ctx)
    ctxr = th.randn(bs * hw, 128, device="cuda").to(BF).requires_grad_(True) if with_ctx else None
    bank = P.WeightBank(P.chain_specs(chain, 256 if with_ctx else 128))
    e, red = P.embed_stage(bank, (0, 1, 2), bank.apply(), chain, x, ctxr, bs, spp, hw)
    ge = th.randn_like(e.float())
    gr = th.randn_like(red.float())
    ((e.float() * ge).sum() + (red.float() * gr).sum()).backward()
    got = _grads(chain)
    gx = x.grad.clone() if with_ctx else None
    gc = ctxr.grad.clone() if with_ctx else None
    chain.zero_grad()

    xr = x.detach().float().requires_grad_(True)
    full = xr[:, :96] if not with_ctx else None
    if with_ctx:
        cr = ctxr.detach().float().requires_grad_(True)
        full = th.cat([xr.view(bs, spp, hw, 128),
                       cr.view(bs, 1, hw, 128).expand(bs, spp, hw, 128)], 3).reshape(-1, 256)
    er = chain(full.t().reshape(1, cin, -1, 1)).reshape(128, -1).t()
    rr = er.view(bs, spp, hw, 128).mean(1).reshape(-1, 128)
    ((er * ge).sum() + (rr * gr).sum()).backward()
    assert rel(e, er) < 2e-2 and rel(red, rr) < 2e-2
    for k, p in chain.named_parameters():
        assert rel(got[k], p.grad) < 8e-2, k
    if with_ctx:
        assert rel(gx, xr.grad) < 8e-2
        assert rel(gc, cr.grad) < 8e-2


def test_regress_stage_matches_fp32_chain():
    th.manual_seed(4)
    bs, spp, h, w, k2 = 2, 2, 16, 16, 441
    hw = h * w
    chain = modules.ConvChain(256, k2, depth=3, width=128, ksize=1, activation="leaky_relu",
                              pad=False, output_type="linear").cuda()
    e = th.randn(bs * spp * hw, 128, device="cuda").to(BF).requires_grad_(True)
    c = th.randn(bs * hw, 128, device="cuda").to(BF).requires_grad_(True)
    bank = P.WeightBank(P.chain_specs(chain, 256, 512))
    outs = P.regress_stage(bank, (0, 1, 2), bank.apply(), chain, e, c, bs, spp, hw)
    assert len(outs) == spp and outs[0].shape == (bs, k2, hw)
    gs = [th.randn(bs, k2, hw, device="cuda") for _ in range(spp)]
    gs[1][:, 100:] = 0
    sum((o * g).sum() for o, g in zip(outs, gs)).backward()
    got = _grads(chain)
    ge, gc = e.grad.clone(), c.grad.clone()
    chain.zero_grad()
    er = e.detach().float().requires_grad_(True)
    cr = c.detach().float().requires_grad_(True)
    full = th.cat([er.view(bs, spp, hw, 128), cr.view(bs, 1, hw, 128).expand(bs, spp, hw, 128)], 3)
    y = chain(full.reshape(-1, 256).t().reshape(1, 256, -1, 1)).reshape(k2, bs, spp, hw)
    sum((y[:, :, s].permute(1, 0, 2) * gs[s]).sum() for s in range(spp)).backward()
    for s in range(spp):
        assert rel(outs[s], y[:, :, s].permute(1, 0, 2)) < 2e-2
    for k, p in chain.named_parameters():
        assert rel(got[k], p.grad) < 8e-2, k
    assert rel(ge, er.grad) < 8e-2 and rel(gc, cr.grad) < 8e-2


@pytest.mark.parametrize("h,w", [(32, 48), (20, 36)])
def test_unet_stage_matches_fp32_autoencoder(h, w):
    th.manual_seed(5)
    net = modules.Autoencoder(128, 128, num_levels=3, increase_factor=2.0, num_convs=3, width=128,
                              ksize=3, output_type="leaky_relu", pooling="max").cuda()
    for p in net.parameters():
        if p.dim() == 1:
            p.data.normal_(0, 0.05)
    x = th.randn(2, h, w, 128, device="cuda").to(BF).requires_grad_(True)
    plan, specs = P.unet_specs(net)
    bank = P.WeightBank(specs)
    y = P.unet_stage(bank, plan, list(range(len(specs))), bank.apply(), x)
    g = th.randn_like(y.float())
    (y.float() * g).sum().backward()
    got = _grads(net)
    gx = x.grad.clone()
    net.zero_grad()
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    prev = th.backends.cudnn.allow_tf32
    th.backends.cudnn.allow_tf32 = False
    try:
        yr = net(xr)
        (yr * g.permute(0, 3, 1, 2)).sum().backward()
    finally:
        th.backends.cudnn.allow_tf32 = prev
    assert rel(y.permute(0, 3, 1, 2), yr) < 3e-2
    assert rel(gx.permute(0, 3, 1, 2), xr.grad) < 0.15
    num = sum(((got[k] - p.grad) ** 2).sum() for k, p in net.named_parameters())
    den = sum((p.grad ** 2).sum() for p in net.parameters())
    assert (num / den).sqrt().item() < 0.15
    worst = max(rel(got[k], p.grad) for k, p in net.named_parameters() if p.grad.norm() > 1e-6)
    assert worst < 0.35, worst


def test_weight_bank_matches_torch_weight_norm():
    """One launch prepares the bf16 operands of every convolution; one launch turns the
    weight gradients into those of weight_v / weight_g (torch._weight_norm under autograd)."""
    th.manual_seed(8)
    convs = [th.nn.utils.weight_norm(th.nn.Conv2d(ci, co, k, padding=k // 2)).cuda()
             for ci, co, k in ((96, 128, 1), (128, 441, 1), (128, 256, 3), (384, 128, 3), (256, 128, 1))]
    for c in convs:
        c.weight_g.data.uniform_(0.5, 2.0)
    bank = P.WeightBank([(convs[0], 128, 0), (convs[1], 0, 512), (convs[2], 0, 0), (convs[3], 0, 0),
                         (convs[4], 0, 0)])
    toks = bank.apply()
    gs = []
    for k, c in enumerate(convs):
        w = th._weight_norm(c.weight_v, c.weight_g, 0).detach()
        co, ci, kh, _ = w.shape
        f, d = bank.fwd(k).float(), bank.dgrad(k).float()
        if kh == 1:
            assert rel(f[:co, :ci], w.view(co, ci)) < 3e-3 and rel(d[:ci, :co], w.view(co, ci).t()) < 3e-3
            assert (f[co:] == 0).all() and (f[:, ci:] == 0).all()
            assert (d[ci:] == 0).all() and (d[:, co:] == 0).all()
        else:
            assert rel(f, w.permute(2, 3, 0, 1).reshape(9, co, ci)) < 3e-3
            assert rel(d, w.flip(2, 3).permute(2, 3, 1, 0).reshape(9, ci, co)) < 3e-3
        g = th.randn(toks[k].shape, device="cuda")
        bank.dw(k).copy_(g)
        gs.append(g)
    th.autograd.backward(list(toks), [bank.dw(k) for k in range(len(convs))])
    for k, c in enumerate(convs):
        v = c.weight_v.detach().clone().requires_grad_(True)
        g = c.weight_g.detach().clone().requires_grad_(True)
        w = th._weight_norm(v, g, 0)
        co, ci, kh, _ = w.shape
        gw = gs[k].view(kh, kh, co, ci).permute(2, 3, 0, 1) if kh == 3 else gs[k].view(co, ci, 1, 1)
        w.backward(gw)
        assert rel(c.weight_v.grad, v.grad) < 1e-5, k
        assert rel(c.weight_g.grad, g.grad) < 1e-5, k


def test_multisteps_pipeline_is_used_and_close_to_fp32():
    th.manual_seed(0)
    net = models.Multisteps(12, 3, ksize=5, nsteps=2).cuda().train()
    bs, spp, h, w = 2, 2, 32, 48
    batch = {"radiance": th.rand(bs, spp, 3, h, w, device="cuda"),
             "features": th.randn(bs, spp, 12, h, w, device="cuda"),
             "global_features": th.randn(bs, 3, 1, 1, device="cuda")}
    assert P.supported(net, 12, 3, h, w)
    prev = th.backends.cudnn.allow_tf32
    th.backends.cudnn.allow_tf32 = False
    try:
        ref = net(batch)["radiance"]
        ref.square().mean().backward()
        want = _grads(net)
        net.zero_grad()
        net.bf16_train = True
        got = net(batch)["radiance"]
        got.square().mean().backward()
    finally:
        th.backends.cudnn.allow_tf32 = prev
    assert rel(got, ref) < 3e-2
    num = sum(((p.grad - want[k]) ** 2).sum() for k, p in net.named_parameters())
    den = sum((want[k] ** 2).sum() for k in want)
    assert (num / den).sqrt().item() < 0.2
    assert all(p.grad is not None and th.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("bf16", [True, False])
def test_cuda_graph_step_follows_the_eager_step(bf16):
    """interfaces.SampleBasedDenoiserInterface(cuda_graph=True): three replayed steps on
    three different batches against the same three steps run eagerly."""
    from sbmc_b200 import interfaces

    def make():
        th.manual_seed(1)
        net = models.Multisteps(12, 3, ksize=5, nsteps=2).cuda().train()
        net.bf16_train = bf16
        return net
    bs, spp, h, w = 2, 2, 32, 48
    g = th.Generator(device="cuda").manual_seed(5)
    batches = [{"radiance": th.rand(bs, spp, 3, h, w, device="cuda", generator=g),
                "features": th.randn(bs, spp, 12, h, w, device="cuda", generator=g),
                "global_features": th.randn(bs, 3, 1, 1, device="cuda", generator=g),
                "target_image": th.rand(bs, 3, h, w, device="cuda", generator=g)}
               for _ in range(3)]
    eager = interfaces.SampleBasedDenoiserInterface(make(), lr=1e-3, cuda=True, fused_optimizer=True)
    graph = interfaces.SampleBasedDenoiserInterface(make(), lr=1e-3, cuda=True, fused_optimizer=True,
                                                    cuda_graph=True)
    for b in batches:
        fe, be = eager.train_step(dict(b))
        fg, bg = graph.train_step(dict(b))
        assert abs(be["loss"] - bg["loss"]) <= 1e-4 * abs(be["loss"]) + 1e-7
        assert rel(fg["radiance"], fe["radiance"]) < 1e-4
    # same arithmetic in both up to the last bit of Adam's bias corrections (derived on the
    # device under capture); Adam's normalised update amplifies last-bit gradient
    # differences on components whose gradient is ~0 (each step moves a parameter by at
    # most lr = 1e-3), so the bar is per tensor, not per element
    for (k, p), q in zip(eager.model.named_parameters(), graph.model.parameters()):
        assert rel(q, p) < 2e-2, k
    num = sum(((q - p) ** 2).sum() for p, q in zip(eager.model.parameters(), graph.model.parameters()))
    den = sum((p ** 2).sum() for p in eager.model.parameters())
    assert (num / den).sqrt().item() < 1e-4
    assert float(graph.optimizer.state[next(graph.model.parameters())]["step"]) == 3.0
    assert len(graph._graphs) == 1


def test_cuda_graph_step_survives_an_optimizer_state_reload():
    """Checkpoint / resume under cuda_graph: loading an optimizer state dict replaces the
    state tensors, the captured step is rebuilt and training continues where the
    uninterrupted run goes."""
    from sbmc_b200 import interfaces

    def make():
        th.manual_seed(2)
        net = models.Multisteps(12, 3, ksize=5, nsteps=1).cuda().train()
        net.bf16_train = True
        return interfaces.SampleBasedDenoiserInterface(net, lr=1e-3, cuda=True, fused_optimizer=True,
                                                       cuda_graph=True)
    bs, spp, h, w = 2, 2, 16, 16
    g = th.Generator(device="cuda").manual_seed(9)
    batches = [{"radiance": th.rand(bs, spp, 3, h, w, device="cuda", generator=g),
                "features": th.randn(bs, spp, 12, h, w, device="cuda", generator=g),
                "global_features": th.randn(bs, 3, 1, 1, device="cuda", generator=g),
                "target_image": th.rand(bs, 3, h, w, device="cuda", generator=g)}
               for _ in range(4)]
    straight = make()
    for b in batches:
        straight.train_step(dict(b))
    first = make()
    for b in batches[:2]:
        first.train_step(dict(b))
    resumed = make()
    resumed.train_step(dict(batches[0]))                     # a graph exists before the reload
    resumed.model.load_state_dict(first.model.state_dict())
    resumed.optimizer.load_state_dict(first.optimizer.state_dict())
    for b in batches[2:]:
        resumed.train_step(dict(b))
    assert float(resumed.optimizer.state[next(resumed.model.parameters())]["step"]) == 4.0
    num = sum(((q - p) ** 2).sum() for p, q in zip(straight.model.parameters(), resumed.model.parameters()))
    den = sum((p ** 2).sum() for p in straight.model.parameters())
    assert (num / den).sqrt().item() < 1e-4


def test_training_curves_of_the_pipeline_and_the_fp32_path_agree():
    """40 Adam steps on a fixed set of batches: the mixed-precision pipeline (one CUDA graph)
    and the reference-arithmetic fp32 path start from the same weights, both reduce the
    loss, and their loss curves stay within a few percent of each other."""
    from sbmc_b200 import interfaces
    bs, spp, h, w = 2, 2, 32, 32
    g = th.Generator(device="cuda").manual_seed(11)
    target = th.rand(bs, 3, h, w, device="cuda", generator=g)
    batches = []
    for _ in range(4):
        rad = target.unsqueeze(1) + 0.3 * th.randn(bs, spp, 3, h, w, device="cuda", generator=g)
        batches.append({"radiance": rad.clamp_min(0),
                        "features": th.cat([rad, th.randn(bs, spp, 9, h, w, device="cuda", generator=g)], 2),
                        "global_features": th.randn(bs, 3, 1, 1, device="cuda", generator=g),
                        "target_image": target})
    curves = {}
    for name, bf16, graph in (("fp32", False, False), ("pipeline", True, True)):
        th.manual_seed(3)
        net = models.Multisteps(12, 3, ksize=5, nsteps=2).cuda().train()
        net.bf16_train = bf16
        iface = interfaces.SampleBasedDenoiserInterface(net, lr=3e-4, cuda=True, fused_optimizer=True,
                                                        cuda_graph=graph)
        curves[name] = [iface.train_step(dict(batches[i % 4]))[1]["loss"] for i in range(40)]
    a, b = th.tensor(curves["fp32"]), th.tensor(curves["pipeline"])
    assert a[-4:].mean() < 0.8 * a[:4].mean() and b[-4:].mean() < 0.8 * b[:4].mean()
    assert ((a - b).abs() / a).max().item() < 0.1
    assert ((a[-8:] - b[-8:]).abs() / a[-8:]).mean().item() < 0.05
