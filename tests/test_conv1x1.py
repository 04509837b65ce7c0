"""Fused tcgen05 1x1 ConvChain: host-side preparation on CPU; on the GPU the
kernel against (a) a torch emulation with the same bf16 rounding points (tight)
and (b) the plain fp32 ConvChain (bf16-level tolerance)."""
import pytest
import torch as th

from sbmc_b200 import conv1x1, models, modules


def _chains():
    th.manual_seed(0)
    return {
        "embedding_00": modules.ConvChain(96, 128, width=128, depth=3, ksize=1, pad=False),
        "embedding_01": modules.ConvChain(256, 128, width=128, depth=3, ksize=1, pad=False),
        "kernel_regressor": modules.ConvChain(256, 441, depth=3, width=128, ksize=1,
                                              activation="leaky_relu", pad=False,
                                              output_type="linear"),
        "small_out": modules.ConvChain(40, 24, depth=3, width=128, ksize=1, pad=False,
                                       weight_norm=False),
    }


def test_supports_and_prepare():
    for name, chain in _chains().items():
        assert conv1x1.supports(chain), name
        with th.no_grad():
            for prm in chain.parameters():       # non-trivial weight-norm gains / biases
                prm.add_(0.1 * th.randn_like(prm))
        p = conv1x1.prepare(chain)
        c1, c2, c3 = conv1x1._convs(chain)
        assert p.k1p in (128, 256) and p.k1p >= c1.in_channels and p.n3p % 16 == 0
        assert p.w1.shape == (128, p.k1p) and p.w3.shape == (p.n3p, 128)
        # the folded weights reproduce the convolution
        x = th.randn(2, c1.in_channels, 3, 5)
        want = c1(x)
        got = th.einsum("oc,nchw->nohw", conv1x1._effective_weight(c1), x) + c1.bias.view(1, -1, 1, 1)
        assert th.allclose(got, want, rtol=1e-4, atol=1e-5)
        assert (p.w1[:, c1.in_channels:] == 0).all() and (p.w3[c3.out_channels:] == 0).all()
        assert conv1x1.prepare(chain) is p       # cached until a parameter changes
        with th.no_grad():
            c2.bias.add_(1.0)
        assert conv1x1.prepare(chain) is not p
    assert not conv1x1.supports(modules.ConvChain(96, 128, width=64, depth=3, ksize=1))
    assert not conv1x1.supports(modules.ConvChain(96, 128, width=128, depth=3, ksize=3))
    assert not conv1x1.supports(modules.ConvChain(96, 128, width=128, depth=2, ksize=1))
    assert not conv1x1.supports(modules.ConvChain(96, 128, width=128, depth=3, ksize=1,
                                                  normalize=True))


def test_prepare_nhwc_channel_mapping():
    """Input channel order of a chain: [ca real channels | ngf broadcast features |
    cb second-source channels]; the NHWC kernel sees xa padded to 128 channels and
    the broadcast features folded into a per-image bias."""
    chains = _chains()
    emb0 = chains["embedding_00"]                       # 93 + 3 inputs
    p = conv1x1._prepare_nhwc(emb0, 93, 0, 3)
    w = conv1x1._effective_weight(conv1x1._convs(emb0)[0])
    assert p.w1.shape == (128, 128) and p.w1_gf.shape == (128, 3)
    assert th.equal(p.w1[:, :93], w[:, :93].to(th.bfloat16)) and (p.w1[:, 93:] == 0).all()
    assert th.equal(p.w1_gf, w[:, 93:96])
    reg = chains["kernel_regressor"]                    # 128 + 128 inputs
    p = conv1x1._prepare_nhwc(reg, 128, 128, 0)
    w = conv1x1._effective_weight(conv1x1._convs(reg)[0])
    assert p.w1.shape == (128, 256) and th.equal(p.w1, w.to(th.bfloat16))
    assert p.n3p == 448 and p.cout == 441 and p.act == 1
    assert conv1x1._prepare_nhwc(reg, 128, 128, 0) is p
    with pytest.raises(RuntimeError):
        conv1x1._prepare_nhwc(reg, 100, 128, 0)
    x = th.randn(2, 5, 3, 4)                            # CPU fallback of the layout change
    y = conv1x1.to_nhwc_bf16(x, channels=8)
    assert y.shape == (2, 12, 8) and (y[..., 5:] == 0).all()
    assert th.equal(y[..., :5], x.reshape(2, 5, 12).transpose(1, 2).to(th.bfloat16))


def _emulate(chain, x):
    """The chain with the kernel's rounding points: bf16 operands, fp32 accumulate."""
    p = conv1x1.prepare(chain)
    n, c, h, w = x.shape
    a = x.permute(0, 2, 3, 1).reshape(-1, c).to(th.bfloat16).double()
    act = (lambda t: th.nn.functional.leaky_relu(t, 0.01)) if p.act else th.relu
    h1 = act(a @ p.w1[:, :c].double().t() + p.b1.double()).float().to(th.bfloat16).double()
    h2 = act(h1 @ p.w2.double().t() + p.b2.double()).float().to(th.bfloat16).double()
    y = h2 @ p.w3[:p.cout].double().t() + p.b3[:p.cout].double()
    return y.float().reshape(n, h, w, p.cout).permute(0, 3, 1, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["embedding_00", "embedding_01", "kernel_regressor", "small_out"])
@pytest.mark.parametrize("hw", [(16, 24), (9, 13), (64, 128)])
def test_chain_matches_emulation_and_fp32(name, hw):
    chain = _chains()[name].cuda().eval()
    with th.no_grad():
        for prm in chain.parameters():
            prm.add_(0.05 * th.randn_like(prm))
    h, w = hw
    cin = conv1x1._convs(chain)[0].in_channels
    th.manual_seed(1)
    x = th.randn(3, cin, h, w, device="cuda")
    with th.no_grad():
        got = conv1x1.chain_forward(chain, x)
        emu = _emulate(chain, x)
        ref = chain(x)
    scale = ref.abs().max().item()
    # same rounding points: only the fp32 accumulation order differs -- except
    # where an activation lands on a bf16 rounding boundary (rare flips of 1 bf16 ulp)
    err = (got - emu).abs()
    assert err.max().item() <= 2e-2 * scale
    assert (err > 1e-4 * scale).float().mean().item() < 0.02
    assert ((got - emu).norm() / emu.norm()).item() < 2e-3
    # against the fp32 reference module: bf16-operand accuracy
    assert ((got - ref).norm() / ref.norm()).item() < 2e-2


@pytest.mark.gpu
def test_chain_two_sources_broadcast_and_strided_output():
    """cat([features[:, sp], ctx]) without the cat: second source as a tensor and
    as a per-image broadcast vector; output written into a strided slice."""
    chains = _chains()
    reg = chains["kernel_regressor"].cuda().eval()
    emb0 = chains["embedding_00"].cuda().eval()
    th.manual_seed(2)
    bs, spp, h, w = 2, 3, 20, 36
    feats = th.randn(bs, spp, 128, h, w, device="cuda")
    prop = th.randn(bs, 128, h, w, device="cuda")
    with th.no_grad():
        got = conv1x1.chain_forward(reg, feats[:, 1], prop)
        one = conv1x1.chain_forward(reg, th.cat([feats[:, 1], prop], 1))
        assert th.equal(got, one)
        f0 = th.randn(bs, spp, 93, h, w, device="cuda")
        gf = th.randn(bs, 3, 1, 1, device="cuda")
        new = th.zeros(bs, spp, 128, h, w, device="cuda")
        conv1x1.chain_forward(emb0, f0[:, 2], gf, out=new[:, 2])
        one = conv1x1.chain_forward(emb0, th.cat([f0[:, 2], gf.expand(bs, 3, h, w)], 1))
        assert th.equal(new[:, 2], one)
        assert (new[:, :2] == 0).all()


@pytest.mark.gpu
def test_multisteps_bf16_chains_close_to_fp32():
    th.manual_seed(0)
    net = models.Multisteps(12, 3, ksize=5, nsteps=2).cuda().eval()
    bs, spp, h, w = 2, 2, 32, 48
    samples = {"radiance": th.rand(bs, spp, 3, h, w, device="cuda"),
               "features": th.randn(bs, spp, 12, h, w, device="cuda"),
               "global_features": th.randn(bs, 3, 1, 1, device="cuda")}
    with th.no_grad():
        ref = net(samples)["radiance"]
        net.bf16_chains = True
        got = net(samples)["radiance"]
    assert got.shape == ref.shape
    assert ((got - ref).norm() / ref.norm()).item() < 3e-2


@pytest.mark.gpu
@pytest.mark.parametrize("name,two", [("embedding_00", False), ("embedding_01", True),
                                      ("kernel_regressor", True)])
@pytest.mark.parametrize("hw", [(16, 24), (9, 13), (64, 128)])
def test_nhwc_chain_matches_fp32_kernel_inputs(name, two, hw):
    """The TMA-fed NHWC variant against the emulation (same rounding points)."""
    chain = _chains()[name].cuda().eval()
    with th.no_grad():
        for prm in chain.parameters():
            prm.add_(0.05 * th.randn_like(prm))
    h, w = hw
    n = 3
    th.manual_seed(3)
    if name == "embedding_00":            # 93 features + 3 global features
        xa = th.randn(n, 93, h, w, device="cuda")
        gf = th.randn(n, 3, device="cuda")
        full = th.cat([xa, gf.view(n, 3, 1, 1).expand(n, 3, h, w)], 1)
        a_nhwc, ca, xb_nhwc = conv1x1.to_nhwc_bf16(xa), 93, None
    else:
        xa = th.randn(n, 128, h, w, device="cuda")
        xb = th.randn(n, 128, h, w, device="cuda")
        gf = None
        full = th.cat([xa, xb], 1)
        a_nhwc, ca, xb_nhwc = conv1x1.to_nhwc_bf16(xa), 128, conv1x1.to_nhwc_bf16(xb)
    nhwc_out = name != "kernel_regressor"
    with th.no_grad():
        got = conv1x1.chain_forward_nhwc(chain, a_nhwc, ca, xb=xb_nhwc, gf=gf, nhwc_out=nhwc_out)
        ref = chain(full)
    if nhwc_out:
        assert got.dtype == th.bfloat16 and got.shape == (n, h * w, 128)
        got = got.float().view(n, h, w, 128).permute(0, 3, 1, 2)
    else:
        got = got.view(n, -1, h, w)
    assert ((got - ref).norm() / ref.norm()).item() < 2e-2
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 0.1 * scale


@pytest.mark.gpu
def test_nhwc_chain_strided_views():
    reg = _chains()["embedding_01"].cuda().eval()
    th.manual_seed(4)
    bs, spp, h, w = 2, 3, 12, 20
    feats = conv1x1.to_nhwc_bf16(th.randn(bs, spp, 128, h, w, device="cuda"))
    prop = conv1x1.to_nhwc_bf16(th.randn(bs, 128, h, w, device="cuda"))
    new = th.zeros(bs, spp, h * w, 128, device="cuda", dtype=th.bfloat16)
    with th.no_grad():
        conv1x1.chain_forward_nhwc(reg, feats[:, 1], 128, xb=prop, out=new[:, 2])
        one = conv1x1.chain_forward_nhwc(reg, feats[:, 1].contiguous(), 128, xb=prop)
    assert th.equal(new[:, 2], one)
    assert (new[:, :2] == 0).all()


@pytest.mark.gpu
def test_to_nhwc_bf16_kernel_matches_torch():
    th.manual_seed(6)
    for shape in [(2, 3, 93, 9, 13), (1, 1, 128, 16, 32), (3, 5, 7, 11)]:
        x = th.randn(*shape, device="cuda")
        got = conv1x1.to_nhwc_bf16(x)
        c, h, w = shape[-3:]
        want = th.zeros(shape[:-3] + (h * w, 128), device="cuda", dtype=th.bfloat16)
        want[..., :c] = x.reshape(shape[:-3] + (c, h * w)).transpose(-1, -2)
        assert th.equal(got, want)


# -- the pipelined all-samples kernel (csrc/chain_v3.cu) ---------------------------
def _sample_inputs(name, n, spp, h, w, seed):
    th.manual_seed(seed)
    if name == "embedding_00":            # 93 features + 3 global features, no prop
        f = th.randn(n, spp, 93, h, w, device="cuda")
        gf = th.randn(n, 3, device="cuda")
        return f, 93, None, gf
    f = th.randn(n, spp, 128, h, w, device="cuda")
    prop = th.randn(n, 128, h, w, device="cuda")
    return f, 128, prop, None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["embedding_00", "embedding_01", "kernel_regressor"])
@pytest.mark.parametrize("n,spp,hw", [(1, 1, (9, 13)), (2, 2, (16, 24)), (3, 3, (9, 13)),
                                      (1, 4, (64, 128)), (2, 5, (20, 36))])
def test_pipelined_chain_matches_the_serial_kernel_and_fp32(name, n, spp, hw):
    """All samples in one launch (two streams in lock step, hidden activations in
    tensor memory, sample mean accumulated by the tensor cores) against the serial
    per-sample kernel (same rounding points) and the fp32 module."""
    chain = _chains()[name].cuda().eval()
    with th.no_grad():
        for prm in chain.parameters():
            prm.add_(0.05 * th.randn_like(prm))
    h, w = hw
    f, ca, prop, gf = _sample_inputs(name, n, spp, h, w, 11)
    feats = conv1x1.to_nhwc_bf16(f)                                   # [n, spp, hw, 128]
    prop_n = None if prop is None else conv1x1.to_nhwc_bf16(prop)
    regress = name == "kernel_regressor"
    with th.no_grad():
        if regress:
            got = conv1x1.chain_samples_nhwc(chain, feats, ca, prop=prop_n, regress=True)
            assert got.shape == (n, spp, 441, h * w) and got.dtype == th.float32
        else:
            got, mean = conv1x1.chain_samples_nhwc(chain, feats, ca, prop=prop_n, gf=gf,
                                                   want_mean=True, mean_dtype=th.float32)
            assert got.shape == (n, spp, h * w, 128) and mean.shape == (n, h * w, 128)
        acc = 0
        for s in range(spp):
            one = conv1x1.chain_forward_nhwc(chain, feats[:, s], ca, xb=prop_n, gf=gf,
                                             nhwc_out=not regress)
            ctx = prop if prop is not None else gf.view(n, 3, 1, 1).expand(n, 3, h, w)
            ref = chain(th.cat([f[:, s], ctx], 1))
            if regress:
                g = got[:, s]
                assert ((g - one).norm() / one.norm()).item() < 1e-5     # same rounding points
                g = g.view(n, 441, h, w)
            else:
                g = got[:, s].float()
                assert ((g - one.float()).norm() / one.float().norm()).item() < 2e-3
                g = g.view(n, h, w, 128).permute(0, 3, 1, 2)
                acc = acc + ref
            assert ((g - ref).norm() / ref.norm()).item() < 2e-2
        if not regress:
            want = (acc / spp).permute(0, 2, 3, 1).reshape(n, h * w, 128)
            assert ((mean - want).norm() / want.norm()).item() < 2e-2
            # the mean is taken on the fp32 accumulators: at least as close to the fp32
            # module as the mean of the bf16-rounded outputs
            e_fused = (mean - want).norm().item()
            e_round = (got.float().mean(1) - want).norm().item()
            assert e_fused <= 1.05 * e_round + 1e-6


@pytest.mark.gpu
def test_pipelined_chain_sample_ranges_and_bf16_mean():
    """Sample sub-ranges (the regressor is run a few samples at a time to bound the
    logits buffer) and the bf16 mean output."""
    chains = _chains()
    reg = chains["kernel_regressor"].cuda().eval()
    emb = chains["embedding_01"].cuda().eval()
    n, spp, h, w = 2, 5, 12, 20
    f, ca, prop, _ = _sample_inputs("kernel_regressor", n, spp, h, w, 12)
    feats, prop_n = conv1x1.to_nhwc_bf16(f), conv1x1.to_nhwc_bf16(prop)
    with th.no_grad():
        full = conv1x1.chain_samples_nhwc(reg, feats, ca, prop=prop_n, regress=True)
        a = conv1x1.chain_samples_nhwc(reg, feats, ca, prop=prop_n, regress=True, sample0=0,
                                       nsamples=2)
        b = conv1x1.chain_samples_nhwc(reg, feats, ca, prop=prop_n, regress=True, sample0=2,
                                       nsamples=3)
        assert th.equal(th.cat([a, b], 1), full)
        out = th.zeros(n, spp, h * w, 128, device="cuda", dtype=th.bfloat16)
        _, m16 = conv1x1.chain_samples_nhwc(emb, feats, ca, prop=prop_n, out=out, want_mean=True,
                                            sample0=1, nsamples=3)
        o32, m32 = conv1x1.chain_samples_nhwc(emb, feats, ca, prop=prop_n, want_mean=True,
                                              mean_dtype=th.float32, sample0=1, nsamples=3)
        assert (out[:, 0] == 0).all() and (out[:, 4] == 0).all()
        assert th.equal(out[:, 1:4], o32[:, 1:4])
        assert m16.dtype == th.bfloat16 and th.equal(m16, m32.to(th.bfloat16))
