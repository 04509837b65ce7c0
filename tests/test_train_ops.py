"""Kernels of the mixed-precision training pipeline (sbmc_b200/train_ops.py) against
plain PyTorch fp32 references of the same operations (autograd where the kernel is a
backward pass)."""
import pytest
import torch as th
import torch.nn.functional as F

from sbmc_b200 import train_ops as T

pytestmark = pytest.mark.gpu
BF = th.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _act(x, act):
    return F.relu(x) if act == 1 else (F.leaky_relu(x, 0.01) if act == 2 else x)


def _dact(y, act):
    if act == 0:
        return th.ones_like(y)
    return th.where(y > 0, th.ones_like(y), th.full_like(y, 0.01 if act == 2 else 0.0))


@pytest.mark.parametrize("bs,spp,hw,ca,cb,cout", [(2, 3, 256, 128, 128, 128), (1, 2, 512, 128, 64, 256),
                                                  (2, 1, 256, 64, 128, 128)])
def test_linear_two_sources(bs, spp, hw, ca, cb, cout):
    th.manual_seed(1)
    x = th.randn(bs * spp * hw, ca, device="cuda").to(BF)
    xb = th.randn(bs * hw, cb, device="cuda").to(BF)
    w = (th.randn(cout, ca + cb, device="cuda") / (ca + cb) ** 0.5).to(BF)
    b = th.randn(cout, device="cuda")
    got = T.linear(x, w, b, 2, xb=xb, hw=hw, spp=spp)
    full = th.cat([x.view(bs, spp, hw, ca),
                   xb.view(bs, 1, hw, cb).expand(bs, spp, hw, cb)], 3).reshape(-1, ca + cb)
    ref = F.leaky_relu(full.float() @ w.float().t() + b, 0.01)
    assert rel(got, ref) < 4e-3


@pytest.mark.parametrize("rows,cin,cout,mact", [(700, 128, 128, 1), (513, 512, 128, 2), (256, 128, 256, 2)])
def test_linear_mask_epilogue(rows, cin, cout, mact):
    th.manual_seed(2)
    x = th.randn(rows, cin, device="cuda").to(BF)
    w = (th.randn(cout, cin, device="cuda") / cin ** 0.5).to(BF)
    m = th.randn(rows, cout, device="cuda").to(BF)
    m[::7, ::5] = 0                                     # derivative at exactly 0
    m[1::9, 3::4] = -0.0
    got = T.linear(x, w, None, 0, mask=m, mask_act=mact)
    ref = (x.float() @ w.float().t()) * _dact(m.float(), mact)
    assert rel(got, ref) < 4e-3
    # exact zeros where ReLU's derivative vanishes
    if mact == 1:
        assert (got[m.float() <= 0] == 0).all()


def test_linear_plane_output():
    th.manual_seed(3)
    bs, spp, hw, cin, cout, valid = 2, 3, 320, 128, 512, 441
    x = th.randn(bs * spp * hw, cin, device="cuda").to(BF)
    w = (th.randn(cout, cin, device="cuda") / cin ** 0.5).to(BF)
    b = th.randn(cout, device="cuda")
    out = th.full((spp, bs, valid, hw), float("nan"), device="cuda")
    T.linear(x, w, b, 0, hw=hw, spp=spp, out_mode=2, out=out, out_img_stride=valid * hw,
             out_smp_stride=bs * valid * hw, cout_valid=valid)
    ref = (x.float() @ w.float().t() + b).view(bs, spp, hw, cout)[..., :valid].permute(1, 0, 3, 2)
    assert th.isfinite(out).all()
    assert rel(out, ref) < 2e-5


@pytest.mark.parametrize("rows,cout,cin,cv,civ", [(1000, 128, 128, 0, 0), (4096 + 77, 512, 128, 441, 0),
                                                   (300, 128, 256, 0, 0), (65536, 128, 128, 0, 96),
                                                   (128, 128, 128, 0, 0)])
def test_wgrad_matches_matmul(rows, cout, cin, cv, civ):
    th.manual_seed(rows)
    dy = th.randn(rows, cout, device="cuda").to(BF)
    x = th.randn(rows, cin, device="cuda").to(BF)
    dw, db = T.wgrad(dy, x, cout_valid=cv, cin_valid=civ)
    ref = (dy.double().t() @ x.double())[:cv or cout, :civ or cin]
    assert dw.shape == ref.shape
    assert rel(dw, ref) < 1e-5
    assert rel(db, dy.double().sum(0)[:cv or cout]) < 1e-5


def test_wgrad_strided_operand_and_destination():
    th.manual_seed(5)
    rows = 2000
    dy = th.randn(rows, 128, device="cuda").to(BF)
    wide = th.randn(rows, 384, device="cuda").to(BF)
    dw = th.zeros(128, 384, device="cuda")
    T.wgrad(dy, wide[:, 128:256], dw=dw[:, 128:256], want_bias=False)
    ref = dy.double().t() @ wide[:, 128:256].double()
    assert rel(dw[:, 128:256], ref) < 1e-5
    assert (dw[:, :128] == 0).all() and (dw[:, 256:] == 0).all()


def test_wgrad_is_deterministic():
    th.manual_seed(6)
    dy = th.randn(30000, 128, device="cuda").to(BF)
    x = th.randn(30000, 128, device="cuda").to(BF)
    a = T.wgrad(dy, x)
    b = T.wgrad(dy, x)
    assert th.equal(a[0], b[0]) and th.equal(a[1], b[1])


@pytest.mark.parametrize("cin,cout,mact", [(128, 128, 1), (256, 128, 2), (128, 256, 1)])
def test_conv3x3_mask_epilogue(cin, cout, mact):
    th.manual_seed(7)
    n, h, w = 2, 18, 140
    x = th.randn(n, h, w, cin, device="cuda").to(BF)
    w9 = (th.randn(9, cout, cin, device="cuda") / (9 * cin) ** 0.5).to(BF)
    zero = th.zeros(cout, device="cuda")
    m = th.randn(n, h, w, cout, device="cuda").to(BF)
    got = T.conv3x3(x, w9, zero, 0, mask=m, mask_act=mact)
    wt = w9.float().view(3, 3, cout, cin).permute(2, 3, 0, 1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1)
    ref = ref * _dact(m.float(), mact)
    assert rel(got, ref) < 4e-3


def test_spp_reduce_and_bcast_add():
    th.manual_seed(8)
    n, spp, hw, c = 3, 5, 77, 128
    x = th.randn(n * spp * hw, c, device="cuda").to(BF)
    mean = T.spp_reduce(x, n, spp, 1.0 / spp)
    ref = x.float().view(n, spp, hw, c).mean(1).reshape(n * hw, c)
    assert rel(mean, ref) < 3e-3
    s32 = T.spp_reduce(x, n, spp, 1.0, out_f32=True)
    assert rel(s32, x.float().view(n, spp, hw, c).sum(1).reshape(n * hw, c)) < 1e-6
    r = th.randn(n * hw, c, device="cuda").to(BF)
    got = T.bcast_add(x, r, n, spp, 0.25)
    want = (x.float().view(n, spp, hw, c) + 0.25 * r.float().view(n, 1, hw, c)).reshape(-1, c)
    assert rel(got, want) < 3e-3
    only = T.bcast_add(None, r, n, spp, 2.0)
    assert rel(only, (2.0 * r.float()).view(n, 1, hw, c).expand(n, spp, hw, c).reshape(-1, c)) < 3e-3


@pytest.mark.parametrize("h,w,act", [(16, 24, 1), (17, 23, 2), (8, 8, 0)])
def test_maxpool_backward_with_skip_and_activation(h, w, act):
    th.manual_seed(9)
    n, c = 2, 64
    pre = th.randn(n, c, h, w, device="cuda")
    pre = (pre * 4).round() / 4                         # ties inside the windows
    pre = pre.to(BF).float().requires_grad_(True)
    x = _act(pre, act)
    pooled = F.max_pool2d(x, 2, 2)
    dpool = th.randn_like(pooled).to(BF).float()
    dcat = th.randn(n, h, w, 2 * c, device="cuda").to(BF)
    dskip = dcat[..., c:]
    (pooled * dpool).sum().backward(retain_graph=True)
    (x * dskip.float().permute(0, 3, 1, 2)).sum().backward()
    got = T.maxpool2x2_bwd(x.detach().permute(0, 2, 3, 1).contiguous().to(BF),
                           dpool.permute(0, 2, 3, 1).contiguous().to(BF), dskip, act)
    ref = pre.grad.permute(0, 2, 3, 1)
    assert rel(got, ref) < 4e-3
    nos = T.maxpool2x2_bwd(x.detach().permute(0, 2, 3, 1).contiguous().to(BF),
                           dpool.permute(0, 2, 3, 1).contiguous().to(BF), None, 0)
    assert th.isfinite(nos.float()).all()


@pytest.mark.parametrize("hl,wl,h,w,act", [(8, 12, 16, 24, 1), (8, 11, 17, 23, 0), (5, 5, 10, 10, 2)])
def test_upsample_backward(hl, wl, h, w, act):
    th.manual_seed(10)
    n, c = 2, 64
    pre = th.randn(n, c, hl, wl, device="cuda").to(BF).float().requires_grad_(True)
    coarse = _act(pre, act)
    up = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False)
    dcat = th.randn(n, h, w, 2 * c, device="cuda").to(BF)
    dup = dcat[..., :c]
    (up * dup.float().permute(0, 3, 1, 2)).sum().backward()
    got = T.upsample_bwd(dup, (hl, wl), coarse.detach().permute(0, 2, 3, 1).contiguous().to(BF), act)
    assert rel(got, pre.grad.permute(0, 2, 3, 1)) < 4e-3


def test_dact_and_colsum():
    th.manual_seed(11)
    y = th.randn(1000, 128, device="cuda").to(BF)
    g = th.randn(1000, 128, device="cuda").to(BF)
    for act in (1, 2):
        assert rel(T.dact(y, g, act), g.float() * _dact(y.float(), act)) < 3e-3
    for rows, c in ((1000, 128), (70000, 384), (5, 512)):
        x = th.randn(rows, c + 64, device="cuda").to(BF)
        assert rel(T.colsum(x[:, 64:]), x[:, 64:].double().sum(0)) < 1e-5


def test_planes_to_rows_into_a_row_buffer():
    th.manual_seed(12)
    bs, spp, hw, c, cpad = 2, 3, 200, 441, 512
    rows = th.zeros(bs, spp, hw, cpad, device="cuda", dtype=BF)
    g = th.randn(bs, c, hw, device="cuda")
    T.planes_to_rows(g, cpad, out=rows[:, 1], out_img_stride=spp * hw * cpad)
    assert rel(rows[:, 1, :, :c], g.permute(0, 2, 1)) < 3e-3
    assert (rows[:, 1, :, c:] == 0).all() and (rows[:, 0] == 0).all() and (rows[:, 2] == 0).all()


@pytest.mark.parametrize("n,h,w,cout,cin", [(2, 16, 64, 128, 128), (1, 17, 70, 128, 256), (3, 8, 32, 256, 128),
                                            (8, 128, 128, 128, 128), (2, 5, 130, 128, 384)])
def test_wgrad3x3_matches_conv2d_weight(n, h, w, cout, cin):
    th.manual_seed(n * h + w)
    dp = th.randn(n, h, w, cout, device="cuda").to(BF)
    x = th.randn(n, h, w, cin, device="cuda").to(BF)
    got, db = T.wgrad3x3(dp, x, want_bias=True)
    assert rel(db, dp.double().sum((0, 1, 2))) < 1e-5
    ref = th.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (cout, cin, 3, 3),
                                   dp.double().permute(0, 3, 1, 2), padding=1)
    ref9 = ref.permute(2, 3, 0, 1).reshape(9, cout, cin)
    assert rel(got, ref9) < 1e-5
    for t in range(9):
        assert rel(got[t], ref9[t]) < 1e-5, t
    assert th.equal(got, T.wgrad3x3(dp, x))
