"""3x3 implicit-GEMM convolution (csrc/conv3x3.cu) against torch's fp32 conv2d on the
same bf16-rounded operands (tight: only the fp32 accumulation order differs) and the
U-net fast path built on it against the fp32 module."""
import pytest
import torch as th
import torch.nn.functional as F

from sbmc_b200 import conv3x3, modules, unet_fast


def test_prepare_weight_layout():
    w = th.randn(128, 64, 3, 3)
    w9 = conv3x3.prepare_weight(w)
    assert w9.shape == (9, 128, 64) and w9.dtype == th.bfloat16
    assert th.equal(w9[3 * 2 + 1], w[:, :, 2, 1].to(th.bfloat16))
    assert conv3x3.supports_conv(th.nn.Conv2d(128, 256, 3, padding=1))
    assert not conv3x3.supports_conv(th.nn.Conv2d(100, 256, 3, padding=1))
    assert not conv3x3.supports_conv(th.nn.Conv2d(128, 256, 3, padding=0))
    assert not conv3x3.supports_conv(th.nn.Conv2d(128, 256, 5, padding=2))


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout", [(128, 128), (64, 256), (384, 128), (256, 512)])
@pytest.mark.parametrize("n,h,w", [(1, 8, 128), (2, 7, 45), (1, 33, 300), (3, 2, 129),
                                   (4, 32, 32), (2, 64, 64), (1, 9, 85), (2, 5, 3), (1, 1, 1)])
@pytest.mark.parametrize("act", [0, 2])
def test_conv3x3_matches_torch(cin, cout, n, h, w, act):
    th.manual_seed(cin + cout + h + w)
    x = th.randn(n, h, w, cin, device="cuda").to(th.bfloat16)
    wt = (th.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)).to(th.bfloat16)
    bias = th.randn(cout, device="cuda")
    got = conv3x3.conv3x3_nhwc(x, conv3x3.prepare_weight(wt), bias, act=act)
    assert got.shape == (n, h, w, cout) and got.dtype == th.bfloat16
    prev = th.backends.cudnn.allow_tf32
    th.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1)
    finally:
        th.backends.cudnn.allow_tf32 = prev
    if act == 2:
        ref = F.leaky_relu(ref, 0.01)
    ref = ref.permute(0, 2, 3, 1)
    err = (got.float() - ref).abs()
    scale = ref.abs().max().item()
    # bf16 output rounding: half an ulp = 2^-9 relative
    assert err.max().item() <= 6e-3 * scale
    assert ((got.float() - ref).norm() / ref.norm()).item() < 3e-3


@pytest.mark.gpu
def test_unet_fast_path_with_own_convs_matches_fp32_module():
    th.manual_seed(0)
    net = modules.Autoencoder(128, 128, num_levels=3, increase_factor=2.0, num_convs=3,
                              width=128, ksize=3, output_type="leaky_relu", pooling="max").cuda().eval()
    x = th.randn(2, 128, 40, 72, device="cuda")
    prev = th.backends.cudnn.allow_tf32
    th.backends.cudnn.allow_tf32 = False
    try:
        with th.no_grad():
            ref = net(x)
            lib = unet_fast.autoencoder_forward(net, x, own_convs=False).float()
            got = unet_fast.autoencoder_forward(net, x, own_convs=True).float()
    finally:
        th.backends.cudnn.allow_tf32 = prev
    assert got.shape == ref.shape
    e_own = ((got - ref).norm() / ref.norm()).item()
    e_lib = ((lib - ref).norm() / ref.norm()).item()
    assert e_own < 3e-2 and e_own < 1.5 * e_lib + 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 128, 40, 72), (1, 256, 7, 9), (3, 8, 2, 2)])
def test_maxpool2x2_matches_torch(shape):
    th.manual_seed(1)
    x = th.randn(*shape, device="cuda").to(th.bfloat16).contiguous(memory_format=th.channels_last)
    got = unet_fast._maxpool2x2(x)
    want = F.max_pool2d(x, 2, 2)
    assert got.shape == want.shape and th.equal(got, want)


@pytest.mark.gpu
def test_upsample_concat_and_layout_change_kernels_after_their_rewrite():
    """Both kernels were re-indexed / re-tiled in round 2: same results as torch."""
    from sbmc_b200 import conv1x1
    th.manual_seed(2)
    for (n, cu, hl, wl, cs, h, w) in [(1, 256, 20, 36, 128, 40, 72), (2, 16, 3, 5, 8, 7, 11)]:
        low = th.randn(n, cu, hl, wl, device="cuda").to(th.bfloat16).contiguous(memory_format=th.channels_last)
        skip = th.randn(n, cs, h, w, device="cuda").to(th.bfloat16).contiguous(memory_format=th.channels_last)
        got = modules._upsample_concat(low, skip)
        up = F.interpolate(low.float(), size=(h, w), mode="bilinear", align_corners=False)
        want = th.cat([up, skip.float()], 1)
        assert got.shape == want.shape
        assert (got.float() - want).abs().max().item() <= 2e-2 * want.abs().max().item()
        assert th.equal(got[:, cu:], skip)
    for shape in [(2, 3, 93, 9, 13), (1, 1, 128, 16, 32), (3, 5, 7, 11), (1, 93, 64, 128)]:
        x = th.randn(*shape, device="cuda")
        got = conv1x1.to_nhwc_bf16(x)
        c, h, w = shape[-3:]
        want = th.zeros(shape[:-3] + (h * w, 128), device="cuda", dtype=th.bfloat16)
        want[..., :c] = x.reshape(shape[:-3] + (c, h * w)).transpose(-1, -2)
        assert th.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout", [(128, 128), (128, 256)])
@pytest.mark.parametrize("n,h,w", [(8, 32, 32), (2, 64, 64), (1, 9, 85), (3, 7, 20)])
def test_linear_mode_gives_the_same_result(cin, cout, n, h, w):
    """Narrow images: tiles of 256 consecutive pixels of the [H][W + 1] sequence against
    the 128-pixel row-segment tiling (same accumulation order: bit-identical)."""
    from sbmc_b200 import _lib
    th.manual_seed(6)
    x = th.randn(n, h, w, cin, device="cuda").to(th.bfloat16)
    w9 = conv3x3.prepare_weight((th.randn(cout, cin, 3, 3, device="cuda") / 34.0).to(th.bfloat16))
    bias = th.randn(cout, device="cuda")
    lib = _lib.load()
    prev = lib.sbmc_b200_conv3x3_linear(0)
    try:
        tiled = conv3x3.conv3x3_nhwc(x, w9, bias, act=1)
        lib.sbmc_b200_conv3x3_linear(1)
        linear = conv3x3.conv3x3_nhwc(x, w9, bias, act=1)
    finally:
        lib.sbmc_b200_conv3x3_linear(prev)
    assert th.equal(tiled, linear)


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w", [(1, 8, 128), (2, 7, 45), (1, 33, 300)])
def test_cta_pair_kernel_gives_the_same_result(n, h, w):
    """The tcgen05.mma.cta_group::2 kernel (two spatial tiles per CTA pair, half of every
    weight stage per CTA; odd tile counts end in a dummy tile) against the single-CTA
    kernel."""
    from sbmc_b200 import _lib
    th.manual_seed(5)
    x = th.randn(n, h, w, 128, device="cuda").to(th.bfloat16)
    wt = (th.randn(128, 128, 3, 3, device="cuda") / 34.0).to(th.bfloat16)
    bias = th.randn(128, device="cuda")
    w9 = conv3x3.prepare_weight(wt)
    prev = _lib.load().sbmc_b200_conv3x3_pair(0)
    try:
        one = conv3x3.conv3x3_nhwc(x, w9, bias, act=2)
        _lib.load().sbmc_b200_conv3x3_pair(1)
        two = conv3x3.conv3x3_nhwc(x, w9, bias, act=2)
    finally:
        _lib.load().sbmc_b200_conv3x3_pair(prev)
    assert th.equal(one, two)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,act", [(128, 128, 2), (256, 128, 1), (128, 256, 0)])
def test_conv3x3_autograd_matches_torch(cin, cout, act):
    """Conv3x3BiasAct (forward + data gradient on conv3x3.cu, weight gradient on cuDNN
    bf16) against torch autograd in fp32 on the same bf16-rounded operands."""
    th.manual_seed(7)
    n, h, w = 2, 12, 40
    x = th.randn(n, h, w, cin, device="cuda").to(th.bfloat16).requires_grad_(True)
    wt = (th.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)).to(th.bfloat16)
    w9 = conv3x3.prepare_weight(wt).requires_grad_(True)
    bias = th.randn(cout, device="cuda", requires_grad=True)
    y = conv3x3.Conv3x3BiasAct.apply(x, w9, bias, act)
    gy = th.randn_like(y)
    y.backward(gy)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.float().requires_grad_(True)
    br = bias.detach().clone().requires_grad_(True)
    prev = th.backends.cudnn.allow_tf32
    th.backends.cudnn.allow_tf32 = False
    try:
        yr = F.conv2d(xr, wr, br, padding=1)
        yr = F.relu(yr) if act == 1 else (F.leaky_relu(yr, 0.01) if act == 2 else yr)
        yr.backward(gy.float().permute(0, 3, 1, 2))
    finally:
        th.backends.cudnn.allow_tf32 = prev

    def rel(a, b):
        return ((a.float() - b).norm() / b.norm()).item()
    assert rel(y.permute(0, 3, 1, 2), yr.detach()) < 5e-3
    assert rel(x.grad.permute(0, 3, 1, 2), xr.grad) < 1e-2
    assert rel(w9.grad, wr.grad.permute(2, 3, 0, 1).reshape(9, cout, cin)) < 1e-2
    assert rel(bias.grad, br.grad) < 1e-2


@pytest.mark.gpu
def test_unet_training_path_gradients_close_to_fp32_module():
    th.manual_seed(3)
    net = modules.Autoencoder(128, 128, num_levels=3, increase_factor=2.0, num_convs=3,
                              width=128, ksize=3, output_type="leaky_relu", pooling="max").cuda().train()
    assert unet_fast.supports_training(net)
    x = th.randn(2, 128, 24, 40, device="cuda")
    gy = th.randn(2, 128, 24, 40, device="cuda")
    prev = th.backends.cudnn.allow_tf32
    th.backends.cudnn.allow_tf32 = False
    try:
        ref = net(x)
        ref.backward(gy)
        want = {k: p.grad.clone() for k, p in net.named_parameters()}
        net.zero_grad()
        # the library's mixed precision (autocast bf16 through cuDNN) as the yardstick for
        # what bf16 activations / gradients cost in a 15-convolution-deep network
        with th.autocast("cuda", dtype=th.bfloat16):
            lib = net(x)
        lib.float().backward(gy)
        lib_grads = {k: p.grad.clone() for k, p in net.named_parameters()}
        net.zero_grad()
        got = unet_fast.autoencoder_forward_train(net, x)
        got.backward(gy)
    finally:
        th.backends.cudnn.allow_tf32 = prev

    def grad_err(grads):
        num = sum(((grads[k] - want[k]) ** 2).sum() for k in want)
        den = sum((want[k] ** 2).sum() for k in want)
        return (num / den).sqrt().item()
    assert ((got - ref).norm() / ref.norm()).item() < 3e-2
    e_own = grad_err({k: p.grad for k, p in net.named_parameters()})
    e_lib = grad_err(lib_grads)
    assert e_own < 0.12 and e_own < 1.5 * e_lib + 1e-2, (e_own, e_lib)
