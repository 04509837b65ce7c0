"""Fused 3-layer 1x1 ConvChain on the tcgen05 tensor cores (inference).

Binds ``sbmc_conv1x1_chain_f32`` (include/sbmc_b200.h, csrc/conv1x1.cu) to the
per-sample networks of the reference model (sbmc/models.py:86-102): the
``embedding_XX`` chains and the ``kernel_regressor`` are 1x1 convolutions, i.e. a
3-layer MLP per pixel.  bf16 operands, fp32 accumulation, fp32 output.
"""
import torch as th

from . import _lib

__all__ = ["supports", "prepare", "chain_forward", "chain_forward_nhwc", "chain_samples_nhwc",
           "to_nhwc_bf16"]

_HID = 128


def _convs(chain):
    """The three nn.Conv2d of a depth-3 ConvChain (layer_0, layer_1, prediction)."""
    return [chain.layer_0.layer[0], chain.layer_1.layer[0], chain.prediction]


def supports(chain):
    """Whether `chain` (a modules.ConvChain) has the shape the fused kernel serves:
    three 1x1 convolutions with hidden width 128, ReLU / LeakyReLU in between,
    linear output, no normalization layers, at most 256 input channels."""
    try:
        kids = dict(chain.named_children())
        if sorted(kids) != ["layer_0", "layer_1", "prediction"]:
            return False
        for name in ("layer_0", "layer_1"):
            seq = kids[name].layer
            if len(seq) != 2 or not isinstance(seq[1], (th.nn.ReLU, th.nn.LeakyReLU)):
                return False
            if isinstance(seq[1], th.nn.LeakyReLU) and abs(seq[1].negative_slope - 0.01) > 1e-12:
                return False
        c1, c2, c3 = _convs(chain)
        if type(chain.layer_0.layer[1]) is not type(chain.layer_1.layer[1]):
            return False
        for c in (c1, c2, c3):
            if c.kernel_size != (1, 1) or c.stride != (1, 1) or c.padding != (0, 0) \
                    or c.groups != 1 or c.bias is None:
                return False
        return (c1.out_channels == _HID and c2.in_channels == _HID and c2.out_channels == _HID
                and c3.in_channels == _HID and c1.in_channels <= 256 and c3.out_channels <= 512)
    except AttributeError:
        return False


def _effective_weight(conv):
    """[out, in] fp32 weight with the (old-style) weight normalization folded in."""
    if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
        w = th._weight_norm(conv.weight_v, conv.weight_g, 0)
    else:
        w = conv.weight
    return w.detach().reshape(w.shape[0], w.shape[1]).float()


class _Prepared(object):
    __slots__ = ("w1", "w2", "w3", "b1", "b2", "b3", "k1p", "cin", "cout", "n3p", "act",
                 "versions")


def _versions(chain):
    return tuple((p.data_ptr(), p._version) for p in chain.parameters())


def prepare(chain):
    """bf16 / padded copies of the chain's weights, cached on the module and
    refreshed when a parameter changes."""
    cached = getattr(chain, "_sbmc_b200_prepared", None)
    ver = _versions(chain)
    if cached is not None and cached.versions == ver:
        return cached
    c1, c2, c3 = _convs(chain)
    dev = c1.bias.device
    p = _Prepared()
    p.cin, p.cout = c1.in_channels, c3.out_channels
    p.k1p = 128 if p.cin <= 128 else 256
    p.n3p = (p.cout + 15) // 16 * 16
    w1 = th.zeros(_HID, p.k1p, device=dev)
    w1[:, :p.cin] = _effective_weight(c1)
    w3 = th.zeros(p.n3p, _HID, device=dev)
    w3[:p.cout] = _effective_weight(c3)
    b3 = th.zeros(p.n3p, device=dev)
    b3[:p.cout] = c3.bias.detach().float()
    p.w1 = w1.to(th.bfloat16).contiguous()
    p.w2 = _effective_weight(c2).to(th.bfloat16).contiguous()
    p.w3 = w3.to(th.bfloat16).contiguous()
    p.b1 = c1.bias.detach().float().contiguous()
    p.b2 = c2.bias.detach().float().contiguous()
    p.b3 = b3
    p.act = 1 if isinstance(chain.layer_0.layer[1], th.nn.LeakyReLU) else 0
    p.versions = ver
    object.__setattr__(chain, "_sbmc_b200_prepared", p)
    return p


def _image_view(x):
    """(tensor, elements between images) for x [n, c, h, w] whose images are
    contiguous [c, h, w] blocks (the batch stride may be anything, e.g. a
    `features[:, sp]` slice)."""
    n, c, h, w = x.shape
    if x.dtype != th.float32:
        x = x.float()
    if x.stride()[1:] != (h * w, w, 1):
        x = x.contiguous()
    return x, (x.stride(0) if n > 1 else c * h * w)


def chain_forward(chain, xa, xb=None, out=None):
    """act-chain(th.cat([xa, xb], 1)) without materialising the concatenation.

    xa [n, ca, h, w]; xb [n, cb, h, w], or [n, cb, 1, 1] (broadcast over the
    pixels, the global features), or None.  `out` may be a preallocated
    [n, cout, h, w] view whose images are contiguous (e.g. `new_features[:, sp]`).
    """
    p = prepare(chain)
    n, ca, h, w = xa.shape
    xa, a_img = _image_view(xa)
    cb, b_img, bcast, xb_t = 0, 0, 0, None
    if xb is not None:
        cb = xb.shape[1]
        if xb.shape[-2:] == (1, 1) and (h, w) != (1, 1):
            xb_t = xb.reshape(xb.shape[0], cb).float().contiguous()
            if xb_t.shape[0] == 1 and n > 1:
                xb_t = xb_t.expand(n, cb).contiguous()
            b_img, bcast = cb, 1
        else:
            xb_t, b_img = _image_view(xb)
    if ca + cb != p.cin:
        raise RuntimeError("conv1x1 chain expects %d input channels, got %d" % (p.cin, ca + cb))
    if out is None:
        out = xa.new_empty(n, p.cout, h, w)
    if out.dtype != th.float32 or out.stride()[1:] != (h * w, w, 1):
        raise RuntimeError("conv1x1 chain: `out` must be float32 with contiguous images")
    y_img = out.stride(0) if n > 1 else p.cout * h * w
    lib = _lib.load()
    with th.cuda.device(xa.device):
        rc = lib.sbmc_conv1x1_chain_f32(
            xa.data_ptr(), ca, a_img, xb_t.data_ptr() if xb_t is not None else None, cb, b_img,
            bcast, p.w1.data_ptr(), p.b1.data_ptr(), p.w2.data_ptr(), p.b2.data_ptr(),
            p.w3.data_ptr(), p.b3.data_ptr(), p.k1p, p.cout, p.n3p, p.act,
            out.data_ptr(), y_img, n, h * w, th.cuda.current_stream(xa.device).cuda_stream)
    _lib.check(rc, "conv1x1_chain")
    return out


# -- bf16 channels-innermost pipeline ---------------------------------------------
def to_nhwc_bf16(x, channels=_HID):
    """[..., c, h, w] fp32 -> [..., h*w, channels] bf16 (zero-padded channels)."""
    c, h, w = x.shape[-3:]
    lead = x.shape[:-3]
    if x.is_cuda and x.dtype == th.float32 and channels % 8 == 0 and c <= channels:
        x = x.contiguous()
        out = th.empty(lead + (h * w, channels), device=x.device, dtype=th.bfloat16)
        n = 1
        for d in lead:
            n *= d
        lib = _lib.load()
        with th.cuda.device(x.device):
            rc = lib.sbmc_nchw_to_nhwc_bf16(
                x.data_ptr(), c * h * w, out.data_ptr(), h * w * channels, n, c, h * w,
                channels, th.cuda.current_stream(x.device).cuda_stream)
        _lib.check(rc, "nchw_to_nhwc")
        return out
    out = x.new_zeros(lead + (h * w, channels), dtype=th.bfloat16)
    out[..., :c] = x.reshape(lead + (c, h * w)).transpose(-1, -2)
    return out


class _PreparedNhwc(object):
    __slots__ = ("w1", "w1_gf", "w2", "w3", "b1", "b2", "b3", "cout", "n3p", "act", "versions",
                 "key")


def _prepare_nhwc(chain, ca, cb, ngf):
    """Weights for the NHWC kernel: input channel order of the chain is
    [ca real channels of xa | ngf broadcast features | cb channels of xb];
    xa is stored padded to 128 channels."""
    key = (ca, cb, ngf)
    cached = getattr(chain, "_sbmc_b200_prepared_nhwc", None)
    ver = _versions(chain)
    if cached is not None and cached.versions == ver and cached.key == key:
        return cached
    c1, c2, c3 = _convs(chain)
    if ca + ngf + cb != c1.in_channels or ca > _HID or cb not in (0, _HID):
        raise RuntimeError("conv1x1 chain: %d + %d + %d input channels do not match the "
                           "chain's %d" % (ca, ngf, cb, c1.in_channels))
    dev = c1.bias.device
    w = _effective_weight(c1)
    p = _PreparedNhwc()
    w1 = th.zeros(_HID, _HID + cb, device=dev)
    w1[:, :ca] = w[:, :ca]
    if cb:
        w1[:, _HID:] = w[:, ca + ngf:]
    p.w1 = w1.to(th.bfloat16).contiguous()
    p.w1_gf = w[:, ca:ca + ngf].contiguous()          # fp32, folded into the bias per image
    p.w2 = _effective_weight(c2).to(th.bfloat16).contiguous()
    p.cout = c3.out_channels
    p.n3p = (p.cout + 15) // 16 * 16
    w3 = th.zeros(p.n3p, _HID, device=dev)
    w3[:p.cout] = _effective_weight(c3)
    p.w3 = w3.to(th.bfloat16).contiguous()
    p.b1 = c1.bias.detach().float().contiguous()
    p.b2 = c2.bias.detach().float().contiguous()
    b3 = th.zeros(p.n3p, device=dev)
    b3[:p.cout] = c3.bias.detach().float()
    p.b3 = b3
    p.act = 1 if isinstance(chain.layer_0.layer[1], th.nn.LeakyReLU) else 0
    p.versions, p.key = ver, key
    object.__setattr__(chain, "_sbmc_b200_prepared_nhwc", p)
    return p


def _nhwc_view(x):
    """x bf16 [n, hw, 128] with contiguous images -> (tensor, image stride)."""
    if x.dtype != th.bfloat16 or x.dim() != 3 or x.shape[-1] != _HID:
        raise RuntimeError("expected a bf16 [n, pixels, 128] tensor, got %s %s"
                           % (x.dtype, tuple(x.shape)))
    if x.stride()[1:] != (_HID, 1):
        x = x.contiguous()
    return x, (x.stride(0) if x.shape[0] > 1 else x.shape[1] * _HID)


def chain_forward_nhwc(chain, xa, ca, xb=None, gf=None, out=None, nhwc_out=True):
    """The chain on bf16 channels-innermost activations.

    xa bf16 [n, hw, 128] holding `ca` real channels; gf [n, ngf] fp32 features
    broadcast over the pixels (enter through a per-image first-layer bias); xb bf16
    [n, hw, 128] or None.  Output: bf16 [n, hw, 128] (nhwc_out, needs cout == 128)
    or fp32 [n, cout, hw]."""
    n, hw, _ = xa.shape
    ngf = 0 if gf is None else gf.shape[1]
    p = _prepare_nhwc(chain, ca, 0 if xb is None else _HID, ngf)
    xa, a_img = _nhwc_view(xa)
    xb_t, b_img = (None, 0) if xb is None else _nhwc_view(xb)
    if gf is not None:
        b1 = (p.b1.unsqueeze(0) + gf.float().reshape(n, ngf) @ p.w1_gf.t()).contiguous()
        b1_img = _HID
    else:
        b1, b1_img = p.b1, 0
    if out is None:
        out = (xa.new_empty(n, hw, _HID) if nhwc_out
               else th.empty(n, p.cout, hw, device=xa.device, dtype=th.float32))
    if nhwc_out:
        if p.cout != _HID or out.dtype != th.bfloat16 or out.stride()[1:] != (_HID, 1):
            raise RuntimeError("conv1x1 chain: bad bf16 NHWC output tensor")
        y_img = out.stride(0) if n > 1 else hw * _HID
    else:
        if out.dtype != th.float32 or out.stride()[1:] != (hw, 1):
            raise RuntimeError("conv1x1 chain: bad fp32 NCHW output tensor")
        y_img = out.stride(0) if n > 1 else p.cout * hw
    lib = _lib.load()
    with th.cuda.device(xa.device):
        rc = lib.sbmc_conv1x1_chain_nhwc_bf16(
            xa.data_ptr(), a_img, xb_t.data_ptr() if xb_t is not None else None, b_img,
            p.w1.data_ptr(), b1.data_ptr(), b1_img, p.w2.data_ptr(), p.b2.data_ptr(),
            p.w3.data_ptr(), p.b3.data_ptr(), p.cout, p.n3p, p.act, out.data_ptr(), y_img,
            1 if nhwc_out else 0, n, hw, th.cuda.current_stream(xa.device).cuda_stream)
    _lib.check(rc, "conv1x1_chain_nhwc")
    return out


def chain_samples_nhwc(chain, feats, ca, prop=None, gf=None, out=None, want_mean=False,
                       mean_dtype=th.bfloat16, regress=False, sample0=0, nsamples=None):
    """The chain for all samples of every pixel in ONE pipelined launch
    (csrc/chain_v3.cu; reference: the per-sample loops of sbmc/models.py:143-181 and
    :195-199).

    feats bf16 [n, spp, hw, 128] holding `ca` real channels per sample; prop bf16
    [n, hw, 128] shared by the samples of a pixel (or None); gf [n, ngf] fp32 global
    features (enter through a per-image first-layer bias).  Samples
    [sample0, sample0 + nsamples) are processed.

    regress=False (embedding chains): returns (out, mean) with out bf16
    [n, spp, hw, 128] (only the processed samples are written) and, if want_mean,
    mean [n, hw, 128] = the mean of the chain output over the processed samples,
    taken on the fp32 accumulators (models.py:181), else None.
    regress=True (kernel regressor): returns fp32 logits [n, nsamples, cout, hw]."""
    n, spp, hw, _ = feats.shape
    if feats.dtype != th.bfloat16 or feats.shape[-1] != _HID or feats.stride()[2:] != (_HID, 1):
        raise RuntimeError("expected bf16 [n, spp, pixels, 128] features with contiguous samples")
    nsamples = spp - sample0 if nsamples is None else nsamples
    ngf = 0 if gf is None else gf.shape[1]
    p = _prepare_nhwc(chain, ca, 0 if prop is None else _HID, ngf)
    prop_t, p_img = (None, 0) if prop is None else _nhwc_view(prop)
    if gf is not None:
        b1 = (p.b1.unsqueeze(0) + gf.float().reshape(n, ngf) @ p.w1_gf.t()).contiguous()
        b1_img = _HID
    else:
        b1, b1_img = p.b1, 0
    f_img = feats.stride(0) if n > 1 else spp * hw * _HID
    f_smp = feats.stride(1) if spp > 1 else hw * _HID
    mean = None
    if regress:
        if want_mean:
            raise RuntimeError("the regressor has no sample mean")
        if out is None:
            out = th.empty(n, nsamples, p.cout, hw, device=feats.device, dtype=th.float32)
        if out.dtype != th.float32 or out.shape != (n, nsamples, p.cout, hw) \
                or not out.is_contiguous():
            raise RuntimeError("conv1x1 chain: bad fp32 logits tensor")
        o_img, o_smp = nsamples * p.cout * hw, p.cout * hw
        # the kernel indexes samples absolutely: shift the base by sample0 samples
        o_ptr = out.data_ptr() - sample0 * o_smp * 4
    else:
        if p.cout != _HID:
            raise RuntimeError("embedding chains must have 128 output channels")
        if out is None:
            out = feats.new_empty(n, spp, hw, _HID)
        if out.dtype != th.bfloat16 or out.shape != (n, spp, hw, _HID) \
                or out.stride()[2:] != (_HID, 1):
            raise RuntimeError("conv1x1 chain: bad bf16 output tensor")
        o_img = out.stride(0) if n > 1 else spp * hw * _HID
        o_smp = out.stride(1) if spp > 1 else hw * _HID
        o_ptr = out.data_ptr()
        if want_mean:
            mean = th.empty(n, hw, _HID, device=feats.device, dtype=mean_dtype)
    lib = _lib.load()
    with th.cuda.device(feats.device):
        rc = lib.sbmc_chain_samples_nhwc_bf16(
            feats.data_ptr(), f_img, f_smp, spp,
            prop_t.data_ptr() if prop_t is not None else None, p_img,
            p.w1.data_ptr(), b1.data_ptr(), b1_img, p.w2.data_ptr(), p.b2.data_ptr(),
            p.w3.data_ptr(), p.b3.data_ptr(), p.cout, p.n3p, p.act, 1 if regress else 0,
            o_ptr, o_img, o_smp, mean.data_ptr() if mean is not None else None, hw * _HID,
            1 if (mean is not None and mean.dtype == th.float32) else 0, n, sample0, nsamples,
            hw, th.cuda.current_stream(feats.device).cuda_stream)
    _lib.check(rc, "chain_samples")
    return out if regress else (out, mean)
