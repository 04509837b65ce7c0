"""Fused 3-layer 1x1 ConvChain on the tcgen05 tensor cores (inference).

Binds ``sbmc_conv1x1_chain_f32`` (include/sbmc_b200.h, csrc/conv1x1.cu) to the
per-sample networks of the reference model (sbmc/models.py:86-102): the
``embedding_XX`` chains and the ``kernel_regressor`` are 1x1 convolutions, i.e. a
3-layer MLP per pixel.  bf16 operands, fp32 accumulation, fp32 output.
"""
import torch as th

from . import _lib

__all__ = ["supports", "prepare", "chain_forward"]

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
