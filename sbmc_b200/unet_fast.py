"""bf16 channels_last inference path of the U-net (`modules.Autoencoder`,
sbmc/modules.py:195-320).  The 3x3 convolutions run on this repo's tcgen05
implicit-GEMM kernel with bias + activation fused into its epilogue
(csrc/conv3x3.cu, `own_convs=True`, the default) on cached bf16 weights with the
weight normalization folded in; `own_convs=False` keeps the cuDNN library
convolutions + the separate bias / activation pass (the comparison baseline).
Bilinear upsample + skip concatenation is one pass (csrc/unet_ops.cu) instead of
eager PyTorch's interpolate and cat kernels.
"""
import torch as th
import torch.nn.functional as F

from . import _lib
from . import conv3x3 as _conv3x3

__all__ = ["supports", "supports_training", "autoencoder_forward", "autoencoder_forward_train"]

_ACT = {th.nn.ReLU: 1, th.nn.LeakyReLU: 2}


def _chain_layers(chain):
    """[(conv, act_code)] of a ConvChain without normalization layers, or None."""
    layers = []
    kids = list(chain.named_children())
    for name, m in kids:
        if name.startswith("layer_"):
            seq = m.layer
            if len(seq) != 2 or type(seq[1]) not in _ACT:
                return None
            if isinstance(seq[1], th.nn.LeakyReLU) and abs(seq[1].negative_slope - 0.01) > 1e-12:
                return None
            layers.append((seq[0], _ACT[type(seq[1])]))
        elif name == "prediction":
            layers.append((m, 0))
        elif name == "output_activation":
            if type(m) not in _ACT or not layers:
                return None
            if isinstance(m, th.nn.LeakyReLU) and abs(m.negative_slope - 0.01) > 1e-12:
                return None
            layers[-1] = (layers[-1][0], _ACT[type(m)])
        else:
            return None
    for conv, _ in layers:
        if not isinstance(conv, th.nn.Conv2d) or conv.bias is None or conv.groups != 1 \
                or conv.out_channels % 8 or conv.stride != (1, 1) or conv.dilation != (1, 1):
            return None
    return layers


def _levels(level):
    while level is not None:
        yield level
        level = None if level.is_last else level.next_level


def supports(autoencoder):
    try:
        for lvl in _levels(autoencoder.net):
            if _chain_layers(lvl.left) is None:
                return False
            if not lvl.is_last:
                if _chain_layers(lvl.right) is None or not isinstance(lvl.downsample, th.nn.MaxPool2d):
                    return False
        return True
    except AttributeError:
        return False


def _prepared(conv):
    """(bf16 channels_last weight with weight-norm folded, fp32 bias), cached."""
    ver = tuple((p.data_ptr(), p._version) for p in conv.parameters())
    cached = getattr(conv, "_sbmc_b200_fast", None)
    if cached is not None and cached[0] == ver:
        return cached[1], cached[2]
    if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
        w = th._weight_norm(conv.weight_v, conv.weight_g, 0)
    else:
        w = conv.weight
    w = w.detach().to(th.bfloat16).contiguous(memory_format=th.channels_last)
    b = conv.bias.detach().float().contiguous()
    object.__setattr__(conv, "_sbmc_b200_fast", (ver, w, b))
    return w, b


def _prepared9(conv):
    """(bf16 [9, cout, cin] tap-major weight with weight-norm folded, fp32 bias), cached."""
    ver = tuple((p.data_ptr(), p._version) for p in conv.parameters())
    cached = getattr(conv, "_sbmc_b200_fast9", None)
    if cached is not None and cached[0] == ver:
        return cached[1], cached[2]
    if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
        w = th._weight_norm(conv.weight_v, conv.weight_g, 0)
    else:
        w = conv.weight
    w9 = _conv3x3.prepare_weight(w)
    b = conv.bias.detach().float().contiguous()
    object.__setattr__(conv, "_sbmc_b200_fast9", (ver, w9, b))
    return w9, b


def _bias_act_(y, bias, act):
    n, c, h, w = y.shape
    lib = _lib.load()
    with th.cuda.device(y.device):
        rc = lib.sbmc_bias_act_nhwc_bf16(y.data_ptr(), bias.data_ptr(), n * h * w, c, act,
                                         th.cuda.current_stream(y.device).cuda_stream)
    _lib.check(rc, "bias_act")
    return y


def _chain(chain, x, own_convs=True):
    for conv, act in _chain_layers(chain):
        if own_convs and _conv3x3.supports_conv(conv):
            # NCHW-shaped channels_last tensor <-> contiguous [n, h, w, c] view: no copies
            w9, b = _prepared9(conv)
            if not x.is_contiguous(memory_format=th.channels_last):
                x = x.contiguous(memory_format=th.channels_last)
            y = _conv3x3.conv3x3_nhwc(x.permute(0, 2, 3, 1), w9, b, act)
            x = y.permute(0, 3, 1, 2)
            continue
        w, b = _prepared(conv)
        x = F.conv2d(x, w, None, conv.stride, conv.padding)
        if not x.is_contiguous(memory_format=th.channels_last):
            x = x.contiguous(memory_format=th.channels_last)
        _bias_act_(x, b, act)
    return x


def _maxpool2x2(x):
    """nn.MaxPool2d(2, 2) on an NCHW-shaped bf16 channels_last tensor (own kernel)."""
    n, c, h, w = x.shape
    if not x.is_contiguous(memory_format=th.channels_last):
        x = x.contiguous(memory_format=th.channels_last)
    y = th.empty((n, c, h // 2, w // 2), device=x.device, dtype=th.bfloat16,
                 memory_format=th.channels_last)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_maxpool2x2_nhwc_bf16(x.data_ptr(), y.data_ptr(), n, h, w, c,
                                           th.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "maxpool2x2")
    return y


def _is_pool2x2(m):
    def pair(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    return (isinstance(m, th.nn.MaxPool2d) and pair(m.kernel_size) == (2, 2)
            and pair(m.stride) == (2, 2) and pair(m.padding) == (0, 0)
            and pair(m.dilation) == (1, 1) and not m.ceil_mode)


def _level(level, x, own_convs):
    from .modules import _upsample_concat
    left = _chain(level.left, x, own_convs)
    if level.is_last:
        return left
    if own_convs and _is_pool2x2(level.downsample) and left.shape[1] % 8 == 0 \
            and left.shape[2] >= 2 and left.shape[3] >= 2:
        pooled = _maxpool2x2(left)
    else:
        pooled = level.downsample(left)
    coarse = _level(level.next_level, pooled, own_convs)
    return _chain(level.right, _upsample_concat(coarse, left), own_convs)


def autoencoder_forward(autoencoder, x, own_convs=True):
    """x: [n, c, h, w] (any float dtype; converted to bf16 channels_last) ->
    bf16 channels_last [n, c_out, h, w].  Inference only."""
    x = x.to(th.bfloat16).contiguous(memory_format=th.channels_last)
    with th.no_grad():
        return _level(autoencoder.net, x, own_convs)


# -- opt-in mixed-precision TRAINING path ------------------------------------------------
def supports_training(autoencoder):
    """Every convolution of the U-net can run forward and data-gradient on conv3x3.cu."""
    if not supports(autoencoder):
        return False
    for lvl in _levels(autoencoder.net):
        chains = [lvl.left] + ([] if lvl.is_last else [lvl.right])
        for chain in chains:
            for conv, _ in _chain_layers(chain):
                if not _conv3x3.supports_conv_training(conv):
                    return False
        if not lvl.is_last and not isinstance(lvl.downsample, th.nn.MaxPool2d):
            return False
    return True


def _weight9(conv):
    """bf16 [9, cout, cin] weight as a DIFFERENTIABLE function of the module's
    parameters (weight normalization included), recomputed every step."""
    if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
        w = th._weight_norm(conv.weight_v, conv.weight_g, 0)
    else:
        w = conv.weight
    cout, cin = w.shape[:2]
    return w.permute(2, 3, 0, 1).reshape(9, cout, cin).to(th.bfloat16).contiguous()


def _chain_train(chain, x):
    """x bf16 [n, h, w, c] contiguous -> same layout."""
    for conv, act in _chain_layers(chain):
        x = _conv3x3.Conv3x3BiasAct.apply(x.contiguous(), _weight9(conv), conv.bias.float(), act)
    return x


def _level_train(level, x):
    left = _chain_train(level.left, x)
    if level.is_last:
        return left
    # pooling / upsampling / concatenation: torch ops (autograd), NCHW-shaped views of the
    # channels-innermost tensors
    pooled = level.downsample(left.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    coarse = _level_train(level.next_level, pooled.contiguous())
    up = F.interpolate(coarse.permute(0, 3, 1, 2), size=left.shape[1:3], mode="bilinear",
                       align_corners=False).permute(0, 2, 3, 1)
    return _chain_train(level.right, th.cat([up, left], 3))


def autoencoder_forward_train(autoencoder, x):
    """Differentiable bf16 forward of the U-net: x [n, c, h, w] (fp32 or bf16) ->
    [n, c_out, h, w] in x's dtype.  Convolutions: `Conv3x3BiasAct` (forward and data
    gradient on this repo's tcgen05 kernel, weight gradient on cuDNN bf16)."""
    xn = x.permute(0, 2, 3, 1).to(th.bfloat16).contiguous()
    y = _level_train(autoencoder.net, xn)
    return y.permute(0, 3, 1, 2).to(x.dtype)
