"""Modules of the kernel-splatting hot path, API-identical to the reference's
``sbmc/modules.py``: ``ConvChain`` (:34-192), ``Autoencoder`` (:195-320),
``KernelApply`` (:323-361), ``ProgressiveKernelApply`` (:364-473).

Constructor signatures, child-module names (so reference state dicts load:
``layer_i.layer.0.weight_g/_v``, ``prediction``, ``output_activation``,
``net.left / right / downsample / next_level``), error behaviour and return
tuples follow the reference.  The two custom ops underneath are the sm_100a
kernels of libsbmc_b200 (``sbmc_b200.functions``); ``ProgressiveKernelApply``
additionally has a fused CUDA path for inference (``sbmc_b200.splat``).
"""
import math

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from . import functions as funcs
from . import splat as _splat
from ._compat import get_logger

__all__ = ["ConvChain", "Autoencoder", "KernelApply", "ProgressiveKernelApply"]

LOG = get_logger(__name__)

_ACTIVATIONS = {"relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "tanh": nn.Tanh,
                "elu": nn.ELU}


def _init_conv(conv, kind):
    """Zero bias, Xavier-uniform weights with the gain of the following
    non-linearity (modules.py:88-94,183-188; elu / softplus use the relu gain)."""
    conv.bias.data.zero_()
    gain_of = "relu" if kind in ("elu", "softplus") else kind
    nn.init.xavier_uniform_(conv.weight.data, nn.init.calculate_gain(gain_of))


class ConvChain(nn.Module):
    """A stack of `depth` convolutions: `depth - 1` conv(+norm)+activation groups
    named ``layer_0 ...`` followed by a last conv named ``prediction`` and an
    optional ``output_activation``.

    Args:
        ninputs(int): number of input channels.
        noutputs(int): number of output channels.
        ksize(int): size of all the convolution kernels.
        width(int): number of channels per intermediate layer.
        depth(int): number of conv layers.
        stride(int): stride of the intermediate convolutions.
        pad(bool): if True keep the spatial size by zero padding (ksize // 2).
        normalize(bool): add a normalization layer after each intermediate conv.
        normalization_type(str): "batch" or "instance".
        output_type(str): linear, relu, leaky_relu, sigmoid, tanh, elu, softplus.
        activation(str): relu, leaky_relu, tanh or elu.
        weight_norm(bool): wrap the convolutions in weight normalization.
    """

    def __init__(self, ninputs, noutputs, ksize=3, width=64, depth=3, stride=1,
                 pad=True, normalize=False, normalization_type="batch",
                 output_type="linear", activation="relu", weight_norm=True):
        super(ConvChain, self).__init__()
        if depth <= 0:
            LOG.error("ConvChain should have non-negative depth.")
            raise ValueError("negative network depth.")
        padding = ksize // 2 if pad else 0

        n_in = ninputs
        for d in range(depth - 1):
            self.add_module("layer_{}".format(d), ConvChain._ConvBNRelu(
                n_in, ksize, width, normalize=normalize,
                normalization_type=normalization_type, padding=padding,
                stride=stride, activation=activation, weight_norm=weight_norm))
            n_in = width

        last = nn.Conv2d(n_in, noutputs, ksize, bias=True, padding=padding)
        if weight_norm:
            last = nn.utils.weight_norm(last)
        if output_type not in ("linear", "relu", "leaky_relu", "sigmoid", "tanh",
                               "elu", "softplus"):
            raise ValueError("Unknon output type '{}'".format(output_type))
        _init_conv(last, output_type)
        self.add_module("prediction", last)

        out_act = {"relu": lambda: nn.ReLU(inplace=True),
                   "leaky_relu": lambda: nn.LeakyReLU(inplace=True),
                   "sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "elu": nn.ELU,
                   "softplus": nn.Softplus}.get(output_type)
        if out_act is not None:
            self.add_module("output_activation", out_act())

    def forward(self, x):
        for m in self.children():
            x = m(x)
        return x

    class _ConvBNRelu(nn.Module):
        """Conv-(Norm)-Activation group; the layers live in ``self.layer``
        (an ``nn.Sequential``), so parameters are ``layer.0.*``."""

        def __init__(self, ninputs, ksize, noutputs, normalize=False,
                     normalization_type="batch", stride=1, padding=0,
                     activation="relu", weight_norm=True):
            super(ConvChain._ConvBNRelu, self).__init__()
            if activation not in _ACTIVATIONS:
                LOG.error("Incorrect activation %s", activation)
                raise ValueError("activation should be one of: "
                                 "relu, leaky_relu, tanh, elu")
            act_fn = _ACTIVATIONS[activation]
            if normalize:
                conv = nn.Conv2d(ninputs, noutputs, ksize, stride=stride,
                                 padding=padding, bias=False)
                if normalization_type == "batch":
                    nrm = nn.BatchNorm2d(noutputs)
                elif normalization_type == "instance":
                    # (the reference spells this nn.InstanceNorm2D, which does
                    # not exist: modules.py:166; affine so it has weight / bias)
                    nrm = nn.InstanceNorm2d(noutputs, affine=True)
                else:
                    LOG.error("Incorrect normalization %s", normalization_type)
                    raise ValueError("Unkown normalization type {}".format(
                        normalization_type))
                nrm.bias.data.zero_()
                nrm.weight.data.fill_(1.0)
                self.layer = nn.Sequential(conv, nrm, act_fn())
                nn.init.xavier_uniform_(conv.weight.data, nn.init.calculate_gain(
                    "relu" if activation == "elu" else activation))
            else:
                conv = nn.Conv2d(ninputs, noutputs, ksize, stride=stride,
                                 padding=padding)
                if weight_norm:
                    conv = nn.utils.weight_norm(conv)
                _init_conv(conv, activation)
                self.layer = nn.Sequential(conv, act_fn())

        def forward(self, x):
            return self.layer(x)


class Autoencoder(nn.Module):
    """A U-net style autoencoder, built coarsest level first (modules.py:221-243).

    Args:
        ninputs(int): number of input channels.
        noutputs(int): number of output channels.
        ksize(int): size of all the convolution kernels.
        width(int): number of channels per conv layer at the finest scale.
        num_levels(int): number of spatial scales.
        num_convs(int): number of conv layers per scale.
        max_width(int): max number of features per conv layer.
        increase_factor(float): channel multiplier from one scale to the next
            coarser one, up to `max_width`.
        normalize(bool), normalization_type(str): as in ConvChain.
        output_type(str), activation(str): as in ConvChain.
        pooling(str): "max", "average" or "conv".
    """

    def __init__(self, ninputs, noutputs, ksize=3, width=64, num_levels=3,
                 num_convs=2, max_width=512, increase_factor=1.0,
                 normalize=False, normalization_type="batch",
                 output_type="linear", activation="relu", pooling="max"):
        super(Autoencoder, self).__init__()

        def chans(lvl):
            return min(int(width * increase_factor ** lvl), max_width)

        level = None
        for lvl in reversed(range(num_levels)):
            finest, coarsest = lvl == 0, lvl == num_levels - 1
            level = Autoencoder._Level(
                ninputs if finest else chans(lvl - 1),
                noutputs if finest else chans(lvl),
                next_level=level, num_us=None if coarsest and not finest else chans(lvl + 1),
                ksize=ksize, width=chans(lvl), num_convs=num_convs,
                output_type=output_type if finest else activation,
                normalize=normalize, normalization_type=normalization_type,
                activation=activation, pooling=pooling)
        self.add_module("net", level)

    def forward(self, x):
        return self.net(x)

    class _Level(nn.Module):
        """One scale: ``left`` convs, then (unless coarsest) ``downsample`` ->
        ``next_level`` -> bilinear upsample -> concat skip -> ``right`` convs."""

        def __init__(self, num_inputs, num_outputs, next_level=None,
                     num_us=None, ksize=3, width=64, num_convs=2,
                     output_type="linear", normalize=True,
                     normalization_type="batch", pooling="max",
                     activation="relu"):
            super(Autoencoder._Level, self).__init__()
            self.is_last = next_level is None
            common = dict(ksize=ksize, width=width, depth=num_convs, stride=1,
                          pad=True, normalize=normalize,
                          normalization_type=normalization_type)
            if self.is_last:
                self.left = ConvChain(num_inputs, num_outputs,
                                      output_type=output_type, **common)
                return
            assert num_us is not None
            self.left = ConvChain(num_inputs, width, output_type=activation,
                                  activation=activation, **common)
            if pooling == "max":
                self.downsample = nn.MaxPool2d(2, 2)
            elif pooling == "average":
                self.downsample = nn.AvgPool2d(2, 2)
            elif pooling == "conv":
                self.downsample = nn.Conv2d(width, width, 2, stride=2)
            else:
                raise ValueError("unknown pooling'{}'".format(pooling))
            self.next_level = next_level
            self.right = ConvChain(num_us + width, num_outputs,
                                   output_type=output_type, **common)

        def forward(self, x):
            left = self.left(x)
            if self.is_last:
                return left
            coarse = self.next_level(self.downsample(left))
            if _fused_upsample_ok(coarse, left):
                return self.right(_upsample_concat(coarse, left))
            up = F.interpolate(coarse, size=left.shape[-2:], mode="bilinear",
                               align_corners=False)
            return self.right(th.cat([up, left], 1))


def _fused_upsample_ok(coarse, left):
    """bf16 channels_last CUDA tensors that need no gradient (the inference
    pipeline): upsample + concat run as one pass (csrc/unet_ops.cu)."""
    if th.is_grad_enabled() and (coarse.requires_grad or left.requires_grad):
        return False
    for t in (coarse, left):
        if not (t.is_cuda and t.dtype == th.bfloat16 and t.dim() == 4 and t.shape[1] % 8 == 0
                and t.is_contiguous(memory_format=th.channels_last)):
            return False
    return True


def _upsample_concat(coarse, left):
    from . import _lib
    n, cu, hl, wl = coarse.shape
    _, cs, h, w = left.shape
    out = th.empty((n, cu + cs, h, w), device=left.device, dtype=th.bfloat16,
                   memory_format=th.channels_last)
    lib = _lib.load()
    with th.cuda.device(left.device):
        rc = lib.sbmc_upsample_concat_nhwc_bf16(
            coarse.data_ptr(), left.data_ptr(), out.data_ptr(), n, hl, wl, h, w, cu, cs,
            th.cuda.current_stream(left.device).cuda_stream)
    _lib.check(rc, "upsample_concat")
    return out


def _ksize(k2):
    k = int(math.sqrt(k2) + 0.5)
    if k * k != k2:
        k = int(math.sqrt(k2))      # the reference truncates (modules.py:350)
    return k


class KernelApply(nn.Module):
    """Applies kernel-based averaging to the input tensor.

    Args:
        softmax(bool): softmax-normalize the kernels over the k*k taps of each
            output pixel.
        splat(bool): the kernels are splatting kernels; they are transposed to
            gather kernels (Scatter2Gather) before being applied.
    """

    def __init__(self, softmax=True, splat=True):
        super(KernelApply, self).__init__()
        self.softmax = softmax
        self.splat = splat

    def forward(self, data, kernels):
        """data [bs, chans, h, w], kernels [bs, k*k, h, w] ->
        (output [bs, chans, h, w], sum_w [bs, 1, h, w])."""
        bs, k2, h, w = kernels.shape
        k = _ksize(k2)
        if self.softmax and not self.splat and k * k == k2 and \
                _splat.fused_available(data, kernels):
            # inference: softmax over the taps + weighting in one pass over the
            # logits (the fused splat kernel in gather mode); the softmax weights
            # of a pixel sum to one by construction
            sum_r, sum_w, _ = _splat.progressive_splat_update(
                data.float(), kernels.float(), None, None, None, False)
            return sum_r / sum_w, th.ones_like(sum_w)
        # the custom ops are fp32 (reduced-precision convs may feed them under autocast)
        kernels = kernels.float().contiguous().view(bs, k, k, h, w)
        data = data.float()
        if self.splat:
            kernels = funcs.Scatter2Gather.apply(kernels)
        if self.softmax:
            kernels = F.softmax(kernels.view(bs, k * k, h, w), dim=1)
            kernels = kernels.view(bs, k, k, h, w)
        output, sum_w = funcs.KernelWeighting.apply(data, kernels)
        return output, sum_w.unsqueeze(1)


class ProgressiveKernelApply(nn.Module):
    """Applies progressive kernel-based averaging: an online softmax over the
    samples (running maximum `max_w`, running sums `sum_r`, `sum_w`); the
    normalized reconstruction is ``sum_r / sum_w`` (modules.py:364-473).

    Args:
        splat(bool): the kernels are splatting kernels (see KernelApply).
    """

    def __init__(self, splat=False):
        super(ProgressiveKernelApply, self).__init__()
        self.splat = splat
        # use the single-pass sm_100a kernel when no gradient is needed
        self.fused = True

    def forward(self, data, kernels, sum_r, sum_w, max_w):
        """First call: pass sum_r = sum_w = max_w = None.

        Args:
            data [bs, chans, h, w], kernels [bs, k*k, h, w] (logits; modified in
            place in gather mode, as in the reference), sum_r [bs, chans, h, w],
            sum_w / max_w [bs, 1, h, w] or None.
        Returns: updated (sum_r, sum_w, max_w).
        """
        bs, k2, h, w = kernels.shape
        k = _ksize(k2)
        if kernels.dtype != th.float32 or data.dtype != th.float32:
            kernels, data = kernels.float(), data.float()   # the custom ops are fp32
        first = sum_r is None
        if first and (sum_w is not None or max_w is not None):
            LOG.error("sum_r is None, this is the initialization step: "
                      "sum_w and max_w should be None as well.")
            raise RuntimeError("all of sum_r, sum_w, max_w should be none")

        if getattr(self, "fused", False) and k * k == k2 and \
                _splat.fused_available(data, kernels, sum_r, sum_w, max_w):
            return _splat.progressive_splat_update(data, kernels, sum_r, sum_w, max_w,
                                                   self.splat)
        if getattr(self, "fused", False) and self.splat and \
                _splat.fused_training_available(data, kernels, sum_r, sum_w, max_w):
            if first and (sum_w is not None or max_w is not None):
                raise RuntimeError("all of sum_r, sum_w, max_w should be none")
            return _splat.ProgressiveSplat.apply(data, kernels, sum_r, sum_w, max_w)

        kernels = kernels.view(bs, k, k, h, w)
        if self.splat:
            kernels = funcs.Scatter2Gather.apply(kernels)
        kmax = kernels.view(bs, k * k, h, w).max(1, keepdim=True)[0]

        if first:
            max_w = kmax
        else:
            new_max = th.max(kmax, max_w)
            scaler = th.exp(max_w - new_max)       # rescale the running sums
            sum_r = sum_r * scaler
            sum_w = sum_w * scaler
            max_w = new_max

        kernels.sub_(max_w.unsqueeze(1))            # for numerical stability
        kernels.exp_()
        new_r, new_w = funcs.KernelWeighting.apply(data.contiguous(),
                                                   kernels.contiguous())
        new_w = new_w.unsqueeze(1)
        if first:
            return new_r, new_w, max_w
        return sum_r + new_r, sum_w + new_w, max_w
