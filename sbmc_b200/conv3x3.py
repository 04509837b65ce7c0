"""3x3 convolution on the tcgen05 tensor cores (inference): binds
``sbmc_conv3x3_nhwc_bf16`` (include/sbmc_b200.h, csrc/conv3x3.cu) to the U-net
convolutions of the reference (sbmc/modules.py:248-320: every conv of
``Autoencoder`` is 3x3, stride 1, padding 1, followed by ReLU / LeakyReLU).

bf16 channels-innermost activations, fp32 accumulation, bias + activation fused
into the kernel's epilogue.
"""
import torch as th

from . import _lib

__all__ = ["supports_conv", "prepare_weight", "conv3x3_nhwc"]


def supports_conv(conv):
    """Whether an nn.Conv2d has the shape the kernel serves."""
    return (isinstance(conv, th.nn.Conv2d) and conv.kernel_size == (3, 3)
            and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1)
            and conv.groups == 1 and conv.bias is not None
            and getattr(conv, "padding_mode", "zeros") == "zeros"
            and conv.in_channels % 64 == 0 and conv.out_channels % 128 == 0)


def prepare_weight(w):
    """[cout, cin, 3, 3] (any float dtype) -> bf16 [9, cout, cin], tap = 3 * dy + dx."""
    cout, cin = w.shape[:2]
    return w.detach().permute(2, 3, 0, 1).reshape(9, cout, cin).to(th.bfloat16).contiguous()


def conv3x3_nhwc(x, w9, bias, act=0, out=None):
    """x bf16 [n, h, w, cin] contiguous; w9 bf16 [9, cout, cin]; bias fp32 [cout].
    Returns bf16 [n, h, w, cout] = act(conv3x3(x) + bias); act: 0 none, 1 ReLU,
    2 LeakyReLU(0.01)."""
    n, h, w, cin = x.shape
    cout = w9.shape[1]
    if x.dtype != th.bfloat16 or not x.is_contiguous() or w9.dtype != th.bfloat16 \
            or w9.shape != (9, cout, cin) or not w9.is_contiguous():
        raise RuntimeError("conv3x3: expected contiguous bf16 [n,h,w,cin] and [9,cout,cin]")
    if bias.dtype != th.float32 or bias.numel() != cout:
        raise RuntimeError("conv3x3: bias must be float32 [cout]")
    if out is None:
        out = th.empty(n, h, w, cout, device=x.device, dtype=th.bfloat16)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_conv3x3_nhwc_bf16(x.data_ptr(), w9.data_ptr(), bias.data_ptr(),
                                        out.data_ptr(), n, h, w, cin, cout, act,
                                        th.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "conv3x3")
    return out
