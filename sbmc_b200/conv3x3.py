"""3x3 convolution on the tcgen05 tensor cores (inference): binds
``sbmc_conv3x3_nhwc_bf16`` (include/sbmc_b200.h, csrc/conv3x3.cu) to the U-net
convolutions of the reference (sbmc/modules.py:248-320: every conv of
``Autoencoder`` is 3x3, stride 1, padding 1, followed by ReLU / LeakyReLU).

bf16 channels-innermost activations, fp32 accumulation, bias + activation fused
into the kernel's epilogue.
"""
import torch as th

from . import _lib

__all__ = ["supports_conv", "supports_conv_training", "prepare_weight", "conv3x3_nhwc",
           "Conv3x3BiasAct"]


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


# -- training: the same kernel as forward AND as the data gradient --------------------
def supports_conv_training(conv):
    """Forward needs cin % 64 == 0 and cout % 128 == 0; the data gradient is the same
    kernel with the roles swapped, so it needs cout % 64 == 0 and cin % 128 == 0."""
    return supports_conv(conv) and conv.in_channels % 128 == 0


def _flip_transpose(w9):
    """Weights of the data-gradient convolution: dX = conv3x3(dPre, W') with
    W'[3 dy + dx][ci][co] = W[3 (2 - dy) + (2 - dx)][co][ci]."""
    return w9.flip(0).transpose(1, 2).contiguous()


class Conv3x3BiasAct(th.autograd.Function):
    """y = act(conv3x3(x, w9) + bias) on bf16 channels-innermost tensors with a
    backward pass: an opt-in mixed-precision training path for the U-net
    (sbmc/modules.py:248-320; the reference trains in fp32).

    forward   : csrc/conv3x3.cu (tcgen05 implicit GEMM, bias + activation fused)
    d_input   : the SAME kernel on (dPre, flipped / transposed weights)
    d_weights : cuDNN's bf16 weight-gradient kernel (library call; a tcgen05 split-K
                weight-gradient kernel is DESIGN.md section 9 item 1)
    d_bias    : fp32 sum of dPre over the pixels
    dPre = dY * act'(y): ReLU / LeakyReLU(0.01) derivatives are read off the sign of y."""

    @staticmethod
    def forward(ctx, x, w9, bias, act):
        y = conv3x3_nhwc(x, w9, bias, act)
        ctx.save_for_backward(x, w9, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w9, y = ctx.saved_tensors
        dpre = dy.contiguous()
        if ctx.act == 1:
            dpre = th.where(y > 0, dpre, th.zeros_like(dpre))
        elif ctx.act == 2:
            dpre = th.where(y > 0, dpre, dpre * 0.01)
        n, h, w, cin = x.shape
        cout = w9.shape[1]
        dx = dw9 = db = None
        if ctx.needs_input_grad[0]:
            zero = th.zeros(cin, device=x.device, dtype=th.float32)
            dx = conv3x3_nhwc(dpre, _flip_transpose(w9), zero, 0)
        if ctx.needs_input_grad[1]:
            dw = th.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (cout, cin, 3, 3),
                                          dpre.permute(0, 3, 1, 2), padding=1)
            dw9 = dw.permute(2, 3, 0, 1).reshape(9, cout, cin).to(w9.dtype)
        if ctx.needs_input_grad[2]:
            db = dpre.float().sum((0, 1, 2))
        return dx, dw9, db, None
