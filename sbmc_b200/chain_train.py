"""Opt-in mixed-precision TRAINING path of the per-sample 1x1 ConvChains
(``embedding_XX`` / ``kernel_regressor``, sbmc/models.py:86-102; sbmc/modules.py:34-125).

The reference trains in fp32 through cuDNN.  Here a depth-3 chain runs layer by layer
on bf16 channels-innermost rows [pixels, channels]:

  forward        three launches of the tcgen05 GEMM layer (csrc/linear.cu), the two hidden
                 activations are kept for the backward pass;
  data gradient  three launches of the SAME kernel on the transposed weights, the
                 activation derivative applied between them (read off the sign of the
                 saved activations);
  weight / bias gradients   library GEMMs (`torch.matmul` in bf16, fp32 accumulate) and
                 fp32 sums -- a tcgen05 pixel-reduction GEMM is DESIGN.md section 9 item 1.

Gradients carry bf16 rounding (about 1e-2 relative); `Multisteps.bf16_train` switches it
on, the default training path is unchanged.
"""
import torch as th

from . import _lib

__all__ = ["linear_nhwc", "ChainFn", "chain_weights", "supports_training"]

_HID = 128


def linear_nhwc(x, w, bias=None, act=0, out_dtype=th.bfloat16):
    """x bf16 [pixels, cin]; w bf16 [cout, cin]; bias fp32 [cout] or None ->
    act(x @ w.T + bias) as bf16 or fp32 [pixels, cout] (csrc/linear.cu)."""
    p, cin = x.shape
    cout = w.shape[0]
    if x.dtype != th.bfloat16 or w.dtype != th.bfloat16 or not x.is_contiguous() \
            or not w.is_contiguous() or w.shape[1] != cin:
        raise RuntimeError("linear: expected contiguous bf16 [pixels, cin] and [cout, cin]")
    if bias is not None and (bias.dtype != th.float32 or bias.numel() != cout):
        raise RuntimeError("linear: bias must be float32 [cout]")
    y = th.empty(p, cout, device=x.device, dtype=out_dtype)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_linear_nhwc_bf16(
            x.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
            y.data_ptr(), p, cin, cout, act, 1 if out_dtype == th.float32 else 0,
            th.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "linear")
    return y


def _dact(h, g, act):
    """g * act'(pre-activation), the derivative read off the sign of the output h."""
    if act == 1:
        return th.where(h > 0, g, th.zeros_like(g))
    if act == 2:
        return th.where(h > 0, g, g * 0.01)
    return g


class ChainFn(th.autograd.Function):
    """y = W3 act(W2 act(W1 x + b1) + b2) + b3 on rows x [pixels, cin] (bf16).

    w1 [128, cin], w2 [128, 128], w3 [cout_p, 128] bf16 (cout_p a multiple of 128: pad
    with zero rows), biases fp32; act 1 ReLU / 2 LeakyReLU(0.01); out_f32: fp32 output
    (the kernel regressor's logits keep fp32 storage, SURVEY.md section 8a note 4)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, act, out_f32):
        h1 = linear_nhwc(x, w1, b1, act)
        h2 = linear_nhwc(h1, w2, b2, act)
        y = linear_nhwc(h2, w3, b3, 0, th.float32 if out_f32 else th.bfloat16)
        ctx.save_for_backward(x, w1, w2, w3, h1, h2)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, w3, h1, h2 = ctx.saved_tensors
        act = ctx.act
        dyb = dy.to(th.bfloat16).contiguous()
        need = ctx.needs_input_grad
        # data gradients: the same GEMM kernel on the transposed weights
        dh2 = _dact(h2, linear_nhwc(dyb, w3.t().contiguous()), act)
        dh1 = _dact(h1, linear_nhwc(dh2, w2.t().contiguous()), act)
        dx = linear_nhwc(dh1, w1.t().contiguous()) if need[0] else None
        # weight gradients: reductions over the pixels (library GEMMs)
        dw1 = th.matmul(dh1.t(), x) if need[1] else None
        dw2 = th.matmul(dh2.t(), h1) if need[3] else None
        dw3 = th.matmul(dyb.t(), h2) if need[5] else None
        db1 = dh1.float().sum(0) if need[2] else None
        db2 = dh2.float().sum(0) if need[4] else None
        db3 = dy.float().sum(0) if need[6] else None
        return dx, dw1, db1, dw2, db2, dw3, db3, None, None


def supports_training(chain):
    from . import conv1x1
    return conv1x1.supports(chain)


def chain_weights(chain, cin_pad):
    """(w1, b1, w2, b2, w3, b3, act, cout) of a depth-3 1x1 ConvChain as DIFFERENTIABLE
    functions of its parameters (weight normalization included): bf16 weights, w1's
    input channels zero-padded to `cin_pad`, w3 / b3 zero-padded to a multiple of 128
    output channels."""
    from .conv1x1 import _convs
    c1, c2, c3 = _convs(chain)

    def weight(conv):
        if hasattr(conv, "weight_g") and hasattr(conv, "weight_v"):
            w = th._weight_norm(conv.weight_v, conv.weight_g, 0)
        else:
            w = conv.weight
        return w.reshape(w.shape[0], w.shape[1])
    w1, w2, w3 = weight(c1), weight(c2), weight(c3)
    if w1.shape[1] < cin_pad:
        w1 = th.nn.functional.pad(w1, (0, cin_pad - w1.shape[1]))
    cout = w3.shape[0]
    cout_p = (cout + 127) // 128 * 128
    b3 = c3.bias.float()
    if cout_p != cout:
        w3 = th.nn.functional.pad(w3, (0, 0, 0, cout_p - cout))
        b3 = th.nn.functional.pad(b3, (0, cout_p - cout))
    act = 2 if isinstance(chain.layer_0.layer[1], th.nn.LeakyReLU) else 1
    return (w1.to(th.bfloat16).contiguous(), c1.bias.float(), w2.to(th.bfloat16).contiguous(),
            c2.bias.float(), w3.to(th.bfloat16).contiguous(), b3, act, cout)
