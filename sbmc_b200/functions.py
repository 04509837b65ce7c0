"""The two autograd entry points of the kernel-splatting path.

`Scatter2Gather` and `KernelWeighting` keep the public contract of the
reference's ``sbmc/functions.py`` (:39-71 and :74-115) -- class names,
``apply`` signatures, what is saved for backward, and the tuple each returns --
so ``sbmc/modules.py``-style callers work unchanged.  Underneath, every call is
routed through ``sbmc_b200.halide_ops`` (the drop-in for the reference's native
module) to the sm_100a kernels of libsbmc_b200: CUDA tensors run on the current
stream of their device, host tensors are streamed through the GPU.
"""
import torch as th

from . import halide_ops as ops

__all__ = ["Scatter2Gather", "KernelWeighting"]


def _native(op, *tensors):
    """Pick the ``<op>_{cuda,cpu}_float32`` entry point: the CUDA flavour as soon
    as one operand lives on a GPU (same rule as functions.py:30-36)."""
    flavour = "cuda" if any(t.is_cuda for t in tensors) else "cpu"
    return getattr(ops, "%s_%s_float32" % (op, flavour))


def _transpose_kernels(kernels):
    kernels = kernels.contiguous()
    out = th.empty_like(kernels, memory_format=th.contiguous_format)
    _native("scatter2gather", kernels)(kernels, out)
    return out


class Scatter2Gather(th.autograd.Function):
    """Turns per-sample splatting kernels into per-pixel gathering kernels.

    Input and output are ``[bs, k_h, k_w, h, w]``.  The weight a sample at
    ``(y, x)`` sends to offset ``(dy, dx)`` becomes the weight the pixel at
    ``(y + dy, x + dx)`` applies to offset ``(-dy, -dx)``; taps whose partner lies
    outside the image are zero.  The permutation is its own adjoint, so the
    backward pass applies it to the incoming gradient.
    """

    @staticmethod
    def forward(ctx, data):
        if data.dim() != 5:
            raise AssertionError("data should be 5d")
        return _transpose_kernels(data)

    @staticmethod
    def backward(ctx, d_output):
        return _transpose_kernels(d_output)


class KernelWeighting(th.autograd.Function):
    """Per-pixel weighted sum of a neighbourhood.

    ``data`` is ``[bs, c, h, w]``, ``weights`` ``[bs, k_h, k_w, h, w]`` (gather
    kernels, every channel filtered alike).  Returns ``(output, sum_w)`` with
    ``output[b, c, y, x] = sum_{dy, dx} weights[b, dy, dx, y, x] *
    data[b, c, y + dy - (k_h-1)//2, x + dx - (k_w-1)//2]`` (zero outside the
    image) and ``sum_w[b, y, x]`` the sum of all ``k_h * k_w`` weights of the pixel.
    """

    @staticmethod
    def forward(ctx, data, weights):
        bs, _, h, w = data.shape
        data, weights = data.contiguous(), weights.contiguous()
        output = th.empty_like(data, memory_format=th.contiguous_format)
        sum_w = data.new_empty((bs, h, w))
        _native("kernel_weighting", data, weights)(data, weights, output, sum_w)
        ctx.save_for_backward(data, weights, sum_w)
        return output, sum_w

    @staticmethod
    def backward(ctx, d_output, d_sum_w):
        data, weights, sum_w = ctx.saved_tensors
        d_output, d_sum_w = d_output.contiguous(), d_sum_w.contiguous()
        grads = (th.empty_like(data), th.empty_like(weights))
        _native("kernel_weighting_grad", d_output, d_sum_w)(
            data, weights, sum_w, d_output, d_sum_w, *grads)
        return grads
