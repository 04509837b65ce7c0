"""Autograd functions of the kernel-splatting hot path, API-identical to the
reference's ``sbmc/functions.py`` (``Scatter2Gather`` :39-71, ``KernelWeighting``
:74-115): same class names, ``forward``/``backward`` signatures, tuple returns
and saved tensors.  The native module behind them is ``sbmc_b200.halide_ops``
(hand-written sm_100a CUDA behind a C ABI) instead of the Halide pipelines.
"""
import torch as th

from . import halide_ops as ops

__all__ = ["Scatter2Gather", "KernelWeighting"]


def _is_cuda(*args):
    """True if any of the arguments is on a CUDA device (functions.py:30-36)."""
    for arg in args:
        if arg.is_cuda:
            return True
    return False


class Scatter2Gather(th.autograd.Function):
    """Converts (transposes) scatter kernels into gather kernels.

    Kernel weights at (x, y) for offset (dx, dy) (i.e. scatter[., dy, dx, y,
    x]) are put at gather[., -dy, -dx, y+dy, x+dx].

    Args:
      data(th.Tensor)[bs, k_h, k_w, h, w]: scatter kernel weights.

    Returns:
      (th.Tensor)[bs, k_h, k_w, h, w]: gather kernel weights.
    """
    @staticmethod
    def forward(ctx, data):
        assert len(data.shape) == 5, "data should be 5d"
        output = th.empty_like(data, memory_format=th.contiguous_format)
        if _is_cuda(data):
            ops.scatter2gather_cuda_float32(data, output)
        else:
            ops.scatter2gather_cpu_float32(data, output)
        return output

    @staticmethod
    def backward(ctx, d_output):
        # The op is its own adjoint (functions.py:63-71).
        d_output = d_output.contiguous()
        d_data = th.empty_like(d_output)
        if _is_cuda(d_output):
            ops.scatter2gather_cuda_float32(d_output, d_data)
        else:
            ops.scatter2gather_cpu_float32(d_output, d_data)
        return d_data


class KernelWeighting(th.autograd.Function):
    """Locally-weighted average of the input values using kernel weights.

    Args:
      data(th.Tensor)[bs, c, h, w]: input values to be locally averaged.
      weights(th.Tensor)[bs, k_h, k_w, h, w]: kernel weights. k_h, k_w are
          the kernel's dimensions. Channels are filtered independently.

    Returns:
      output(th.Tensor)[bs, c, h, w]: weighted average of data using weights.
          output[., c, y, x] = sum_{dx, dy} weights[., dy, dx, x, y]*data[., c,
          y+dy, x+dx].
      sum_w(th.Tensor)[bs, h, w]: sum of weights per pixel
    """
    @staticmethod
    def forward(ctx, data, weights):
        bs, c, h, w = data.shape
        output = th.empty_like(data, memory_format=th.contiguous_format)
        sum_w = data.new_empty((bs, h, w))
        if _is_cuda(data, weights):
            ops.kernel_weighting_cuda_float32(data, weights, output, sum_w)
        else:
            ops.kernel_weighting_cpu_float32(data, weights, output, sum_w)
        ctx.save_for_backward(data, weights, sum_w)
        return output, sum_w

    @staticmethod
    def backward(ctx, d_output, d_sum_w):
        data, weights, sum_w = ctx.saved_tensors
        d_output = d_output.contiguous()
        d_sum_w = d_sum_w.contiguous()
        d_data = th.empty_like(data)
        d_weights = th.empty_like(weights)
        if _is_cuda(d_output, d_sum_w):
            ops.kernel_weighting_grad_cuda_float32(
                data, weights, sum_w, d_output, d_sum_w, d_data, d_weights)
        else:
            ops.kernel_weighting_grad_cpu_float32(
                data, weights, sum_w, d_output, d_sum_w, d_data, d_weights)
        return d_data, d_weights
