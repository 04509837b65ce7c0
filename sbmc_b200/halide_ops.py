"""Drop-in for the reference's native-op module ``sbmc.halide_ops``.

The reference synthesises a pybind11 module with six functions
(setup.py:65-84; ``m.def`` at halide_pytorch/halide_pytorch/extension.py:168-173)
and ``sbmc/functions.py:24-27`` imports it as ``ops``.  This module exposes the
same six names with the same argument order (Halide inputs, then the
caller-allocated outputs) and the same failure mode (``RuntimeError``), backed
by libsbmc_b200.so:

* ``*_cuda_float32``: CUDA tensors, asynchronous on PyTorch's current stream of
  the tensors' device;
* ``*_cpu_float32``: HOST tensors.  There is no CPU compute path -- the buffers
  are streamed through the GPU by the library's host entry points (pinned
  tensors avoid a staging copy inside the driver).
"""
import torch as th

from . import _lib

__all__ = [
    "scatter2gather_cuda_float32", "kernel_weighting_cuda_float32",
    "kernel_weighting_grad_cuda_float32", "scatter2gather_cpu_float32",
    "kernel_weighting_cpu_float32", "kernel_weighting_grad_cpu_float32",
]


def _check(name, t, ndim, cuda):
    if not isinstance(t, th.Tensor):
        raise RuntimeError("%s: expected a torch.Tensor" % name)
    if t.dtype != th.float32:
        raise RuntimeError("%s: expected float32, got %s" % (name, t.dtype))
    if t.dim() != ndim:
        raise RuntimeError("%s: expected %d dimensions, got %d" % (name, ndim, t.dim()))
    if not t.is_contiguous():
        raise RuntimeError("%s: tensor must be contiguous" % name)
    if t.is_cuda != cuda:
        raise RuntimeError("%s: tensor must be on %s" % (name, "a CUDA device" if cuda else "the host"))


def _same_device(*ts):
    dev = ts[0].device
    for t in ts[1:]:
        if t.device != dev:
            raise RuntimeError("all tensors must be on the same device")
    return dev


def _shape_eq(name, t, shape):
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))


def _stream(dev):
    return th.cuda.current_stream(dev).cuda_stream


def _host_device():
    if not th.cuda.is_available():
        raise RuntimeError(
            "sbmc_b200 has no CPU compute path: host tensors are streamed "
            "through a CUDA device and none is available")
    return th.cuda.current_device()


# -- scatter2gather ---------------------------------------------------------------
def _s2g(weights, output, cuda):
    _check("weights", weights, 5, cuda)
    _check("output", output, 5, cuda)
    _shape_eq("output", output, weights.shape)
    n, kh, kw, h, w = weights.shape
    lib = _lib.load()
    if cuda:
        dev = _same_device(weights, output)
        with th.cuda.device(dev):
            rc = lib.sbmc_scatter2gather_f32(weights.data_ptr(), output.data_ptr(),
                                             n, kh, kw, h, w, _stream(dev))
    else:
        rc = lib.sbmc_scatter2gather_host_f32(weights.data_ptr(), output.data_ptr(),
                                              n, kh, kw, h, w, _host_device())
    _lib.check(rc, "scatter2gather")
    return 0


def scatter2gather_cuda_float32(weights, output):
    return _s2g(weights, output, True)


def scatter2gather_cpu_float32(weights, output):
    return _s2g(weights, output, False)


# -- kernel_weighting -------------------------------------------------------------
def _kw(data, weights, output, sum_w, cuda):
    _check("data", data, 4, cuda)
    _check("weights", weights, 5, cuda)
    _check("output", output, 4, cuda)
    _check("sum_w", sum_w, 3, cuda)
    n, c, h, w = data.shape
    _, kh, kw, _, _ = weights.shape
    _shape_eq("weights", weights, (n, kh, kw, h, w))
    _shape_eq("output", output, (n, c, h, w))
    _shape_eq("sum_w", sum_w, (n, h, w))
    lib = _lib.load()
    if cuda:
        dev = _same_device(data, weights, output, sum_w)
        with th.cuda.device(dev):
            rc = lib.sbmc_kernel_weighting_fwd_f32(
                data.data_ptr(), weights.data_ptr(), output.data_ptr(),
                sum_w.data_ptr(), n, c, h, w, kh, kw, _stream(dev))
    else:
        rc = lib.sbmc_kernel_weighting_fwd_host_f32(
            data.data_ptr(), weights.data_ptr(), output.data_ptr(),
            sum_w.data_ptr(), n, c, h, w, kh, kw, _host_device())
    _lib.check(rc, "kernel_weighting")
    return 0


def kernel_weighting_cuda_float32(data, weights, output, sum_w):
    return _kw(data, weights, output, sum_w, True)


def kernel_weighting_cpu_float32(data, weights, output, sum_w):
    return _kw(data, weights, output, sum_w, False)


# -- kernel_weighting_grad ----------------------------------------------------------
def _kw_grad(data, weights, sum_w, d_output, d_sum_w, d_data, d_weights, cuda):
    _check("data", data, 4, cuda)
    _check("weights", weights, 5, cuda)
    _check("sum_w", sum_w, 3, cuda)
    _check("d_output", d_output, 4, cuda)
    _check("d_sum_w", d_sum_w, 3, cuda)
    _check("d_data", d_data, 4, cuda)
    _check("d_weights", d_weights, 5, cuda)
    n, c, h, w = data.shape
    _, kh, kw, _, _ = weights.shape
    _shape_eq("weights", weights, (n, kh, kw, h, w))
    _shape_eq("sum_w", sum_w, (n, h, w))
    _shape_eq("d_output", d_output, (n, c, h, w))
    _shape_eq("d_sum_w", d_sum_w, (n, h, w))
    _shape_eq("d_data", d_data, (n, c, h, w))
    _shape_eq("d_weights", d_weights, (n, kh, kw, h, w))
    lib = _lib.load()
    if cuda:
        dev = _same_device(data, weights, sum_w, d_output, d_sum_w, d_data, d_weights)
        with th.cuda.device(dev):
            rc = lib.sbmc_kernel_weighting_bwd_f32(
                data.data_ptr(), weights.data_ptr(), sum_w.data_ptr(),
                d_output.data_ptr(), d_sum_w.data_ptr(), d_data.data_ptr(),
                d_weights.data_ptr(), n, c, h, w, kh, kw, _stream(dev))
    else:
        rc = lib.sbmc_kernel_weighting_bwd_host_f32(
            data.data_ptr(), weights.data_ptr(), sum_w.data_ptr(),
            d_output.data_ptr(), d_sum_w.data_ptr(), d_data.data_ptr(),
            d_weights.data_ptr(), n, c, h, w, kh, kw, _host_device())
    _lib.check(rc, "kernel_weighting_grad")
    return 0


def kernel_weighting_grad_cuda_float32(data, weights, sum_w, d_output, d_sum_w,
                                       d_data, d_weights):
    return _kw_grad(data, weights, sum_w, d_output, d_sum_w, d_data, d_weights, True)


def kernel_weighting_grad_cpu_float32(data, weights, sum_w, d_output, d_sum_w,
                                      d_data, d_weights):
    return _kw_grad(data, weights, sum_w, d_output, d_sum_w, d_data, d_weights, False)
