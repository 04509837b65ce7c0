"""Fused ProgressiveKernelApply update (inference): one pass over the kernel
logits instead of the reference's Scatter2Gather -> max -> sub_ -> exp_ ->
KernelWeighting -> rescale chain (sbmc/modules.py:419-473).  Binds
``sbmc_progressive_splat_fwd_f32`` (include/sbmc_b200.h)."""
import torch as th

from . import _lib

__all__ = ["progressive_splat_update", "fused_available"]


def fused_available(data, kernels, *state):
    """The fused kernel has no backward: it serves calls that need no gradient."""
    ts = [t for t in (data, kernels) + state if t is not None]
    if not all(t.is_cuda and t.dtype == th.float32 for t in ts):
        return False
    if th.is_grad_enabled() and any(t.requires_grad for t in ts):
        return False
    return True


def progressive_splat_update(data, kernels, sum_r, sum_w, max_w, splat):
    """data [bs, c, h, w], kernels [bs, k*k, h, w] (or [bs, k, k, h, w]) logits;
    sum_r [bs, c, h, w], sum_w / max_w [bs, 1, h, w] or all None (first update).
    Returns new (sum_r, sum_w, max_w); the inputs are not modified."""
    data = data.contiguous()
    kernels = kernels.contiguous()
    bs, c, h, w = data.shape
    if kernels.dim() == 4:
        k = int(round(kernels.shape[1] ** 0.5))
        kh = kw = k
        if k * k != kernels.shape[1]:
            raise RuntimeError("kernels: channel count %d is not a square" % kernels.shape[1])
    else:
        kh, kw = kernels.shape[1], kernels.shape[2]
    if tuple(kernels.shape[-2:]) != (h, w) or kernels.shape[0] != bs:
        raise RuntimeError("kernels %s do not match data %s"
                           % (tuple(kernels.shape), tuple(data.shape)))
    first = sum_r is None
    if first:
        if sum_w is not None or max_w is not None:
            raise RuntimeError("all of sum_r, sum_w, max_w should be none")
        sum_r = th.empty_like(data)
        sum_w = data.new_empty(bs, 1, h, w)
        max_w = data.new_empty(bs, 1, h, w)
    else:
        sum_r = sum_r.contiguous().clone()
        sum_w = sum_w.contiguous().clone()
        max_w = max_w.contiguous().clone()
    lib = _lib.load()
    with th.cuda.device(data.device):
        rc = lib.sbmc_progressive_splat_fwd_f32(
            kernels.data_ptr(), data.data_ptr(), sum_r.data_ptr(), sum_w.data_ptr(),
            max_w.data_ptr(), bs, c, h, w, kh, kw, 1 if splat else 0, 1 if first else 0,
            th.cuda.current_stream(data.device).cuda_stream)
    _lib.check(rc, "progressive_splat")
    return sum_r, sum_w, max_w
