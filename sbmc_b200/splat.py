"""Fused ProgressiveKernelApply update (inference): one pass over the kernel
logits instead of the reference's Scatter2Gather -> max -> sub_ -> exp_ ->
KernelWeighting -> rescale chain (sbmc/modules.py:419-473).  Binds
``sbmc_progressive_splat_fwd_f32`` (include/sbmc_b200.h)."""
import torch as th

from . import _lib

__all__ = ["progressive_splat_update", "fused_available", "ProgressiveSplat",
           "fused_training_available"]


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


_BWD_SHAPES = {(3, 21), (3, 5), (3, 3), (3, 7)}       # (channels, k) with a fused backward


def fused_training_available(data, kernels, *state):
    """Whether ProgressiveSplat (fused forward AND backward) can serve this call."""
    ts = [t for t in (data, kernels) + state if t is not None]
    if not all(t.is_cuda and t.dtype == th.float32 for t in ts):
        return False
    k = int(round(kernels.shape[1] ** 0.5))
    return (k * k == kernels.shape[1] and (data.shape[1], k) in _BWD_SHAPES
            and data.shape[-1] % 4 == 0)


class ProgressiveSplat(th.autograd.Function):
    """One splat-mode ProgressiveKernelApply update (sbmc/modules.py:419-473) as a
    single autograd node: fused one-pass forward, fused one-pass backward in
    scatter space (csrc/splat_bwd.cu).  Only the logits are kept for backward,
    not the Scatter2Gather / exp intermediates of the composed chain.

    forward(data, kernels, sum_r, sum_w, max_w) -> (sum_r', sum_w', max_w'); pass
    None for the three state tensors on the first update."""

    @staticmethod
    def forward(ctx, data, kernels, sum_r, sum_w, max_w):
        first = sum_r is None
        data = data.contiguous()
        kernels = kernels.contiguous()
        new_r, new_w, new_m = progressive_splat_update(data, kernels, sum_r, sum_w, max_w, True)
        ctx.first = first
        if first:
            ctx.save_for_backward(data, kernels, new_r, new_w, new_m)
        else:
            ctx.save_for_backward(data, kernels, new_r, new_w, new_m, sum_r, sum_w, max_w)
        return new_r, new_w, new_m

    @staticmethod
    def backward(ctx, g_r, g_w, g_m):
        saved = ctx.saved_tensors
        data, kernels, new_r, new_w, new_m = saved[:5]
        bs, c, h, w = data.shape
        k = int(round(kernels.shape[1] ** 0.5))
        g_r = g_r.contiguous()
        # gradient reaching the running max: dL/dm' minus what every rescaled
        # term loses when m' grows (d sum'/dm' = -sum')
        t_max = g_m - ((g_r * new_r).sum(1, keepdim=True) + g_w * new_w)
        d_state = (None, None, None)
        if ctx.first:
            t_k = t_max
        else:
            sum_r, sum_w, max_w = saved[5:]
            # Where the running max came from.  Deviation from the composed chain
            # (th.max(kmax, max_w) of sbmc/modules.py:449): on an exact tie torch.max
            # splits the gradient 0.5 / 0.5 between the two operands, here all of it
            # goes to the previous state; and when the arg-max tap is an out-of-image
            # zero logit its share is still scattered to a real tap.  Both only move
            # gradient between terms whose total derivative w.r.t. the max cancels in
            # sum_r / sum_w (the output is invariant to the max), so the output
            # gradient is unaffected; tests/test_modules.py checks values and gradients
            # against the composed chain away from exact ties.
            from_prev = new_m == max_w          # the max came from earlier samples
            a = th.exp(max_w - new_m)
            d_a = (g_r * sum_r).sum(1, keepdim=True) + g_w * sum_w
            zero = th.zeros_like(t_max)
            t_k = th.where(from_prev, zero, t_max)
            d_state = (a * g_r, a * g_w, d_a * a + th.where(from_prev, t_max, zero))
        planes = th.cat([g_r, g_w, new_m, t_k], 1).contiguous()
        d_kernels = th.empty_like(kernels)
        d_data = th.empty_like(data)
        lib = _lib.load()
        with th.cuda.device(data.device):
            rc = lib.sbmc_progressive_splat_bwd_f32(
                planes.data_ptr(), kernels.data_ptr(), data.data_ptr(),
                d_kernels.data_ptr(), d_data.data_ptr(), bs, c, h, w, k, k,
                th.cuda.current_stream(data.device).cuda_stream)
        _lib.check(rc, "progressive_splat_bwd")
        return (d_data, d_kernels) + d_state
