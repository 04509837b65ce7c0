"""Fused gradient clipping + Adam: the optimizer half of the training step.

The reference's step (sbmc/interfaces.py:78-106) is `loss.backward()`,
`clip_grad_norm_(params, 1000)`, `Adam.step()`: on the ~150 parameter tensors of
`Multisteps` eager PyTorch spends hundreds of small launches on the last two.
`FusedAdam.step(max_norm=...)` does them in three launches over device tables of
all tensors (`sbmc_multi_tensor_grad_norm_f32`, `sbmc_multi_tensor_adam_f32`),
without a host synchronisation: the clip coefficient stays on the device.

Same hyper-parameters, update rule and state layout (`step`, `exp_avg`,
`exp_avg_sq`) as `torch.optim.Adam` without weight decay / amsgrad, so state dicts
interchange.  CUDA fp32 parameters only; anything else raises (no fallback).
"""
import math

import torch as th

from . import _lib

__all__ = ["FusedAdam"]

_CHUNK = 65536      # SBMC_MT_CHUNK_ELEMS (include/sbmc_b200.h)


class _CudaBackend(object):
    """The launches of a step: libsbmc_b200's kernels on the parameters' CUDA device.
    (The test-suite swaps in the device sources compiled for the host,
    tests/native/optim_emul.cpp, to run the Python side without a GPU.)"""
    device_type = "cuda"

    def __init__(self):
        self.lib = _lib.load()

    def scope(self, dev):
        return th.cuda.device(dev)

    def grad_norm(self, dev, tensors, chunks, nchunks, partial, max_norm, out):
        _lib.check(self.lib.sbmc_multi_tensor_grad_norm_f32(
            tensors.data_ptr(), chunks.data_ptr(), nchunks, partial.data_ptr(), max_norm,
            out.data_ptr(), th.cuda.current_stream(dev).cuda_stream), "grad_norm")

    def adam(self, dev, tensors, chunks, nchunks, coef_ptr, *scalars):
        _lib.check(self.lib.sbmc_multi_tensor_adam_f32(
            tensors.data_ptr(), chunks.data_ptr(), nchunks, coef_ptr, *scalars,
            th.cuda.current_stream(dev).cuda_stream), "adam")

    def adam_devstep(self, dev, tensors, chunks, nchunks, coef_ptr, lr, beta1, beta2, eps, step):
        _lib.check(self.lib.sbmc_multi_tensor_adam_devstep_f32(
            tensors.data_ptr(), chunks.data_ptr(), nchunks, coef_ptr, lr, beta1, beta2, eps,
            step.data_ptr(), th.cuda.current_stream(dev).cuda_stream), "adam")

    def capturing(self):
        return th.cuda.is_current_stream_capturing()


def _backend():
    return _CudaBackend()


def _bump_version(p):
    try:
        th.autograd.graph.increment_version(p)
    except (AttributeError, RuntimeError):
        p.add_(0)                                 # in-place no-op: same effect


class FusedAdam(th.optim.Optimizer):
    """capturable=True: the step can be captured in a CUDA graph (interfaces.py
    `cuda_graph`): the step count is ONE device float shared by all parameters
    (`state[p]["step"]`, like torch.optim.Adam(capturable=True) keeps device steps), the
    bias corrections are derived from it on the device, and the device tables are built
    once per set of tensor addresses -- a captured step contains no host-to-device copy,
    so parameters, gradients and state must keep their addresses (static `.grad`s)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super(FusedAdam, self).__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.last_grad_norm = None      # device tensor [2]: norm, clip coefficient
        self.capturable = bool(capturable)
        self._dev_step = None
        self._table_cache = {}          # pointer tuple -> (tensors, chunks, nchunks, partial)
        self.generation = 0             # bumped when the state tensors are replaced

    def load_state_dict(self, state_dict):
        """torch replaces every state tensor: device tables and the shared step count are
        rebuilt on the next step, and anything that captured the old addresses (a CUDA graph
        of the training step) must be rebuilt too -- `generation` tells it."""
        super(FusedAdam, self).load_state_dict(state_dict)
        self._table_cache.clear()
        self._dev_step = None
        self.generation += 1

    def _cached_tables(self, rows, dev, backend):
        key = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
                    for p, g, m, v in rows)
        hit = self._table_cache.get(key)
        if hit is None:
            if getattr(backend, "capturing", lambda: False)():
                raise _lib.SbmcB200Error(
                    "FusedAdam: tensor addresses changed inside a CUDA-graph capture (run one "
                    "eager step with the same static gradients first)")
            tensors, chunks, nchunks = self._tables(rows, dev)
            partial = th.empty(max(nchunks, 1), dtype=th.float32, device=dev)
            if len(self._table_cache) > 8:
                self._table_cache.clear()
            hit = self._table_cache[key] = (tensors, chunks, nchunks, partial)
        return hit

    def _step_capturable(self, backend, groups, everything, dev, max_norm):
        if self._dev_step is None:
            steps = [self.state[r[0]].get("step") for r in everything]
            start = max([float(s) for s in steps if s is not None] + [0.0])
            self._dev_step = th.full((), start, dtype=th.float32, device=dev)
        for r in everything:
            self.state[r[0]]["step"] = self._dev_step
        coef_ptr = None
        with backend.scope(dev):
            tensors, chunks, nchunks, partial = self._cached_tables(everything, dev, backend)
            if max_norm is not None:
                if self.last_grad_norm is None or self.last_grad_norm.device != dev:
                    self.last_grad_norm = th.empty(2, dtype=th.float32, device=dev)
                backend.grad_norm(dev, tensors, chunks, nchunks, partial, float(max_norm),
                                  self.last_grad_norm)
                coef_ptr = self.last_grad_norm.data_ptr() + 4
            if len(groups) == 1:
                parts = [(groups[0][0], (tensors, chunks, nchunks))]
            else:
                parts = [(g, self._cached_tables(rows, dev, backend)[:3]) for g, rows in groups]
            for i, (group, (tn, ch, nc)) in enumerate(parts):
                beta1, beta2 = group["betas"]
                if i + 1 < len(parts):
                    raise _lib.SbmcB200Error("FusedAdam(capturable=True) supports one "
                                             "parameter group")
                backend.adam_devstep(dev, tn, ch, nc, coef_ptr, group["lr"], beta1, beta2,
                                     group["eps"], self._dev_step)
            if not getattr(backend, "capturing", lambda: False)():
                for r in everything:
                    _bump_version(r[0])

    @staticmethod
    def _tables(rows, dev):
        """rows: [(param, grad, exp_avg, exp_avg_sq)] -> device tables."""
        tensors, chunks = [], []
        for t, (p, g, m, v) in enumerate(rows):
            n = p.numel()
            tensors.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n))
            chunks.extend((t, start) for start in range(0, n, _CHUNK))
        host = th.tensor(tensors, dtype=th.int64).reshape(-1, 5)
        hchunks = th.tensor(chunks, dtype=th.int64).reshape(-1, 2)
        if dev.type == "cuda":
            host, hchunks = host.pin_memory(), hchunks.pin_memory()
        return host.to(dev, non_blocking=True), hchunks.to(dev, non_blocking=True), len(chunks)

    @th.no_grad()
    def step(self, closure=None, max_norm=None):
        """One Adam step on every parameter that has a gradient.  With `max_norm`
        the gradients are first scaled by min(1, max_norm / (total norm + 1e-6))
        like `clip_grad_norm_`; the norm / coefficient are left in
        `self.last_grad_norm` (device tensor, read it only when needed)."""
        loss = None
        if closure is not None:
            with th.enable_grad():
                loss = closure()
        backend = _backend()
        coef_ptr = None
        groups = []
        everything = []
        for group in self.param_groups:
            rows = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if not (p.device.type == backend.device_type and p.dtype == th.float32
                        and p.is_contiguous()
                        and g.dtype == th.float32 and g.is_contiguous() and g.device == p.device
                        and not g.is_sparse):
                    raise _lib.SbmcB200Error(
                        "FusedAdam wants contiguous float32 CUDA parameters and gradients")
                state = self.state[p]
                if not state:
                    state["step"] = th.zeros((), dtype=th.float32) if not self.capturable else None
                    state["exp_avg"] = th.zeros_like(p, memory_format=th.contiguous_format)
                    state["exp_avg_sq"] = th.zeros_like(p, memory_format=th.contiguous_format)
                rows.append((p, g, state["exp_avg"], state["exp_avg_sq"]))
            if rows:
                groups.append((group, rows))
                everything.extend(rows)
        if not everything:
            return loss
        dev = everything[0][0].device
        if any(r[0].device != dev for r in everything):
            raise _lib.SbmcB200Error("FusedAdam: all parameters must live on one device")
        if self.capturable:
            self._step_capturable(backend, groups, everything, dev, max_norm)
            return loss
        # tensors that have taken the same number of steps share one launch
        parts = []
        for group, rows in groups:
            by_step = {}
            for r in rows:
                by_step.setdefault(float(self.state[r[0]]["step"]), []).append(r)
            parts.extend((group, step, part) for step, part in sorted(by_step.items()))
        with backend.scope(dev):
            tables = None
            if max_norm is not None:
                tables = self._tables(everything, dev)
                tensors, chunks, nchunks = tables
                partial = th.empty(max(nchunks, 1), dtype=th.float32, device=dev)
                self.last_grad_norm = th.empty(2, dtype=th.float32, device=dev)
                backend.grad_norm(dev, tensors, chunks, nchunks, partial, float(max_norm),
                                  self.last_grad_norm)
                coef_ptr = self.last_grad_norm.data_ptr() + 4
            for group, step, part in parts:
                beta1, beta2 = group["betas"]
                t = step + 1
                if tables is None or len(parts) > 1:
                    tables = self._tables(part, dev)
                tensors, chunks, nchunks = tables
                backend.adam(dev, tensors, chunks, nchunks, coef_ptr, group["lr"], beta1, beta2,
                             group["eps"], 1.0 - beta1 ** t, math.sqrt(1.0 - beta2 ** t))
                for r in part:
                    self.state[r[0]]["step"] += 1
                    # the kernel wrote through raw pointers: bump the autograd version
                    # counter so that caches keyed on it (conv1x1.prepare, unet_fast)
                    # see the update
                    _bump_version(r[0])
        return loss
