"""One optimisation step of the denoiser: the caller behind BASELINE.json's
config 4.  Same public surface as the reference's training glue
(``sbmc/interfaces.py:35-132``: ``forward`` / ``backward`` / ``init_validation`` /
``update_validation``, Adam on the model parameters, tonemapped relative MSE as
the loss, gradient-norm clip at 1000, non-finite-loss guard) without the external
``ttools.ModelInterface`` base class.
"""
import math

import torch as th

from . import losses
from ._compat import crop_like, get_logger

__all__ = ["SampleBasedDenoiserInterface"]

LOG = get_logger(__name__)
_GRAD_CLIP = 1000


class _GraphedStep(object):
    """One training step of `SampleBasedDenoiserInterface` on static buffers, captured in a
    CUDA graph.  Gradients are static too (allocated once, zeroed inside the graph), so
    the fused optimizer's device tables never change.  The two eager warm-up steps that
    PyTorch's capture protocol asks for run on the first batch and are undone (parameters
    and optimizer state restored) before the capture."""

    def __init__(self, iface, batch):
        self.iface = iface
        self.static = {k: (v.detach().clone() if isinstance(v, th.Tensor) else v)
                       for k, v in batch.items()}
        self.params = [p for p in iface.model.parameters() if p.requires_grad]
        for p in self.params:
            if p.grad is None:                  # (distributed: already views of the flat buffer)
                p.grad = th.zeros_like(p, memory_format=th.contiguous_format)
        opt = iface.optimizer
        saved_p = [p.detach().clone() for p in self.params]
        saved_s = {p: {k: (v.clone() if isinstance(v, th.Tensor) else v)
                       for k, v in opt.state[p].items()}
                   for p in self.params if opt.state.get(p)}
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):
            for _ in range(2):
                self._body()
        th.cuda.current_stream().wait_stream(side)
        th.cuda.synchronize()
        with th.no_grad():
            for p, q in zip(self.params, saved_p):
                p.copy_(q)
                for k, v in opt.state[p].items():
                    if isinstance(v, th.Tensor):
                        old = saved_s.get(p, {}).get(k)
                        v.zero_() if old is None else v.copy_(old)
        del saved_p, saved_s
        self.generation = getattr(opt, "generation", 0)
        self.graph = th.cuda.CUDAGraph()
        # thread_local: other threads (the PrefetchLoader's reader, NCCL's watchdog) may issue
        # CUDA calls while this thread captures
        with th.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._body()
        LOG.info("training step captured in a CUDA graph (%s)",
                 ", ".join("%s %s" % (k, tuple(v.shape)) for k, v in self.static.items()
                           if isinstance(v, th.Tensor)))

    def _body(self):
        iface = self.iface
        if iface.distributed:
            iface._flat_grad.zero_()
        else:
            th._foreach_zero_([p.grad for p in self.params])
        fwd = iface.model(self.static)
        loss, out, tgt = iface._scores(self.static, fwd)
        loss.backward()
        iface._average_grads()                  # NCCL all-reduce: a node of the captured graph
        iface.optimizer.step(max_norm=_GRAD_CLIP)
        with th.no_grad():
            self.scalars = th.stack([loss.detach().float().reshape(()),
                                     iface.rmse_fn(out, tgt).float().reshape(()),
                                     iface.optimizer.last_grad_norm[0].reshape(())])
            self.out = out.detach()

    def run(self, batch):
        for k, v in batch.items():
            if isinstance(v, th.Tensor):
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        from .optim import _bump_version
        for p in self.params:               # the replay wrote the parameters behind autograd's back
            _bump_version(p)
        value, rmse, norm = self.scalars.tolist()
        if not math.isfinite(value):
            kind = "NaN" if math.isnan(value) else "Infinite"
            LOG.error("%s loss, there might be outliers in the data.", kind)
            raise RuntimeError("%s loss at train time." % kind)
        if norm > _GRAD_CLIP:
            LOG.info("Clipped gradients %s -> %s", _GRAD_CLIP, norm)
        return {"radiance": self.out}, {"loss": value, "rmse": rmse}


class SampleBasedDenoiserInterface(object):
    """model: nn.Module taking / returning dicts; lr: Adam step size; cuda: move
    the model (and every batch) to the GPU; fused_optimizer: see below;
    allow_tf32: the reference trains in fp32 (PyTorch 1.2 had no TF32), so the
    cuDNN / cuBLAS TF32 paths PyTorch enables by default for convolutions are
    switched OFF unless asked for -- the policy is set here, explicitly, and
    logged."""

    def __init__(self, model, lr=1e-4, cuda=False, fused_optimizer=False, allow_tf32=False,
                 cuda_graph=False, distributed=False):
        if cuda_graph and not (cuda and fused_optimizer):
            raise ValueError("cuda_graph needs cuda=True and fused_optimizer=True")
        self.allow_tf32 = bool(allow_tf32)
        th.backends.cudnn.allow_tf32 = self.allow_tf32
        th.backends.cuda.matmul.allow_tf32 = self.allow_tf32
        LOG.info("fp32 convolution policy: TF32 %s", "allowed" if self.allow_tf32 else "off")
        self.model = model.cuda() if cuda else model
        self.device = "cuda" if cuda else "cpu"
        self.loss_fn = losses.TonemappedRelativeMSE()
        self.rmse_fn = losses.RelativeMSE()
        # fused_optimizer (extra, CUDA only): clipping + Adam over all parameter
        # tensors in three launches (sbmc_b200.optim.FusedAdam)
        # cuda_graph (extra, CUDA only): `train_step` captures forward + loss + backward +
        # clip + Adam of one batch shape in a CUDA graph and replays it (the step is a few
        # hundred small launches: without the graph the host, not the GPU, sets its pace)
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        self.fused_optimizer = bool(fused_optimizer)
        if self.fused_optimizer:
            from .optim import FusedAdam
            self.optimizer = FusedAdam(self.model.parameters(), lr=lr,
                                       capturable=self.cuda_graph)
        else:
            self.optimizer = th.optim.Adam(self.model.parameters(), lr=lr)
        # distributed (extra): data-parallel training over torch.distributed (one process
        # per GPU, NCCL; gloo on CPU in the tests).  The reference is single-device.  Every
        # rank trains on its own batches; after backward the gradients -- views of ONE flat
        # fp32 buffer -- are averaged with one all-reduce, before clipping and Adam.
        self.distributed = bool(distributed)
        self._flat_grad = None
        if self.distributed:
            self._setup_distributed()

    def _setup_distributed(self):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("distributed=True needs an initialised torch.distributed group")
        self._dist = dist
        self.world = dist.get_world_size()
        params = [p for p in self.model.parameters() if p.requires_grad]
        with th.no_grad():
            for p in params:                          # same starting point on every rank
                dist.broadcast(p.data, 0)
        self._flat_grad = th.zeros(sum(p.numel() for p in params), dtype=th.float32,
                                   device=params[0].device)
        off = 0
        for p in params:
            if p.dtype != th.float32 or not p.is_contiguous():
                raise RuntimeError("distributed training wants contiguous float32 parameters")
            p.grad = self._flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()

    def _zero_grads(self):
        if self.distributed:
            self._flat_grad.zero_()                   # the gradients stay views of the buffer
        else:
            self.optimizer.zero_grad()

    def _average_grads(self):
        """One all-reduce over the flat gradient buffer, then the mean."""
        if self.distributed and self.world > 1:
            self._dist.all_reduce(self._flat_grad)
            self._flat_grad.mul_(1.0 / self.world)

    # -- helpers -----------------------------------------------------------------
    def _to_device(self, batch):
        for key, value in batch.items():
            if isinstance(value, th.Tensor):
                batch[key] = value.to(self.device)
        return batch

    def _scores(self, batch, fwd):
        """(loss, rmse) tensors of a forward result against the batch's target,
        cropped to the (smaller) network output."""
        out = fwd["radiance"]
        tgt = crop_like(batch["target_image"], out)
        return self.loss_fn(out, tgt), out, tgt

    # -- ttools.ModelInterface protocol ---------------------------------------------
    def forward(self, batch):
        return self.model(self._to_device(batch))

    def backward(self, batch, fwd):
        self._zero_grads()
        loss, out, tgt = self._scores(batch, fwd)
        loss.backward()
        self._average_grads()
        value = loss.item()
        if not math.isfinite(value):      # outliers in the data show up here
            kind = "NaN" if math.isnan(value) else "Infinite"
            LOG.error("%s loss, there might be outliers in the data.", kind)
            raise RuntimeError("%s loss at train time." % kind)
        if self.fused_optimizer:
            self.optimizer.step(max_norm=_GRAD_CLIP)
            norm = self.optimizer.last_grad_norm[0]
        else:
            norm = th.nn.utils.clip_grad_norm_(self.model.parameters(), _GRAD_CLIP)
            self.optimizer.step()
        # the gradient norm is only logged: it is read together with the rmse in ONE
        # device-to-host transfer after the optimizer has been launched (no extra
        # synchronisation between the clip and the Adam kernels)
        with th.no_grad():
            rmse, norm = th.stack([self.rmse_fn(out, tgt).float().reshape(()),
                                   norm.float().reshape(()).to(out.device)]).tolist()
        if norm > _GRAD_CLIP:
            LOG.info("Clipped gradients %s -> %s", _GRAD_CLIP, norm)
        return {"loss": value, "rmse": rmse}

    def train_step(self, batch):
        """forward + backward in one call -> (fwd, bwd) as `forward` / `backward` return
        them.  With cuda_graph the whole step is one graph replay per batch signature."""
        if not self.cuda_graph:
            fwd = self.forward(batch)
            return fwd, self.backward(batch, fwd)
        batch = self._to_device(batch)
        key = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in batch.items()
                           if isinstance(v, th.Tensor)))
        step = self._graphs.get(key)
        if step is not None and step.generation != getattr(self.optimizer, "generation", 0):
            self._graphs.clear()          # the optimizer's state tensors were replaced
            step = None
        if step is None:
            # one graph per batch signature: randomised sample counts (2 .. 8 spp) and a short
            # last batch of an epoch all stay resident; beyond that the oldest goes
            if len(self._graphs) >= 16:
                oldest = next(iter(self._graphs))
                self._graphs.pop(oldest).graph = None
            step = self._graphs[key] = _GraphedStep(self, batch)
        return step.run(batch)

    def close(self):
        """Drop the captured graphs (and the NCCL nodes inside them) before the process
        group is destroyed / the process exits: a live graph that references the
        communicator can stall the teardown."""
        import gc
        for step in self._graphs.values():
            step.graph = None
            step.iface = None
        self._graphs.clear()
        gc.collect()
        if self.device == "cuda":
            th.cuda.synchronize()

    def init_validation(self):
        return {"loss": 0.0, "rmse": 0.0, "n": 0}

    def update_validation(self, batch, fwd, running):
        """Running means weighted by the batch size (it may vary)."""
        with th.no_grad():
            loss, out, tgt = self._scores(batch, fwd)
            loss, rmse = loss.item(), self.rmse_fn(out, tgt).item()
        b = out.shape[0]
        n = running["n"] + b
        step = 1.0 / n
        return {"loss": running["loss"] - step * (running["loss"] - b * loss),
                "rmse": running["rmse"] - step * (running["rmse"] - b * rmse), "n": n}
