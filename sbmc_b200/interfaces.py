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


class SampleBasedDenoiserInterface(object):
    """model: nn.Module taking / returning dicts; lr: Adam step size; cuda: move
    the model (and every batch) to the GPU; fused_optimizer: see below;
    allow_tf32: the reference trains in fp32 (PyTorch 1.2 had no TF32), so the
    cuDNN / cuBLAS TF32 paths PyTorch enables by default for convolutions are
    switched OFF unless asked for -- the policy is set here, explicitly, and
    logged."""

    def __init__(self, model, lr=1e-4, cuda=False, fused_optimizer=False, allow_tf32=False):
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
        self.fused_optimizer = bool(fused_optimizer)
        if self.fused_optimizer:
            from .optim import FusedAdam
            self.optimizer = FusedAdam(self.model.parameters(), lr=lr)
        else:
            self.optimizer = th.optim.Adam(self.model.parameters(), lr=lr)

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
        self.optimizer.zero_grad()
        loss, out, tgt = self._scores(batch, fwd)
        loss.backward()
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
