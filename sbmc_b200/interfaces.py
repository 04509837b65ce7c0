"""Training-step glue, API-compatible with the reference's
``sbmc/interfaces.py:35-132`` (``SampleBasedDenoiserInterface``: forward /
backward / init_validation / update_validation) without the external ``ttools``
base class.  It is a caller of the hot path (config 4 of BASELINE.json)."""
import math

import torch as th

from . import losses
from ._compat import crop_like, get_logger

__all__ = ["SampleBasedDenoiserInterface"]

LOG = get_logger(__name__)


class SampleBasedDenoiserInterface(object):
    """Args: model (nn.Module), lr (float), cuda (bool)."""

    def __init__(self, model, lr=1e-4, cuda=False):
        self.device = "cuda" if cuda else "cpu"
        self.model = model
        self.loss_fn = losses.TonemappedRelativeMSE()
        self.rmse_fn = losses.RelativeMSE()
        if cuda:
            self.model.cuda()
        self.optimizer = th.optim.Adam(self.model.parameters(), lr=lr)

    def forward(self, batch):
        for k in batch:
            if isinstance(batch[k], th.Tensor):
                batch[k] = batch[k].to(self.device)
        return self.model(batch)

    def backward(self, batch, fwd):
        self.optimizer.zero_grad()
        out = fwd["radiance"]
        tgt = crop_like(batch["target_image"], out)
        loss = self.loss_fn(out, tgt)
        loss.backward()
        value = loss.item()
        if math.isinf(value):
            LOG.error("Loss is infinite, there might be outliers in the data.")
            raise RuntimeError("Infinite loss at train time.")
        if math.isnan(value):
            LOG.error("NaN in the loss, there might be outliers in the data.")
            raise RuntimeError("NaN loss at train time.")
        clip = 1000
        actual = th.nn.utils.clip_grad_norm_(self.model.parameters(), clip)
        if actual > clip:
            LOG.info("Clipped gradients {} -> {}".format(clip, actual))
        self.optimizer.step()
        with th.no_grad():
            rmse = self.rmse_fn(out, tgt)
        return {"loss": value, "rmse": rmse.item()}

    def init_validation(self):
        return {"loss": 0.0, "rmse": 0.0, "n": 0}

    def update_validation(self, batch, fwd, running):
        with th.no_grad():
            out = fwd["radiance"]
            tgt = crop_like(batch["target_image"], out)
            loss = self.loss_fn(out, tgt).item()
            rmse = self.rmse_fn(out, tgt).item()
        b = out.shape[0]
        n = running["n"] + b
        return {"loss": running["loss"] - (1.0 / n) * (running["loss"] - b * loss),
                "rmse": running["rmse"] - (1.0 / n) * (running["rmse"] - b * rmse), "n": n}
