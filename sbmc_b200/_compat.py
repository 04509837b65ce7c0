"""What the reference takes from the external `ttools` package
(torch-tools==0.0.36, not part of the reference tree and not installed here).
The hot path uses two helpers: `ttools.get_logger` (sbmc/modules.py:31,
sbmc/functions.py:22) and `ttools.modules.image_operators.crop_like`
(sbmc/models.py:27,206).  The two scripts additionally use its checkpointer,
trainer and callbacks (scripts/train.py:96-121, scripts/denoise.py:107,134-135):
small equivalents are at the end of this file.
"""
import logging

__all__ = ["get_logger", "crop_like", "Checkpointer", "Trainer", "LoggingCallback",
           "CheckpointingCallback"]


def get_logger(name):
    log = logging.getLogger(name)
    if not logging.getLogger().handlers and not log.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(
            "[%(asctime)s] %(name)s %(levelname)s: %(message)s", "%H:%M:%S"))
        log.addHandler(handler)
        log.propagate = False
    return log


def crop_like(src, tgt):
    """Centre-crop the last two dims of `src` to those of `tgt` (call sites:
    sbmc/models.py:206,271-283).  The difference is split evenly, the odd row /
    column going to the bottom / right like ttools does."""
    sh, sw = src.shape[-2:]
    th_, tw = tgt.shape[-2:]
    dh, dw = sh - th_, sw - tw
    if dh < 0 or dw < 0:
        raise ValueError("crop_like: source %s smaller than target %s"
                         % (tuple(src.shape), tuple(tgt.shape)))
    if dh == 0 and dw == 0:
        return src
    top, left = dh // 2, dw // 2
    return src[..., top:top + th_, left:left + tw]


# -- the training / checkpoint helpers the reference's scripts take from `ttools` ------
# (scripts/train.py:96-121, scripts/denoise.py:107,134-135).  torch-tools 0.0.36 is
# external and absent: these follow its documented behaviour (one .pth per save
# holding {"model", "optimizers", "schedulers", "meta", "extras"}, newest file wins);
# parity unpinned, no reference test covers them.
import os as _os
import re as _re
import time as _time

import torch as _th


class Checkpointer(object):
    """Saves / restores a model (+ optimizers) under `root`."""

    EXTENSION = ".pth"

    def __init__(self, root, model=None, meta=None, optimizers=None, schedulers=None,
                 prefix=None):
        self.root = root
        self.model = model
        self.meta = meta
        self.prefix = prefix or ""
        as_list = lambda x: [] if x is None else (list(x) if isinstance(x, (list, tuple)) else [x])  # noqa: E731
        self.optimizers = as_list(optimizers)
        self.schedulers = as_list(schedulers)
        self.log = get_logger(__name__)

    def _path(self, name):
        return _os.path.join(self.root, self.prefix + _os.path.splitext(name)[0] + self.EXTENSION)

    def save(self, name, extras=None):
        _os.makedirs(self.root, exist_ok=True)
        path = self._path(name)
        _th.save({"model": self.model.state_dict() if self.model is not None else None,
                  "optimizers": [o.state_dict() for o in self.optimizers],
                  "schedulers": [s.state_dict() for s in self.schedulers],
                  "meta": self.meta, "extras": extras}, path)
        return path

    def sorted_checkpoints(self):
        """Newest first."""
        if not _os.path.isdir(self.root):
            return []
        pattern = _re.compile(_re.escape(self.prefix) + r".*" + _re.escape(self.EXTENSION) + "$")
        found = [_os.path.join(self.root, f) for f in _os.listdir(self.root) if pattern.match(f)]
        return sorted(found, key=_os.path.getmtime, reverse=True)

    def load(self, path):
        chkpt = _th.load(path, map_location="cpu", weights_only=False)
        if self.model is not None and chkpt.get("model") is not None:
            self.model.load_state_dict(chkpt["model"])
        for opt, state in zip(self.optimizers, chkpt.get("optimizers") or []):
            opt.load_state_dict(state)
        for sch, state in zip(self.schedulers, chkpt.get("schedulers") or []):
            sch.load_state_dict(state)
        return chkpt.get("extras"), chkpt.get("meta")

    def load_latest(self):
        """(extras, meta) of the newest loadable checkpoint, else (None, None)."""
        for path in self.sorted_checkpoints():
            try:
                return self.load(path)
            except Exception as e:        # corrupt / partial file: try the next one
                self.log.warning("could not load %s (%s)", path, e)
        return None, None

    @classmethod
    def load_meta(cls, root, prefix=None):
        for path in cls(root, prefix=prefix).sorted_checkpoints():
            return _th.load(path, map_location="cpu", weights_only=False).get("meta")
        return None


class Trainer(object):
    """Minimal epoch loop over a ModelInterface (forward / backward /
    init_validation / update_validation); callbacks get `epoch_start(epoch)`,
    `batch_end(batch, fwd, bwd)`, `validation_end(val)`, `epoch_end()`."""

    def __init__(self, interface):
        self.interface = interface
        self.callbacks = []
        self.log = get_logger(__name__)

    def add_callback(self, callback):
        self.callbacks.append(callback)

    def _emit(self, name, *args):
        for cb in self.callbacks:
            fn = getattr(cb, name, None)
            if fn is not None:
                fn(*args)

    def train(self, dataloader, num_epochs=None, val_dataloader=None, max_steps=None):
        epoch = 0
        steps = 0
        try:
            while num_epochs is None or epoch < num_epochs:
                self._emit("epoch_start", epoch)
                for batch in dataloader:
                    step = getattr(self.interface, "train_step", None)
                    if step is not None:        # forward + backward in one call (CUDA graphs)
                        fwd, bwd = step(batch)
                    else:
                        fwd = self.interface.forward(batch)
                        bwd = self.interface.backward(batch, fwd)
                    self._emit("batch_end", batch, fwd, bwd)
                    steps += 1
                    if max_steps is not None and steps >= max_steps:
                        break
                if val_dataloader is not None:
                    running = self.interface.init_validation()
                    with _th.no_grad():
                        for batch in val_dataloader:
                            fwd = self.interface.forward(batch)
                            running = self.interface.update_validation(batch, fwd, running)
                    self._emit("validation_end", running)
                self._emit("epoch_end")
                epoch += 1
                if max_steps is not None and steps >= max_steps:
                    break
        except KeyboardInterrupt:
            self.log.info("interrupted after %d steps", steps)
        self._emit("training_end")
        return steps


class LoggingCallback(object):
    """Running averages of the interface's scalars, logged every `frequency` steps."""

    def __init__(self, keys, frequency=50):
        self.keys, self.frequency = list(keys), frequency
        self.log = get_logger(__name__)
        self.step = 0
        self.start = _time.time()

    def batch_end(self, batch, fwd, bwd):
        self.step += 1
        if self.step % self.frequency == 0:
            rate = self.step / max(_time.time() - self.start, 1e-9)
            self.log.info("step %d  %s  (%.2f steps/s)", self.step,
                          "  ".join("%s %.5g" % (k, bwd[k]) for k in self.keys), rate)

    def validation_end(self, val):
        self.log.info("validation  %s", "  ".join("%s %.5g" % (k, val[k]) for k in self.keys))


class CheckpointingCallback(object):
    """Saves `epoch_<n>` at the end of every epoch and `training_end` at the end."""

    def __init__(self, checkpointer):
        self.checkpointer = checkpointer
        self.epoch = 0

    def epoch_start(self, epoch):
        self.epoch = epoch

    def epoch_end(self):
        self.checkpointer.save("epoch_%d" % self.epoch, extras={"epoch": self.epoch})

    def training_end(self):
        self.checkpointer.save("training_end", extras={"epoch": self.epoch})
