"""The two helpers the hot path takes from the external `ttools` package
(torch-tools==0.0.36, not part of the reference tree and not installed here):
`ttools.get_logger` (sbmc/modules.py:31, sbmc/functions.py:22) and
`ttools.modules.image_operators.crop_like` (sbmc/models.py:27,206).
"""
import logging

__all__ = ["get_logger", "crop_like"]


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
