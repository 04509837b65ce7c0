"""Image losses used by the training step (callers of the hot path, kept so that
config 4 -- the train.py step -- can run): same classes and formulas as the
reference's ``sbmc/losses.py:26-121``.  Elementwise torch ops; not accelerated."""
import torch as th

__all__ = ["RelativeMSE", "SMAPE", "TonemappedMSE", "TonemappedRelativeMSE"]


def _tonemap(im):
    """Reinhard tonemapper on the non-negative part: x / (1 + x) (losses.py:110-121)."""
    im = th.clamp(im, min=0)
    return im / (1 + im)


class RelativeMSE(th.nn.Module):
    """0.5 * mean((im - ref)^2 / (ref^2 + eps))."""

    def __init__(self, eps=1e-2):
        super(RelativeMSE, self).__init__()
        self.eps = eps

    def forward(self, im, ref):
        return 0.5 * th.mean((im - ref) ** 2 / (ref ** 2 + self.eps))


class SMAPE(th.nn.Module):
    """mean(|im - ref| / (eps + |im| + |ref|)); the denominator carries no gradient."""

    def __init__(self, eps=1e-2):
        super(SMAPE, self).__init__()
        self.eps = eps

    def forward(self, im, ref):
        denom = self.eps + th.abs(im.detach()) + th.abs(ref.detach())
        return (th.abs(im - ref) / denom).mean()


class TonemappedMSE(th.nn.Module):
    """0.5 * mean((t(im) - t(ref))^2) on Reinhard-tonemapped images."""

    def __init__(self, eps=1e-2):
        super(TonemappedMSE, self).__init__()
        self.eps = eps

    def forward(self, im, ref):
        return 0.5 * th.mean((_tonemap(im) - _tonemap(ref)) ** 2)


class TonemappedRelativeMSE(th.nn.Module):
    """RelativeMSE on Reinhard-tonemapped images (the training loss,
    sbmc/interfaces.py:48)."""

    def __init__(self, eps=1e-2):
        super(TonemappedRelativeMSE, self).__init__()
        self.eps = eps

    def forward(self, im, ref):
        im, ref = _tonemap(im), _tonemap(ref)
        return 0.5 * th.mean((im - ref) ** 2 / (ref ** 2 + self.eps))
