"""Analytic values of the four losses (reference tests/test_losses.py:17-128) and
the training-step interface (sbmc/interfaces.py:62-132) on the oracle-backed ops."""
import numpy as np
import pytest
import torch as th

from sbmc_b200 import functions as funcs
from sbmc_b200 import interfaces, losses, models
from tests import kats

VAL, VAL2, EPS = 0.34, 0.7, 1e-2
T, T2 = VAL / (1 + VAL), VAL2 / (1 + VAL2)
CASES = [
    (losses.RelativeMSE, (VAL - VAL2) ** 2 / (VAL ** 2 + EPS) * 0.5),
    (losses.SMAPE, (VAL2 - VAL) / (VAL + VAL2 + EPS)),
    (losses.TonemappedMSE, (T - T2) ** 2 * 0.5),
    (losses.TonemappedRelativeMSE, (T - T2) ** 2 / (T ** 2 + EPS) * 0.5),
]


@pytest.mark.parametrize("cls,target", CASES, ids=[c[0].__name__ for c in CASES])
def test_single_pixel_values(cls, target):
    fn = cls(eps=EPS)
    sz = [1, 3, 4, 5]
    n = int(np.prod(sz))
    im, ref = th.zeros(*sz), th.zeros(*sz)
    assert abs(fn(im, ref).item()) < 1e-7
    for dx in range(n):
        ref.zero_(); im.zero_()
        ref.view(-1)[dx] = VAL
        im.view(-1)[dx] = VAL2
        assert abs(fn(im, ref).item() - target / n) < 1e-4


def test_training_interface_step(monkeypatch):
    KW, S2G = kats.oracle_functions()           # CPU stand-ins for the custom ops
    monkeypatch.setattr(funcs, "KernelWeighting", KW)
    monkeypatch.setattr(funcs, "Scatter2Gather", S2G)
    th.manual_seed(0)
    net = models.Multisteps(6, 2, width=8, embedding_width=8, ksize=3, nsteps=1)
    iface = interfaces.SampleBasedDenoiserInterface(net, lr=1e-3, cuda=False)
    batch = {"radiance": th.rand(2, 2, 3, 12, 12), "features": th.randn(2, 2, 6, 12, 12),
             "global_features": th.randn(2, 2, 1, 1), "target_image": th.rand(2, 3, 12, 12)}
    before = [p.detach().clone() for p in net.parameters()]
    first = iface.backward(batch, iface.forward(batch))
    assert np.isfinite(first["loss"]) and np.isfinite(first["rmse"])
    assert any(not th.equal(a, b) for a, b in zip(before, net.parameters()))
    for _ in range(5):
        last = iface.backward(batch, iface.forward(batch))
    assert last["loss"] < first["loss"]                  # Adam makes progress
    running = iface.update_validation(batch, iface.forward(batch), iface.init_validation())
    assert running["n"] == 2 and np.isfinite(running["loss"])
