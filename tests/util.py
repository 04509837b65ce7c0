"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch as th

import oracle

# fp32 parity bar from BASELINE.json's north_star: 1e-5 relative.  A K*K-tap
# fp32 sum has a forward error bound proportional to sum_i |term_i|, so the
# element-wise check is  |got - ref| <= RTOL * (|ref| + sum|terms|)  (the
# condition-aware form of "1e-5 relative") and the norm-wise check is
# ||got - ref|| <= RTOL * ||ref||.
RTOL = 1e-5


def assert_close_sum(got, ref, mag, what, rtol=RTOL):
    """got/ref: tensors; mag: tensor of sum|terms| per element (same shape)."""
    got = got.detach().cpu().double()
    ref = ref.detach().cpu().double()
    mag = mag.detach().cpu().double()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert th.isfinite(got).all(), what + ": non-finite values"
    err = (got - ref).abs()
    bound = rtol * (ref.abs() + mag) + 1e-30
    worst = (err / bound).max().item() if err.numel() else 0.0
    assert worst <= 1.0, "%s: element-wise error %.3g x the 1e-5 bound" % (what, worst)
    nref = ref.norm().item()
    if nref > 0:
        rel = (got - ref).norm().item() / nref
        assert rel <= rtol, "%s: norm-wise relative error %.3g" % (what, rel)


def kw_magnitudes(data, weights, d_output, d_sum_w):
    """sum|terms| for every output of fwd and bwd, from the oracle on |inputs|."""
    ad, aw = data.abs(), weights.abs()
    mo, ms = oracle.kernel_weighting(ad, aw)
    mdd, mdw = oracle.kernel_weighting_grad(ad, aw, d_output.abs(), d_sum_w.abs())
    return mo, ms, mdd, mdw


def make_inputs(n, c, h, w, kh, kw, seed=0):
    """Synthetic inputs as in SURVEY.md section 8d (seeded, radiance-like data)."""
    g = th.Generator().manual_seed(seed)
    data = 2 * th.randn(n, c, h, w, generator=g)
    weights = th.randn(n, kh, kw, h, w, generator=g)
    d_output = th.randn(n, c, h, w, generator=g)
    d_sum_w = th.randn(n, h, w, generator=g)
    return data, weights, d_output, d_sum_w


def np64(t):
    return t.detach().cpu().numpy().astype(np.float64)
