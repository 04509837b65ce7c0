"""Independent float64 restatement of the three ops (TEST INFRASTRUCTURE ONLY).

Written from the tap formulas, not from sbmc_oracle.c, with zero-padding and
shifted slices instead of per-pixel loops, so that the two restatements can be
checked against each other and against the reference's analytic tests.

Formulas (torch index order; c0 = (K-1)//2; zero outside the image):
  out[n,c,y,x]   = sum_{dy,dx} W[n,dy,dx,y,x] * D[n,c,y+dy-c0h,x+dx-c0w]
  sum_w[n,y,x]   = sum_{dy,dx} W[n,dy,dx,y,x]
      reference: src/kernel_weighting.cpp:45-60
  dD[n,c,y,x]    = sum_{ry,rx} W[n,KH-1-ry,KW-1-rx,y+ry-c0h,x+rx-c0w]
                               * dO[n,c,y+ry-c0h,x+rx-c0w]
  dW[n,dy,dx,y,x]= dSw[n,y,x] + sum_c D[n,c,y+dy-c0h,x+dx-c0w] * dO[n,c,y,x]
      reference: src/kernel_weighting.cpp:86-117
  G[n,dy,dx,y,x] = S[n,KH-1-dy,KW-1-dx,y+dy-c0h,x+dx-c0w]
      reference: src/scatter2gather.cpp:37-47
"""
import numpy as np


def _pad_hw(a, kh, kw):
    """Zero-pad the last two dims so that a_p[..., y+dy, x+dx] == a0(y+dy-c0h, x+dx-c0w)."""
    c0h, c0w = (kh - 1) // 2, (kw - 1) // 2
    pad = [(0, 0)] * (a.ndim - 2) + [(c0h, kh - 1 - c0h), (c0w, kw - 1 - c0w)]
    return np.pad(a, pad)


def kernel_weighting(data, weights):
    data = np.asarray(data, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    n, c, h, w = data.shape
    _, kh, kw, _, _ = weights.shape
    dp = _pad_hw(data, kh, kw)
    out = np.zeros((n, c, h, w))
    for dy in range(kh):
        for dx in range(kw):
            out += weights[:, dy, dx][:, None] * dp[:, :, dy:dy + h, dx:dx + w]
    sum_w = weights.sum(axis=(1, 2))
    return out, sum_w


def kernel_weighting_grad(data, weights, d_output, d_sum_w):
    data = np.asarray(data, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    d_output = np.asarray(d_output, dtype=np.float64)
    d_sum_w = np.asarray(d_sum_w, dtype=np.float64)
    n, c, h, w = data.shape
    _, kh, kw, _, _ = weights.shape
    dp = _pad_hw(data, kh, kw)
    op = _pad_hw(d_output, kh, kw)
    wp = _pad_hw(weights, kh, kw)
    d_data = np.zeros((n, c, h, w))
    d_weights = np.zeros((n, kh, kw, h, w))
    for ry in range(kh):
        for rx in range(kw):
            wsh = wp[:, kh - 1 - ry, kw - 1 - rx, ry:ry + h, rx:rx + w]
            d_data += wsh[:, None] * op[:, :, ry:ry + h, rx:rx + w]
            d_weights[:, ry, rx] = d_sum_w + (
                dp[:, :, ry:ry + h, rx:rx + w] * d_output).sum(axis=1)
    return d_data, d_weights


def scatter2gather(weights):
    weights = np.asarray(weights)
    n, kh, kw, h, w = weights.shape
    wp = _pad_hw(weights, kh, kw)
    out = np.zeros_like(weights)
    for dy in range(kh):
        for dx in range(kw):
            out[:, dy, dx] = wp[:, kh - 1 - dy, kw - 1 - dx, dy:dy + h, dx:dx + w]
    return out
