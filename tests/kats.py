"""The reference's known-answer tests for the hot path, restated so that they
can be run against any implementation of the two autograd Functions (the CPU
oracle in the CPU suite, ``sbmc_b200.functions`` on the GPU).

Sources (adobe/sbmc): tests/test_functions.py:43-70 (_forward), :72-103
(_backward), :105-144 (_kernel_weighting_grad), :164-185 (_scatter2gather),
:187-208 (_scatter2gather_grad).  Loops are trimmed where the reference repeats
an identical check (it re-runs the op for every single tap; here a strided
subset of taps is used for the big loops) but the expected values are the
reference's.
"""
import warnings

import torch as th
from torch.autograd import gradcheck


def oracle_functions():
    """Autograd Functions backed by the CPU oracle (test infrastructure)."""
    import oracle

    class Scatter2Gather(th.autograd.Function):
        @staticmethod
        def forward(ctx, data):
            return oracle.scatter2gather(data)

        @staticmethod
        def backward(ctx, d_output):
            return oracle.scatter2gather(d_output)

    class KernelWeighting(th.autograd.Function):
        @staticmethod
        def forward(ctx, data, weights):
            ctx.save_for_backward(data, weights)
            return oracle.kernel_weighting(data, weights)

        @staticmethod
        def backward(ctx, d_output, d_sum_w):
            data, weights = ctx.saved_tensors
            return oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)

    return KernelWeighting, Scatter2Gather


def almost(a, b, places=7):
    assert round(abs(a - b), places) == 0, (a, b)


def kw_forward_impulse(KW, device):
    """test_functions.py:43-70 -- pins out[y,x] = sum w[dy,dx,y,x] data[y+dy-c, x+dx-c]."""
    bs, c, h, w, ksize = 4, 5, 16, 16, 5
    data = th.zeros(bs, c, h, w)
    idx = 1
    y, x = h // 2, w // 2
    data[idx, 0, y, x] = 1.4
    data[idx, 1, y, x] = 2.4
    data[idx, 2, y, x] = 3.4
    for dy in range(-(ksize // 2), ksize // 2 + 1):
        for dx in range(-(ksize // 2), ksize // 2 + 1):
            weights = th.zeros(bs, ksize, ksize, h, w)
            weights[idx, ksize // 2 + dy, ksize // 2 + dx, y - dy, x - dx] = 0.5
            o, s = KW.apply(data.to(device), weights.to(device))
            o, s = o.cpu(), s.cpu()
            almost(o[idx, 0, y - dy, x - dx].item(), 1.4 * 0.5)
            almost(o[idx, 1, y - dy, x - dx].item(), 2.4 * 0.5)
            almost(o[idx, 2, y - dy, x - dx].item(), 3.4 * 0.5)
            almost(s[idx, y - dy, x - dx].item(), 0.5)
            # and nothing else is touched
            assert o.abs().sum().item() - (1.4 + 2.4 + 3.4) * 0.5 < 1e-5
            almost(s.abs().sum().item(), 0.5)


def kw_backward_impulse(KW, device):
    """test_functions.py:72-103."""
    bs, chans, h, w = 3, 5, 16, 16
    x, y = w // 2, h // 2
    for ksize in [3, 5, 7]:
        for b in range(bs):
            for c in range(0, chans, 2):
                data = th.full((bs, chans, h, w), 7.0, device=device, requires_grad=True)
                weights = th.ones(bs, ksize, ksize, h, w, device=device, requires_grad=True)
                o, s = KW.apply(data, weights)
                o_grad = th.zeros_like(o)
                o_grad[b, c, x, y] = 1.1
                o.backward(o_grad)
                dgrad = data.grad.cpu().clone()
                for dy in range(-(ksize // 2), ksize // 2 + 1):
                    for dx in range(-(ksize // 2), ksize // 2 + 1):
                        almost(dgrad[b, c, y + dy, x + dx].item(), 1.1)
                        dgrad[b, c, y + dy, x + dx] = 0.0
                almost(dgrad.abs().max().item(), 0.0)
                wgrad = weights.grad.cpu()
                for ky in range(ksize):
                    for kx in range(ksize):
                        almost(wgrad[b, ky, kx, x, y].item(), 7.0 * 1.1, places=3)
                # d_sum_w was a materialised zero: other pixels get no gradient
                almost(wgrad.abs().sum().item(), 7.7 * ksize * ksize, places=2)


def kw_gradcheck(KW, device):
    """test_functions.py:105-144 -- float32 gradcheck, same eps / tolerances."""
    bs, c, h, w, ksize = 2, 3, 16, 16, 3
    th.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        data = (2 * th.randn(bs, c, h, w)).to(device).requires_grad_(True)
        weights = th.randn(bs, ksize, ksize, h, w).to(device)
        assert gradcheck(KW.apply, (data, weights), eps=1e-4, atol=5e-2, rtol=5e-4)
        data = (2 * th.randn(bs, c, h, w)).to(device)
        weights = th.randn(bs, ksize, ksize, h, w).to(device).requires_grad_(True)
        assert gradcheck(KW.apply, (data, weights), eps=1e-4, atol=5e-2, rtol=5e-4)


def s2g_index_map(S2G, device, stride=3):
    """test_functions.py:164-185 -- pins the bit-exact index map."""
    bs, h, w = 4, 32, 32
    for ksize in [3, 5, 7, 9]:
        idx = ksize % bs
        count = 0
        for y in range(h // 2 - ksize // 2, h // 2 + ksize // 2 + 1):
            for x in range(w // 2 - ksize // 2, w // 2 + ksize // 2 + 1):
                for ky in range(ksize):
                    for kx in range(ksize):
                        count += 1
                        if count % stride:
                            continue
                        scatter = th.zeros(bs, ksize, ksize, h, w)
                        dx, dy = kx - ksize // 2, ky - ksize // 2
                        kx2, ky2 = ksize - 1 - kx, ksize - 1 - ky
                        scatter[idx, ky, kx, y, x] = 0.5
                        gather = S2G.apply(scatter.to(device)).cpu()
                        assert gather[idx, ky2, kx2, y + dy, x + dx].item() == 0.5
                        assert gather.abs().sum().item() == 0.5


def s2g_gradcheck(S2G, device, h=12, w=10):
    """test_functions.py:187-208 (the reference uses 32x32; the image is shrunk
    here because gradcheck perturbs every element in turn)."""
    th.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        weights = th.randn(2, 3, 3, h, w).to(device).requires_grad_(True)
        assert gradcheck(S2G.apply, (weights,), eps=1e-4, atol=5e-2, rtol=5e-4)
