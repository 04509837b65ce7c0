"""Fused gradient clipping + Adam (sbmc/interfaces.py:78-106).

CPU: the device arithmetic compiled for the host (tests/native/optim_emul.cpp) and
FusedAdam's table building against torch.optim.Adam + clip_grad_norm_; GPU: the
kernels behind sbmc_b200.optim.FusedAdam against the same.  Tolerance: the update
is a handful of fp32 operations per element evaluated in torch's order up to FMA
contraction -> 1e-6 relative on parameters and moments, 1e-6 on the norm; with
clipping active the coefficient inherits the norm's 1e-6 (a 200 K-term fp32 sum
in another order), so moments are compared at 5e-6.
"""
import ctypes
import os
import subprocess

import pytest
import torch as th

from sbmc_b200 import _lib, optim

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(3,), (1,), (17, 5), (70000,), (65536,), (2, 65537), (128, 3, 3, 3)]


def emul():
    src = os.path.join(HERE, "native", "optim_emul.cpp")
    out = os.path.join(HERE, "native", "liboptim_emul.so")
    dep = os.path.join(HERE, "..", "sbmc_b200", "csrc", "optim_body.cuh")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(dep)) > os.path.getmtime(out):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I", "/usr/local/cuda/include", "-o", out, src])
    lib = ctypes.CDLL(out)
    vp, i64, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
    lib.emul_grad_norm.argtypes = [vp, vp, i64, vp, f32, vp]
    lib.emul_adam.argtypes = [vp, vp, i64, vp] + [ctypes.c_double] * 6
    return lib


def make_params(seed, device="cpu", scale=1.0):
    g = th.Generator().manual_seed(seed)
    return [th.nn.Parameter((scale * th.randn(*s, generator=g)).to(device)) for s in SHAPES]


def set_grads(params, seed, scale):
    g = th.Generator().manual_seed(seed)
    for p in params:
        p.grad = (scale * th.randn(*p.shape, generator=g)).to(p.device)


def close(a, b, what, rtol=1e-6):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= rtol * max(ref, 1e-3), "%s: %.3g vs scale %.3g" % (what, err, ref)


@pytest.mark.parametrize("max_norm", [None, 1000.0, 0.5])
def test_emulated_kernels_match_torch_adam(max_norm):
    lib = emul()
    mine, ref = make_params(0), make_params(0)
    opt = th.optim.Adam(ref, lr=1e-2)
    state = [(th.zeros_like(p), th.zeros_like(p)) for p in mine]
    vp = ctypes.c_void_p
    for step in range(1, 5):
        set_grads(mine, step, 3.0)
        set_grads(ref, step, 3.0)
        rows = [(p.data, p.grad, m, v) for p, (m, v) in zip(mine, state)]
        tensors, chunks, nchunks = optim.FusedAdam._tables(rows, th.device("cpu"))
        assert nchunks == sum((p.numel() + 65535) // 65536 for p in mine)
        coef = None
        if max_norm is not None:
            partial = th.empty(nchunks)
            nc = th.empty(2)
            lib.emul_grad_norm(tensors.data_ptr(), chunks.data_ptr(), nchunks, partial.data_ptr(),
                               max_norm, nc.data_ptr())
            want = th.nn.utils.clip_grad_norm_(ref, max_norm)
            assert nc[0].item() == pytest.approx(want.item(), rel=1e-6)
            assert nc[1].item() == pytest.approx(min(1.0, max_norm / (want.item() + 1e-6)), rel=1e-6)
            coef = nc.data_ptr() + 4
        lib.emul_adam(tensors.data_ptr(), chunks.data_ptr(), nchunks, vp(coef) if coef else None,
                      1e-2, 0.9, 0.999, 1e-8, 1 - 0.9 ** step, (1 - 0.999 ** step) ** 0.5)
        opt.step()
        tol = 1e-6 if max_norm is None else 5e-6
        for i, (p, q) in enumerate(zip(mine, ref)):
            close(p, q, "param %d step %d" % (i, step), tol)
            close(p.grad, q.grad, "clipped grad %d" % i, tol)
            close(state[i][0], opt.state[q]["exp_avg"], "exp_avg %d" % i, tol)
            close(state[i][1], opt.state[q]["exp_avg_sq"], "exp_avg_sq %d" % i, tol)


def test_tables_cover_every_element_once():
    rows = [(p.data, p.data, p.data, p.data) for p in make_params(1)]
    tensors, chunks, nchunks = optim.FusedAdam._tables(rows, th.device("cpu"))
    assert tensors.shape == (len(SHAPES), 5) and chunks.shape == (nchunks, 2)
    covered = [0] * len(rows)
    for t, start in chunks.tolist():
        n = int(tensors[t, 4])
        assert start % 65536 == 0 and start < n
        covered[t] += min(65536, n - start)
    assert covered == [r[0].numel() for r in rows]
    assert tensors[3, 0].item() == rows[3][0].data_ptr()


def test_fused_adam_refuses_host_tensors_and_bad_hyperparameters():
    p = th.nn.Parameter(th.zeros(4))
    p.grad = th.ones(4)
    with pytest.raises(_lib.SbmcB200Error):
        optim.FusedAdam([p]).step()
    with pytest.raises(ValueError):
        optim.FusedAdam([p], lr=-1.0)
    assert optim.FusedAdam([th.nn.Parameter(th.zeros(2))]).step() is None     # no gradients: no-op


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    assert lib.sbmc_multi_tensor_adam_f32(None, None, 0, None, 1e-3, 0.9, 0.999, 1e-8, 0.1, 0.03,
                                          None) == 0
    assert lib.sbmc_multi_tensor_adam_f32(None, None, 3, None, 1e-3, 0.9, 0.999, 1e-8, 0.1, 0.03,
                                          None) == -1
    assert lib.sbmc_multi_tensor_grad_norm_f32(None, None, 1, None, 1.0, None, None) == -1


# ------------------------------------------------------------------ FusedAdam, two backends
class EmulBackend(object):
    """Stands in for sbmc_b200.optim._CudaBackend: same Python, kernels built for the host."""
    device_type = "cpu"

    def __init__(self):
        self.lib = emul()

    def scope(self, dev):
        import contextlib
        return contextlib.nullcontext()

    def grad_norm(self, dev, tensors, chunks, nchunks, partial, max_norm, out):
        self.lib.emul_grad_norm(tensors.data_ptr(), chunks.data_ptr(), nchunks, partial.data_ptr(),
                                max_norm, out.data_ptr())

    def adam(self, dev, tensors, chunks, nchunks, coef_ptr, *scalars):
        self.lib.emul_adam(tensors.data_ptr(), chunks.data_ptr(), nchunks, coef_ptr, *scalars)


@pytest.fixture(params=["host-emulation", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    if request.param == "host-emulation":
        monkeypatch.setattr(optim, "_backend", EmulBackend)
        return "cpu"
    return "cuda"


@pytest.mark.parametrize("max_norm", [None, 1000.0, 0.5])
def test_fused_adam_matches_torch(max_norm, backend):
    mine, ref = make_params(0, backend), make_params(0, backend)
    fused = optim.FusedAdam(mine, lr=1e-2)
    opt = th.optim.Adam(ref, lr=1e-2)
    for step in range(1, 6):
        set_grads(mine, step, 3.0)
        set_grads(ref, step, 3.0)
        before = _lib.launch_count()
        versions = [p._version for p in mine]
        fused.step(max_norm=max_norm)
        # the kernels write through raw pointers; caches keyed on the autograd version
        # counter (conv1x1.prepare, unet_fast) must still see the update
        assert all(p._version > v for p, v in zip(mine, versions))
        if backend == "cuda":
            assert _lib.launch_count() - before == (3 if max_norm is not None else 1)
        if max_norm is not None:
            want = th.nn.utils.clip_grad_norm_(ref, max_norm)
            assert fused.last_grad_norm[0].item() == pytest.approx(want.item(), rel=1e-6)
        opt.step()
        tol = 1e-6 if max_norm is None else 5e-6
        for i, (p, q) in enumerate(zip(mine, ref)):
            close(p, q, "param %d step %d" % (i, step), tol)
            close(p.grad, q.grad, "clipped grad %d" % i, tol)
            close(fused.state[p]["exp_avg_sq"], opt.state[q]["exp_avg_sq"], "exp_avg_sq %d" % i,
                  tol)
    # state dicts interchange with torch.optim.Adam
    other = th.optim.Adam(make_params(0, backend), lr=1e-2)
    other.load_state_dict(fused.state_dict())
    assert float(other.state[other.param_groups[0]["params"][0]]["step"]) == 5.0
    # parameters that join later (no gradient so far) start their own step count
    late = th.nn.Parameter(th.ones(5, device=backend))
    fused.add_param_group({"params": [late]})
    set_grads(mine + [late], 9, 1.0)
    fused.step(max_norm=max_norm)
    assert float(fused.state[late]["step"]) == 1.0 and float(fused.state[mine[0]]["step"]) == 6.0
    one = th.nn.Parameter(th.ones(5, device=backend))
    one.grad = late.grad.clone() if max_norm is None else None
    if max_norm is None:
        solo = th.optim.Adam([one], lr=1e-2)
        solo.step()
        close(late, one, "late parameter's first step")


@pytest.mark.gpu
def test_gpu_interface_with_fused_optimizer_follows_the_eager_one():
    from sbmc_b200 import interfaces, models

    def run(fused):
        th.manual_seed(0)
        model = models.Multisteps(8, 3, ksize=3, nsteps=1, width=16, embedding_width=16)
        iface = interfaces.SampleBasedDenoiserInterface(model, lr=1e-3, cuda=True,
                                                        fused_optimizer=fused)
        g = th.Generator().manual_seed(1)
        out = []
        for _ in range(3):
            batch = {"radiance": th.rand(2, 2, 3, 16, 16, generator=g),
                     "features": th.rand(2, 2, 8, 16, 16, generator=g),
                     "global_features": th.rand(2, 3, 1, 1, generator=g),
                     "target_image": th.rand(2, 3, 16, 16, generator=g)}
            out.append(iface.backward(batch, iface.forward(batch))["loss"])
        return out, [p.detach().clone() for p in model.parameters()]

    loss_a, params_a = run(False)
    loss_b, params_b = run(True)
    assert loss_a == pytest.approx(loss_b, rel=1e-4)
    for a, b in zip(params_a, params_b):
        close(b, a, "parameter", rtol=1e-4)
