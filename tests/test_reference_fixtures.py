"""Module / model parity against fixtures produced by RUNNING THE REFERENCE'S OWN
sbmc/functions.py, modules.py and models.py (tests/golden/make_model_golden.py;
native ops = the CPU oracle, ttools stubbed).  The reference modules' state dicts
are loaded into sbmc_b200's modules (same parameter names) and the outputs and
gradients must be reproduced: on CPU with the oracle standing in for the custom
ops, on the GPU with the sm_100a kernels (composed and fused paths)."""
import os

import pytest
import torch as th

from sbmc_b200 import functions as funcs
from sbmc_b200 import models, modules
from tests import kats

PATH = os.path.join(os.path.dirname(__file__), "golden", "model_golden.pt")


@pytest.fixture(scope="module")
def golden():
    return th.load(PATH, map_location="cpu", weights_only=False)


@pytest.fixture
def oracle_ops(monkeypatch):
    KW, S2G = kats.oracle_functions()
    monkeypatch.setattr(funcs, "KernelWeighting", KW)
    monkeypatch.setattr(funcs, "Scatter2Gather", S2G)


def _close(a, b, rtol, atol):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape
    assert th.allclose(a, b, rtol=rtol, atol=atol), (a - b).abs().max().item()


def _check_kernel_apply(g, device, rtol, atol):
    case = g["kernel_apply"]
    data, logits = case["data"].to(device), case["logits"].to(device)
    for splat in (True, False):
        for softmax in (True, False):
            ro, rs = case["kernel_apply_splat%d_softmax%d" % (splat, softmax)]
            mod = modules.KernelApply(softmax=softmax, splat=splat)
            o, s = mod(data[:, 0].contiguous(), logits[:, 0].clone().requires_grad_(True))
            _close(o, ro, rtol, atol)
            _close(s, rs, rtol, atol)
            with th.no_grad():        # the inference (possibly fused) route
                o, s = mod(data[:, 0].contiguous(), logits[:, 0].clone())
            _close(o, ro, rtol, atol)
            _close(s, rs, rtol, atol)
        for fused in (True, False):
            mod = modules.ProgressiveKernelApply(splat=splat)
            mod.fused = fused
            state = (None, None, None)
            with th.no_grad():
                for sp in range(3):
                    state = mod(data[:, sp].contiguous(), logits[:, sp].clone(), *state)
            for got, want in zip(state, case["progressive_splat%d" % splat]):
                _close(got, want, rtol, atol)
            assert th.equal(state[2].cpu(), case["progressive_splat%d" % splat][2])  # max: exact


def _check_multisteps(g, device, rtol, atol, grad_rtol):
    case = g["multisteps"]
    net = models.Multisteps(**case["ctor"])
    net.load_state_dict(case["state_dict"], strict=True)       # reference parameter names
    net = net.to(device)
    samples = {k: v.to(device) for k, v in case["samples"].items()}
    net.eval()
    with th.no_grad():
        _close(net(dict(samples))["radiance"], case["eval"], rtol, atol)
    net.train()
    out = net(dict(samples))["radiance"]
    _close(out, case["train"], rtol, atol)
    (out * case["proj"].to(device)).sum().backward()
    for name, p in net.named_parameters():
        want = case["grads"][name]
        scale = want.abs().max().item() + 1e-12
        assert (p.grad.cpu() - want).abs().max().item() <= grad_rtol * scale, name
    case = g["multisteps_gather"]
    net = models.Multisteps(**case["ctor"])
    net.load_state_dict(case["state_dict"], strict=True)
    net = net.to(device).eval()
    with th.no_grad():
        _close(net(dict(samples))["radiance"], case["eval"], rtol, atol)


def _check_kpcn(g, device, rtol, atol):
    case = g["kpcn"]
    net = models.KPCN(**case["ctor"])
    net.load_state_dict(case["state_dict"], strict=True)
    net = net.to(device).eval()
    with th.no_grad():
        out = net({k: v.to(device) for k, v in case["data"].items()})
    for key, want in case["out"].items():
        _close(out[key], want, rtol, atol)


def test_kernel_apply_matches_reference_run(golden, oracle_ops):
    _check_kernel_apply(golden, "cpu", 1e-5, 1e-6)


def test_multisteps_matches_reference_run(golden, oracle_ops):
    _check_multisteps(golden, "cpu", 1e-5, 1e-6, 1e-4)


def test_kpcn_matches_reference_run(golden, oracle_ops):
    _check_kpcn(golden, "cpu", 1e-5, 1e-6)


def test_full_size_state_dict_names_match_reference(golden):
    """Same parameter names and shapes as the reference's Multisteps(93, 3) / KPCN(27)
    (so that the published checkpoints, Makefile:213, load with strict=True)."""
    mine = {k: tuple(v.shape) for k, v in models.Multisteps(93, 3).state_dict().items()}
    assert mine == golden["multisteps_93_3_keys"]
    mine = {k: tuple(v.shape) for k, v in models.KPCN(27).state_dict().items()}
    assert mine == golden["kpcn_27_keys"]


def test_training_steps_match_reference_run(golden, oracle_ops):
    """Two Adam steps through the reference's SampleBasedDenoiserInterface
    (sbmc/interfaces.py:62-106: loss, clip, Adam, rmse) reproduced by ours."""
    from sbmc_b200 import interfaces
    case = golden["train_steps"]
    net = models.Multisteps(**case["ctor"])
    net.load_state_dict(case["init"], strict=True)
    iface = interfaces.SampleBasedDenoiserInterface(net, lr=case["lr"], cuda=False)
    for want in case["steps"]:
        b = {k: v.clone() for k, v in case["batch"].items()}
        got = iface.backward(b, iface.forward(b))
        assert abs(got["loss"] - want["loss"]) <= 1e-6 * abs(want["loss"]) + 1e-9
        assert abs(got["rmse"] - want["rmse"]) <= 1e-6 * abs(want["rmse"]) + 1e-9
    for name, want in case["final"].items():
        _close(net.state_dict()[name], want, 1e-5, 1e-7)


@pytest.fixture
def fp32_convs(monkeypatch):
    monkeypatch.setattr(th.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(th.backends.cuda.matmul, "allow_tf32", False)


@pytest.mark.gpu
def test_kernel_apply_matches_reference_run_gpu(golden):
    _check_kernel_apply(golden, "cuda", 3e-5, 3e-6)


@pytest.mark.gpu
def test_multisteps_matches_reference_run_gpu(golden, fp32_convs):
    _check_multisteps(golden, "cuda", 1e-4, 1e-5, 2e-3)


@pytest.mark.gpu
def test_kpcn_matches_reference_run_gpu(golden, fp32_convs):
    _check_kpcn(golden, "cuda", 1e-4, 1e-5)
