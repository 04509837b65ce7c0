"""Runs the REFERENCE'S OWN TEST FILES, unmodified, from /root/reference/tests:

* target "reference": the reference's sbmc/functions.py + modules.py + losses.py on
  top of the CPU oracle as `sbmc.halide_ops` -- this pins the ORACLE against every
  known-answer test the reference holds for the path (tests/test_functions.py,
  tests/test_modules.py);
* target "sbmc_b200": the same test files with `sbmc.functions` / `sbmc.modules` /
  `sbmc.losses` resolved to THIS repo's drop-in modules (custom ops stood in for by
  the oracle on CPU) -- the reference's tests are the parity tests.

Only runs where the reference tree is mounted (the build container); the GPU
suite covers the same expectations through tests/kats.py and the committed
fixtures.  `ttools` (external, absent) is stubbed with get_logger.
"""
import importlib.util
import io
import os
import sys
import types
import unittest

import pytest

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REFERENCE, "tests")), reason="reference tree not mounted")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _install(monkeypatch, target):
    import oracle
    from sbmc_b200 import _compat
    ttools = types.ModuleType("ttools")
    ttools.get_logger = _compat.get_logger
    monkeypatch.setitem(sys.modules, "ttools", ttools)
    pkg = types.ModuleType("sbmc")
    pkg.__path__ = [os.path.join(REFERENCE, "sbmc")]
    monkeypatch.setitem(sys.modules, "sbmc", pkg)
    hops = types.ModuleType("sbmc.halide_ops")
    for op in ("scatter2gather", "kernel_weighting", "kernel_weighting_grad"):
        setattr(hops, op + "_cpu_float32", getattr(oracle, op + "_cpu_float32"))
    monkeypatch.setitem(sys.modules, "sbmc.halide_ops", hops)
    pkg.halide_ops = hops
    if target == "reference":
        for name in ("functions", "modules", "losses"):
            monkeypatch.delitem(sys.modules, "sbmc." + name, raising=False)
            mod = _load("sbmc." + name, os.path.join(REFERENCE, "sbmc", name + ".py"))
            monkeypatch.setitem(sys.modules, "sbmc." + name, mod)
            setattr(pkg, name, mod)
    else:
        from sbmc_b200 import functions, losses, modules
        from tests import kats
        KW, S2G = kats.oracle_functions()          # CPU stand-ins for the sm_100a ops
        monkeypatch.setattr(functions, "KernelWeighting", KW)
        monkeypatch.setattr(functions, "Scatter2Gather", S2G)
        for name, mod in (("functions", functions), ("modules", modules), ("losses", losses)):
            monkeypatch.setitem(sys.modules, "sbmc." + name, mod)
            setattr(pkg, name, mod)


def _run(test_file, monkeypatch, target):
    _install(monkeypatch, target)
    name = "_reference_%s_%s" % (os.path.basename(test_file)[:-3], target)
    mod = _load(name, os.path.join(REFERENCE, "tests", test_file))
    try:
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        stream = io.StringIO()
        result = unittest.TextTestRunner(stream=stream, verbosity=0).run(suite)
    finally:
        sys.modules.pop(name, None)
    assert result.testsRun > 0
    assert result.wasSuccessful(), stream.getvalue()
    return result.testsRun


@pytest.mark.parametrize("target", ["reference", "sbmc_b200"])
def test_reference_test_modules(monkeypatch, target):
    # ConvChain structure / errors, KernelApply and ProgressiveKernelApply spreads
    assert _run("test_modules.py", monkeypatch, target) == 3


@pytest.mark.parametrize("target", ["reference", "sbmc_b200"])
def test_reference_test_losses(monkeypatch, target):
    assert _run("test_losses.py", monkeypatch, target) == 4


@pytest.mark.parametrize("target", ["reference", "sbmc_b200"])
def test_reference_test_functions(monkeypatch, target):
    # impulse forward / backward, fp32 gradchecks, scatter2gather index map (CPU
    # variants; the *_cuda variants return early without a GPU, as in the reference)
    assert _run("test_functions.py", monkeypatch, target) == 10
