"""CPU suite: the C-ABI library loads without a GPU and exports every symbol
that include/sbmc_b200.h declares; argument validation happens before any CUDA
call; the Python drop-in module fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch as th

from sbmc_b200 import _lib, halide_ops

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "include", "sbmc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"SBMC_API\s+[\w\s\*]+?\b(sbmc_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for want in ("sbmc_scatter2gather_f32", "sbmc_kernel_weighting_fwd_f32",
                 "sbmc_kernel_weighting_bwd_f32", "sbmc_scatter2gather_host_f32",
                 "sbmc_kernel_weighting_fwd_host_f32",
                 "sbmc_kernel_weighting_bwd_host_f32"):
        assert want in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export: " + name


def test_python_binding_covers_every_declared_symbol():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.sbmc_b200_version() >= 100


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    # negative sizes / zero kernel -> SBMC_EINVAL, with a message
    rc = lib.sbmc_kernel_weighting_fwd_f32(None, None, None, None, -1, 3, 4, 4, 3, 3, None)
    assert rc == -1
    assert b"invalid shape" in lib.sbmc_b200_last_error()
    rc = lib.sbmc_scatter2gather_f32(None, None, 1, 0, 3, 4, 4, None)
    assert rc == -1
    # empty problems succeed without touching the device
    assert lib.sbmc_kernel_weighting_fwd_f32(None, None, None, None, 0, 3, 4, 4, 3, 3, None) == 0
    assert lib.sbmc_scatter2gather_f32(None, None, 2, 3, 3, 0, 4, None) == 0
    # null pointers on a non-empty problem
    rc = lib.sbmc_kernel_weighting_fwd_f32(None, None, None, None, 1, 3, 4, 4, 3, 3, None)
    assert rc == -1
    with pytest.raises(_lib.SbmcB200Error):
        _lib.check(rc, "kernel_weighting")


def test_plain_c_client_builds_links_and_validates(tmp_path):
    """include/sbmc_b200.h is a C header: a C99 client compiles with -pedantic,
    links every declared entry point and gets the documented status codes."""
    import shutil
    import subprocess
    _lib.load()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cabi_client.c")
    text = open(src).read()
    for name in declared_symbols():
        assert name in text, "cabi_client.c does not reference " + name
    exe = str(tmp_path / "cabi_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           "-I", os.path.join(root, "include"), src,
                           "-o", exe, "-L", libdir, "-l:libsbmc_b200.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "failures: 0" in out.stdout


def test_drop_in_module_has_the_reference_names():
    # reference setup.py:65-84
    for op in ("scatter2gather", "kernel_weighting", "kernel_weighting_grad"):
        for dev in ("cpu", "cuda"):
            assert callable(getattr(halide_ops, "%s_%s_float32" % (op, dev)))


def test_drop_in_module_rejects_bad_tensors():
    w = th.zeros(1, 3, 3, 4, 4)
    with pytest.raises(RuntimeError, match="float32"):
        halide_ops.scatter2gather_cpu_float32(w.double(), w.double())
    with pytest.raises(RuntimeError, match="contiguous"):
        halide_ops.scatter2gather_cpu_float32(w.transpose(3, 4), w)
    with pytest.raises(RuntimeError, match="dimensions"):
        halide_ops.scatter2gather_cpu_float32(w[0], w[0])
    with pytest.raises(RuntimeError, match="CUDA"):
        halide_ops.scatter2gather_cuda_float32(w, w.clone())
    d = th.zeros(1, 3, 4, 4)
    with pytest.raises(RuntimeError, match="shape"):
        halide_ops.kernel_weighting_cpu_float32(d, w, d.clone(), th.zeros(1, 4, 5))


@pytest.mark.skipif(th.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Host tensors are streamed through a GPU; without one the op must raise."""
    import sbmc_b200.functions as funcs
    w = th.zeros(1, 3, 3, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU compute path"):
        funcs.Scatter2Gather.apply(w)
    with pytest.raises(RuntimeError, match="no CPU compute path"):
        funcs.KernelWeighting.apply(th.zeros(1, 3, 4, 4), w)


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "sbmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "sbmc_oracle" not in text, f


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "sbmc")),
                    reason="reference tree not present (GPU box)")
@pytest.mark.skipif(th.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_reference_functions_bind_to_the_drop_in(monkeypatch):
    """INTEGRATION.md option A: the reference's own sbmc/functions.py, unmodified,
    on top of sbmc_b200.halide_ops installed as `sbmc.halide_ops`.  Without a GPU
    the call must get through every argument / shape check of the drop-in (the
    reference passes its inputs then its resize_()d outputs, functions.py:53-59,
    91-98,105-114) and stop at the missing device."""
    import importlib.util
    import sys
    import types
    ttools = types.ModuleType("ttools")
    from sbmc_b200._compat import get_logger
    ttools.get_logger = get_logger
    pkg = types.ModuleType("sbmc")
    pkg.__path__ = [os.path.join(REFERENCE, "sbmc")]
    monkeypatch.setitem(sys.modules, "ttools", ttools)
    monkeypatch.setitem(sys.modules, "sbmc", pkg)
    monkeypatch.setitem(sys.modules, "sbmc.halide_ops", halide_ops)
    pkg.halide_ops = halide_ops
    spec = importlib.util.spec_from_file_location(
        "sbmc.functions", os.path.join(REFERENCE, "sbmc", "functions.py"))
    ref = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "sbmc.functions", ref)
    spec.loader.exec_module(ref)
    assert ref.ops is halide_ops
    with pytest.raises(RuntimeError, match="no CPU compute path"):
        ref.KernelWeighting.apply(th.zeros(2, 3, 8, 8), th.zeros(2, 5, 5, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU compute path"):
        ref.Scatter2Gather.apply(th.zeros(2, 5, 5, 8, 8))


def test_package_exports_the_reference_top_level_names():
    """sbmc/__init__.py re-exports its datasets, models and interface: the same
    names resolve on this package."""
    import sbmc_b200 as sbmc
    from sbmc_b200 import datasets, interfaces, models
    assert sbmc.TilesDataset is datasets.TilesDataset
    assert sbmc.FullImagesDataset is datasets.FullImagesDataset
    assert sbmc.MultiSampleCountDataset is datasets.MultiSampleCountDataset
    assert sbmc.Multisteps is models.Multisteps and sbmc.KPCN is models.KPCN
    assert sbmc.SampleBasedDenoiserInterface is interfaces.SampleBasedDenoiserInterface
    assert sbmc.TilesDataset.KPCN_MODE == "kpcn" and sbmc.TilesDataset.SBMC_MODE == "sbmc"
    assert "Multisteps" in dir(sbmc)
    assert sbmc.DenoisingDisplayCallback.__module__ == "sbmc_b200.callbacks"
    with pytest.raises(AttributeError):
        sbmc.NoSuchName
