"""sbmc_b200 -- B200-native kernel-splatting hot path of adobe/sbmc.

``functions`` / ``modules`` / ``models`` keep the reference's autograd-Function
and nn.Module API; ``halide_ops`` is the drop-in for the reference's native op
module; ``_lib`` binds the C ABI of libsbmc_b200.so (include/sbmc_b200.h).
"""
from . import _lib  # noqa: F401
from . import halide_ops  # noqa: F401
from . import functions  # noqa: F401

__version__ = "0.1.0"

# The names the reference package exports at top level (sbmc/__init__.py:19-23:
# `from .datasets import *`, `.models`, `.interfaces`), resolved on first use so
# that `import sbmc_b200 as sbmc; sbmc.Multisteps(...)` reads like the
# reference's scripts.
_EXPORTS = {
    "TilesDataset": "datasets", "FullImagesDataset": "datasets",
    "MultiSampleCountDataset": "datasets", "Multisteps": "models", "KPCN": "models",
    "SampleBasedDenoiserInterface": "interfaces", "DenoisingDisplayCallback": "callbacks",
}


def __getattr__(name):
    if name in _EXPORTS:
        import importlib
        value = getattr(importlib.import_module("." + _EXPORTS[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def __dir__():
    return sorted(list(globals()) + list(_EXPORTS))
