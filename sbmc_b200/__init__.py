"""sbmc_b200 -- B200-native kernel-splatting hot path of adobe/sbmc.

``functions`` / ``modules`` / ``models`` keep the reference's autograd-Function
and nn.Module API; ``halide_ops`` is the drop-in for the reference's native op
module; ``_lib`` binds the C ABI of libsbmc_b200.so (include/sbmc_b200.h).
"""
from . import _lib  # noqa: F401
from . import halide_ops  # noqa: F401
from . import functions  # noqa: F401

__version__ = "0.1.0"
