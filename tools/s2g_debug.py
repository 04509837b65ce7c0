import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from sbmc_b200 import halide_ops
w = th.randn(1, 3, 3, 8, 8, device="cuda")
o = th.zeros_like(w)
halide_ops.scatter2gather_cuda_float32(w, o)
th.cuda.synchronize()
print("flags", os.environ.get("SBMC_S2G_DEBUG"), "ok", o.abs().sum().item())
