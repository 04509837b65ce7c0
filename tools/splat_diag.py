import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from sbmc_b200 import modules, _lib
from tests.test_modules import _softmax_splat_reference

for shape in [(1, 3, 30, 256, 21), (1, 3, 30, 128, 21), (1, 3, 64, 256, 21)]:
    bs, c, h, w, k = shape
    th.manual_seed(sum((1, 3, 30, 256, 21)))
    for spp in (1, 3):
        radiance = th.rand(bs, spp, c, h, w, device="cuda")
        logits = 4 * th.randn(bs, spp, k * k, h, w, device="cuda")
        for force in (0, 1):
            _lib.force_generic(force)
            fused = modules.ProgressiveKernelApply(splat=True)
            a = (None, None, None)
            with th.no_grad():
                for sp in range(spp):
                    a = fused(radiance[:, sp], logits[:, sp].clone(), *a)
            rr, rw, rm = _softmax_splat_reference(radiance.cpu(), logits.cpu(), k)
            err = ((a[1].cpu().double() - rw).abs() / rw)[0, 0]
            top = th.topk(err.flatten(), 5)
            locs = [(int(i) // w, int(i) % w, float(v)) for v, i in zip(top.values, top.indices)]
            print(shape, "spp", spp, "generic" if force else "tuned", "max %.2e mean %.2e" % (err.max(), err.mean()), locs)
        _lib.force_generic(0)
