"""Generates module / model level golden fixtures by RUNNING THE REFERENCE'S OWN
PYTHON CODE (sbmc/functions.py, sbmc/modules.py, sbmc/models.py, imported
unmodified from /root/reference) in this container:

    python tests/golden/make_model_golden.py        # writes tests/golden/model_golden.pt

What is stubbed, because it is absent from the image:
  * `sbmc.halide_ops` (the Halide-generated native module)  -> the CPU oracle's six
    `*_cpu_float32` entry points (same names / argument order, oracle/__init__.py);
  * `ttools` (torch-tools 0.0.36)  -> get_logger and image_operators.crop_like from
    sbmc_b200/_compat.py.
Everything else -- the autograd Functions, ConvChain / Autoencoder / KernelApply /
ProgressiveKernelApply, Multisteps (train and eval branches, including the CPU
staging and the sample-major global-feature tiling) and KPCN -- is reference code.
The fixtures hold seeded inputs, the reference modules' state dicts and their
outputs / gradients; tests/test_reference_fixtures.py loads the state dicts into
sbmc_b200's modules and must reproduce the outputs (CPU: oracle-backed ops; GPU:
the sm_100a kernels).
"""
import importlib.util
import os
import sys
import types

import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"


def import_reference():
    import oracle
    from sbmc_b200 import _compat
    ttools = types.ModuleType("ttools")
    ttools.get_logger = _compat.get_logger
    ttools.ModelInterface = object          # base class of sbmc/interfaces.py:35
    tmods = types.ModuleType("ttools.modules")
    timg = types.ModuleType("ttools.modules.image_operators")
    timg.crop_like = _compat.crop_like
    tmods.image_operators = timg
    ttools.modules = tmods
    sys.modules.update({"ttools": ttools, "ttools.modules": tmods,
                        "ttools.modules.image_operators": timg})
    pkg = types.ModuleType("sbmc")
    pkg.__path__ = [os.path.join(REFERENCE, "sbmc")]
    sys.modules["sbmc"] = pkg
    hops = types.ModuleType("sbmc.halide_ops")
    for name in ("scatter2gather", "kernel_weighting", "kernel_weighting_grad"):
        fn = getattr(oracle, name + "_cpu_float32")
        setattr(hops, name + "_cpu_float32", fn)
    sys.modules["sbmc.halide_ops"] = hops
    pkg.halide_ops = hops
    mods = {}
    for name in ("functions", "modules", "models", "losses", "interfaces"):
        spec = importlib.util.spec_from_file_location(
            "sbmc." + name, os.path.join(REFERENCE, "sbmc", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["sbmc." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        mods[name] = mod
    return mods


def randomize(module, gen):
    """Non-trivial weight-norm gains and biases (the reference zeroes the biases)."""
    with th.no_grad():
        for p in module.parameters():
            p.add_(0.1 * th.randn(p.shape, generator=gen))


def main():
    ref = import_reference()
    g = th.Generator().manual_seed(1234)
    out = {}

    # -- KernelApply / ProgressiveKernelApply ------------------------------------
    bs, c, h, w, k = 2, 3, 14, 18, 5
    data = th.rand(bs, 3, c, h, w, generator=g)
    logits = 2 * th.randn(bs, 3, k * k, h, w, generator=g)
    case = {"data": data, "logits": logits}
    for splat in (True, False):
        for softmax in (True, False):
            o, s = ref["modules"].KernelApply(softmax=softmax, splat=splat)(
                data[:, 0].contiguous(), logits[:, 0].clone())
            case["kernel_apply_splat%d_softmax%d" % (splat, softmax)] = (o, s)
        mod = ref["modules"].ProgressiveKernelApply(splat=splat)
        state = (None, None, None)
        for sp in range(3):
            state = mod(data[:, sp].contiguous(), logits[:, sp].clone(), *state)
        case["progressive_splat%d" % splat] = state
    out["kernel_apply"] = case

    # -- Multisteps: eval and train branches, gradients ----------------------------
    th.manual_seed(7)
    net = ref["models"].Multisteps(6, 2, width=8, embedding_width=8, ksize=5, nsteps=2)
    randomize(net, g)
    bs, spp, h, w = 2, 3, 24, 20
    samples = {"radiance": th.rand(bs, spp, 3, h, w, generator=g),
               "features": th.randn(bs, spp, 6, h, w, generator=g),
               "global_features": th.randn(bs, 2, 1, 1, generator=g)}
    net.eval()
    with th.no_grad():
        out_eval = net({k_: v.clone() for k_, v in samples.items()})["radiance"]
    net.train()
    out_train = net({k_: v.clone() for k_, v in samples.items()})["radiance"]
    proj = th.randn(out_train.shape, generator=g)
    (out_train * proj).sum().backward()
    grads = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    out["multisteps"] = {"ctor": dict(n_features=6, n_global_features=2, width=8,
                                      embedding_width=8, ksize=5, nsteps=2),
                         "state_dict": {k_: v.clone() for k_, v in net.state_dict().items()},
                         "samples": samples, "eval": out_eval, "train": out_train.detach(),
                         "proj": proj, "grads": grads}

    # -- gather-kernel ablation of Multisteps and KPCN -----------------------------
    th.manual_seed(8)
    net = ref["models"].Multisteps(6, 2, width=8, embedding_width=8, ksize=3, nsteps=1,
                                   splat=False)
    randomize(net, g)
    net.eval()
    with th.no_grad():
        o = net({k_: v.clone() for k_, v in samples.items()})["radiance"]
    out["multisteps_gather"] = {"ctor": dict(n_features=6, n_global_features=2, width=8,
                                             embedding_width=8, ksize=3, nsteps=1, splat=False),
                                "state_dict": {k_: v.clone() for k_, v in net.state_dict().items()},
                                "eval": o}
    th.manual_seed(9)
    kpcn = ref["models"].KPCN(4, ksize=5, depth=3, width=8)
    randomize(kpcn, g)
    kpcn.eval()
    kdata = {"kpcn_diffuse_in": th.randn(2, 4, 32, 36, generator=g),
             "kpcn_specular_in": th.randn(2, 4, 32, 36, generator=g),
             "kpcn_diffuse_buffer": th.rand(2, 3, 32, 36, generator=g),
             "kpcn_specular_buffer": th.rand(2, 3, 32, 36, generator=g),
             "kpcn_albedo": th.rand(2, 3, 32, 36, generator=g)}
    with th.no_grad():
        ko = kpcn({k_: v.clone() for k_, v in kdata.items()})
    out["kpcn"] = {"ctor": dict(n_in=4, ksize=5, depth=3, width=8),
                   "state_dict": {k_: v.clone() for k_, v in kpcn.state_dict().items()},
                   "data": kdata, "out": {k_: v for k_, v in ko.items()}}

    # -- two optimisation steps through the reference's training interface ------------
    th.manual_seed(10)
    net = ref["models"].Multisteps(6, 2, width=8, embedding_width=8, ksize=3, nsteps=1)
    randomize(net, g)
    init = {k_: v.clone() for k_, v in net.state_dict().items()}
    iface = ref["interfaces"].SampleBasedDenoiserInterface(net, lr=1e-3, cuda=False)
    batch = {"radiance": th.rand(2, 2, 3, 16, 16, generator=g),
             "features": th.randn(2, 2, 6, 16, 16, generator=g),
             "global_features": th.randn(2, 2, 1, 1, generator=g),
             "target_image": th.rand(2, 3, 16, 16, generator=g)}
    steps = []
    for _ in range(2):
        b = {k_: v.clone() for k_, v in batch.items()}
        steps.append(iface.backward(b, iface.forward(b)))
    out["train_steps"] = {"ctor": dict(n_features=6, n_global_features=2, width=8,
                                       embedding_width=8, ksize=3, nsteps=1),
                          "init": init, "batch": batch, "steps": steps, "lr": 1e-3,
                          "final": {k_: v.clone() for k_, v in net.state_dict().items()}}

    # -- parameter names / shapes of the full-size models (checkpoint compatibility) ---
    full = ref["models"].Multisteps(93, 3)
    out["multisteps_93_3_keys"] = {k_: tuple(v.shape) for k_, v in full.state_dict().items()}
    full = ref["models"].KPCN(27)
    out["kpcn_27_keys"] = {k_: tuple(v.shape) for k_, v in full.state_dict().items()}

    path = os.path.join(HERE, "model_golden.pt")
    th.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
