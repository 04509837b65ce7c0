"""Host logic of the mixed-precision training pipeline that needs no GPU: which models /
shapes it serves, the order in which the U-net's convolutions are planned, and the
single-call training step of the interface."""
import torch as th

from sbmc_b200 import interfaces, models, modules, train_pipeline as P


def test_unet_plan_follows_the_module_order():
    net = modules.Autoencoder(128, 128, num_levels=3, increase_factor=2.0, num_convs=3, width=128,
                              ksize=3, output_type="leaky_relu", pooling="max")
    plan, specs = P.unet_specs(net)
    assert [(len(l), None if r is None else len(r)) for l, r in plan] == [(3, 3), (3, 3), (3, None)]
    # level by level, left then right; channel counts of the reference's U-net (modules.py:221-243)
    shapes = [(c.in_channels, c.out_channels) for c, _, _ in specs]
    assert shapes == [(128, 128), (128, 128), (128, 128), (384, 128), (128, 128), (128, 128),
                      (128, 256), (256, 256), (256, 256), (768, 256), (256, 256), (256, 256),
                      (256, 512), (512, 512), (512, 512)]
    # activations: ReLU inside, LeakyReLU on the finest level's output (output_type)
    assert [a for _, a in plan[0][1]] == [1, 1, 2] and [a for _, a in plan[2][0]] == [1, 1, 1]
    assert all(hasattr(c, "weight_v") for c, _, _ in specs)


def test_chain_specs_pad_layer_one_and_the_prediction():
    reg = modules.ConvChain(256, 441, depth=3, width=128, ksize=1, activation="leaky_relu",
                            pad=False, output_type="linear")
    (c1, p1, _), (c2, _, _), (c3, _, q3) = P.chain_specs(reg, 256, 512)
    assert (c1.in_channels, p1, c3.out_channels, q3) == (256, 256, 441, 512)
    assert P._chain_act(reg) == 2
    emb = modules.ConvChain(96, 128, width=128, depth=3, ksize=1, pad=False)
    assert P._chain_act(emb) == 1 and P.chain_specs(emb, 128)[0][1] == 128


def test_supported_shapes_and_cpu_models():
    net = models.Multisteps(12, 3, ksize=5, nsteps=2)
    # parameters on the CPU: the pipeline (device kernels only) does not serve the model
    assert not P.supported(net, 12, 3, 32, 48)
    wide = models.Multisteps(12, 3, ksize=5, nsteps=1, width=64, embedding_width=64)
    assert not P.supported(wide, 12, 3, 32, 48)
    gather = models.Multisteps(12, 3, ksize=5, nsteps=1, splat=False)
    assert not P.supported(gather, 12, 3, 32, 48)


def test_cuda_graph_needs_cuda_and_the_fused_optimizer():
    net = models.Multisteps(12, 3, ksize=3, nsteps=1)
    for kwargs in (dict(cuda=False, fused_optimizer=True), dict(cuda=True, fused_optimizer=False)):
        try:
            interfaces.SampleBasedDenoiserInterface(net, cuda_graph=True, **kwargs)
        except ValueError:
            continue
        except (RuntimeError, AssertionError):      # .cuda() without a device
            continue
        raise AssertionError("cuda_graph accepted %s" % kwargs)
