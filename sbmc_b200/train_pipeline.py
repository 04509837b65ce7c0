"""Mixed-precision TRAINING pipeline of `Multisteps` (opt-in: `Multisteps.bf16_train`).

The reference trains in fp32 through cuDNN and several hundred eager autograd nodes
(sbmc/models.py:171-209 under sbmc/interfaces.py:78-106).  Here the train branch runs
on bf16 channels-innermost rows with fp32 accumulation, fp32 master weights and fp32
weight gradients, as a handful of autograd nodes whose forward AND backward are this
repo's kernels:

  EmbedStage    embedding_XX + `features.mean(1)`            (models.py:171-181)
                forward   3 tcgen05 GEMM layers (csrc/linear.cu; the broadcast U-net
                          output enters layer 1 as a second operand source instead of
                          `th.cat([features, propagated.repeat(...)])`) + one reduction
                          over the samples (csrc/train_ops.cu)
                backward  data gradients on the same GEMM kernel with the activation
                          derivative fused into its epilogue, weight + bias gradients on
                          the split-K MN-major tcgen05 kernel (csrc/wgrad.cu)
  UNetStage     propagation_XX (modules.Autoencoder)          (modules.py:248-320)
                forward   conv3x3 implicit GEMM + own max-pool / upsample+concat kernels
                backward  data gradients = the same conv kernel on flipped / transposed
                          weights with the previous layer's activation derivative in the
                          epilogue; max-pool / skip / upsample gradients one kernel each;
                          weight gradients on the split-K tcgen05 kernel (csrc/wgrad.cu:
                          the three dx taps of a kernel row share one halo box)
  RegressStage  kernel_regressor                             (models.py:195-199)
                the last layer writes fp32 logits as channel planes per sample -- the
                layout the fused splat reads -- and the backward converts the splat's
                plane gradients back to bf16 rows in one pass per sample
  splat         modules.ProgressiveKernelApply (fused fp32 forward / backward kernels)

Weight normalization of all 57 convolutions is ONE node (`weight_bank.WeightBank`: one
launch prepares every bf16 operand in both layouts, one launch turns all weight gradients
into the gradients of weight_v / weight_g).  The loss stays ordinary PyTorch, the optimizer
is FusedAdam.
Gradients carry bf16 rounding of the activations (a few percent in norm against the
fp32 module after ~40 layers; tests/test_train_pipeline.py states the bars), so this
is not the default training path.
"""
import torch as th

from . import _lib
from . import train_ops as T
from ._compat import crop_like
from .weight_bank import WeightBank

__all__ = ["supported", "forward_train", "EmbedStage", "RegressStage", "UNetStage", "WeightBank",
           "chain_specs", "unet_specs", "embed_stage", "regress_stage", "unet_stage"]

BF = th.bfloat16


# -- the weight bank of a model ---------------------------------------------------------------
def _chain_convs(chain):
    from .conv1x1 import _convs
    return _convs(chain)


def _chain_act(chain):
    return 2 if isinstance(chain.layer_0.layer[1], th.nn.LeakyReLU) else 1


def chain_specs(chain, cin_pad, cout_pad=0):
    """Bank specs of a depth-3 1x1 chain: layer 1's input channels padded to cin_pad, the
    prediction's output channels to cout_pad."""
    c1, c2, c3 = _chain_convs(chain)
    return [(c1, cin_pad, 0), (c2, 0, 0), (c3, 0, cout_pad)]


def unet_specs(autoencoder):
    plan = _unet_plan(autoencoder)
    return plan, [(conv, 0, 0) for left, right in plan for conv, _ in left + (right or [])]


def _model_bank(model):
    """(bank, index) of a Multisteps model; cached on the model, rebuilt when a parameter
    moved.  Layer 1 of the first embedding sees the nf + ngf <= 128 input channels padded to
    128, every later chain 128 sample + 128 pixel channels.  index: {"embed": [(k1, k2, k3)], "unet": [(plan, [k...])], "reg": (k1, k2, k3)}."""
    specs, index = [], {"embed": [], "unet": []}
    for step in range(model.nsteps):
        k = len(specs)
        specs += chain_specs(getattr(model, "embedding_{:02d}".format(step)),
                             128 if step == 0 else 256)
        index["embed"].append((k, k + 1, k + 2))
        plan, us = unet_specs(getattr(model, "propagation_{:02d}".format(step)))
        index["unet"].append((plan, list(range(len(specs), len(specs) + len(us)))))
        specs += us
    k = len(specs)
    k2 = model.kernel_regressor.prediction.out_channels
    specs += chain_specs(model.kernel_regressor, 256, (k2 + 127) // 128 * 128)
    index["reg"] = (k, k + 1, k + 2)
    cached = model.__dict__.get("_sbmc_b200_bank")
    if cached is None or not cached[0].matches(specs):
        cached = (WeightBank(specs), index)
        model.__dict__["_sbmc_b200_bank"] = cached
    return cached[0], index


# -- per-sample 1x1 chains ----------------------------------------------------------------
def _chain_backward(dy, x, ctx_rows, h1, h2, bank, ks, act, n_img, spp, need_dx, need_dctx,
                    cin_valid, cout_valid=0):
    """Backward of y = W3 a(W2 a(W1 [x | ctx] + b1) + b2) + b3 given dy (bf16 rows, channels
    padded to a multiple of 128).  Weight gradients go into the bank's dW buffers.  Returns
    (dx, dctx, dw1, db1, dw2, db2, dw3, db3)."""
    k1, k2, k3 = ks
    ca = x.shape[1]
    cb = 0 if ctx_rows is None else ctx_rows.shape[1]
    dw3, db3 = T.wgrad(dy, h2, dw=bank.dw(k3), cout_valid=cout_valid)
    dh2 = T.linear(dy, bank.dgrad(k3), mask=h2, mask_act=act)
    dw2, db2 = T.wgrad(dh2, h1, dw=bank.dw(k2))
    dh1 = T.linear(dh2, bank.dgrad(k2), mask=h1, mask_act=act)
    dw1 = bank.dw(k1)
    _, db1 = T.wgrad(dh1, x, dw=dw1[:, :cin_valid], cin_valid=cin_valid)
    d1 = bank.dgrad(k1)                                # [cin_pad, 128]: x's rows, then ctx's
    dx = dctx = None
    if need_dx:
        dx = T.linear(dh1, d1[:ca])
    if cb:
        r = T.spp_reduce(dh1, n_img, spp, 1.0)          # sum over the samples of a pixel
        T.wgrad(r, ctx_rows, dw=dw1[:, cin_valid:], want_bias=False)
        if need_dctx:
            dctx = T.linear(r, d1[ca:])
    return dx, dctx, dw1, db1, dw2, db2, dw3, db3


class EmbedStage(th.autograd.Function):
    """(e, reduced) = embedding chain on the sample rows x [S, ca] (+ pixel rows ctx [P, 128])
    and its mean over the samples.  t1..t3: the bank's tokens of the three convolutions `ks`."""

    @staticmethod
    def forward(ctx, x, ctx_rows, t1, b1, t2, b2, t3, b3, bank, ks, act, n_img, spp, hw):
        cb = 0 if ctx_rows is None else ctx_rows.shape[1]
        h1 = T.linear(x, bank.fwd(ks[0]), b1.detach(), act, xb=ctx_rows, hw=hw, spp=spp)
        h2 = T.linear(h1, bank.fwd(ks[1]), b2.detach(), act)
        e = T.linear(h2, bank.fwd(ks[2]), b3.detach(), 0)
        reduced = T.spp_reduce(e, n_img, spp, 1.0 / spp)
        ctx.save_for_backward(x, ctx_rows, h1, h2)
        ctx.cfg = (bank, ks, act, n_img, spp, t1.shape[1] - cb)
        return e, reduced

    @staticmethod
    def backward(ctx, de, dreduced):
        x, ctx_rows, h1, h2 = ctx.saved_tensors
        bank, ks, act, n_img, spp, cin_valid = ctx.cfg
        if dreduced is not None:
            de = T.bcast_add(None if de is None else de.contiguous(), dreduced.contiguous(),
                             n_img, spp, 1.0 / spp)
        else:
            de = de.contiguous()
        need = ctx.needs_input_grad
        dx, dctx, dw1, db1, dw2, db2, dw3, db3 = _chain_backward(
            de, x, ctx_rows, h1, h2, bank, ks, act, n_img, spp, need[0], need[1], cin_valid)
        return (dx, dctx, dw1, db1, dw2, db2, dw3, db3) + (None,) * 6


class RegressStage(th.autograd.Function):
    """Kernel logits of every sample: spp tensors [bs, k2, hw] fp32 (views of one
    [spp, bs, k2, hw] buffer) from the sample rows e [S, 128] and pixel rows ctx [P, 128]."""

    @staticmethod
    def forward(ctx, e, ctx_rows, t1, b1, t2, b2, t3, b3, bank, ks, act, n_img, spp, hw):
        k2 = t3.shape[0]
        w3 = bank.fwd(ks[2])                            # [k2 padded, 128], zero rows past k2
        k2p = w3.shape[0]
        h1 = T.linear(e, bank.fwd(ks[0]), b1.detach(), act, xb=ctx_rows, hw=hw, spp=spp)
        h2 = T.linear(h1, bank.fwd(ks[1]), b2.detach(), act)
        b3p = th.nn.functional.pad(b3.detach(), (0, k2p - k2))
        logits = th.empty(spp, n_img, k2, hw, device=e.device, dtype=th.float32)
        T.linear(h2, w3, b3p, 0, hw=hw, spp=spp, out_mode=2, out=logits,
                 out_img_stride=k2 * hw, out_smp_stride=n_img * k2 * hw, cout_valid=k2)
        ctx.save_for_backward(e, ctx_rows, h1, h2)
        ctx.cfg = (bank, ks, act, n_img, spp, hw, k2, k2p)
        return tuple(logits[s] for s in range(spp))

    @staticmethod
    def backward(ctx, *dlogits):
        e, ctx_rows, h1, h2 = ctx.saved_tensors
        bank, ks, act, n_img, spp, hw, k2, k2p = ctx.cfg
        dy = th.empty(n_img, spp, hw, k2p, device=e.device, dtype=BF)
        for s, g in enumerate(dlogits):
            if g is None:
                dy[:, s].zero_()
            else:
                T.planes_to_rows(g.contiguous().view(n_img, k2, hw), k2p, out=dy[:, s],
                                 out_img_stride=spp * hw * k2p)
        need = ctx.needs_input_grad
        dx, dctx, dw1, db1, dw2, db2, dw3, db3 = _chain_backward(
            dy.view(n_img * spp * hw, k2p), e, ctx_rows, h1, h2, bank, ks, act, n_img, spp,
            need[0], need[1], e.shape[1], cout_valid=k2)
        return (dx, dctx, dw1, db1, dw2, db2, dw3, db3) + (None,) * 6


def embed_stage(bank, ks, toks, chain, x, ctx_rows, n_img, spp, hw):
    c1, c2, c3 = _chain_convs(chain)
    return EmbedStage.apply(x, ctx_rows, toks[ks[0]], c1.bias, toks[ks[1]], c2.bias,
                            toks[ks[2]], c3.bias, bank, ks, _chain_act(chain), n_img, spp, hw)


def regress_stage(bank, ks, toks, chain, e, ctx_rows, n_img, spp, hw):
    c1, c2, c3 = _chain_convs(chain)
    return RegressStage.apply(e, ctx_rows, toks[ks[0]], c1.bias, toks[ks[1]], c2.bias,
                              toks[ks[2]], c3.bias, bank, ks, _chain_act(chain), n_img, spp, hw)


# -- U-net ---------------------------------------------------------------------------------
def _unet_plan(autoencoder):
    """[(left convs, right convs or None)] per level, finest first; a conv entry is
    (module, act_code)."""
    from .unet_fast import _chain_layers, _levels
    plan = []
    for lvl in _levels(autoencoder.net):
        left = _chain_layers(lvl.left)
        right = None if lvl.is_last else _chain_layers(lvl.right)
        plan.append((left, right))
    return plan


def _maxpool(x):
    n, h, w, c = x.shape
    y = th.empty(n, h // 2, w // 2, c, device=x.device, dtype=BF)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_maxpool2x2_nhwc_bf16(x.data_ptr(), y.data_ptr(), n, h, w, c,
                                           th.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "maxpool2x2")
    return y


def _upsample_concat(coarse, skip):
    n, hl, wl, cu = coarse.shape
    _, h, w, cs = skip.shape
    out = th.empty(n, h, w, cu + cs, device=skip.device, dtype=BF)
    lib = _lib.load()
    with th.cuda.device(skip.device):
        rc = lib.sbmc_upsample_concat_nhwc_bf16(
            coarse.data_ptr(), skip.data_ptr(), out.data_ptr(), n, hl, wl, h, w, cu, cs,
            th.cuda.current_stream(skip.device).cuda_stream)
    _lib.check(rc, "upsample_concat")
    return out


OWN_WGRAD3X3 = True      # False: cuDNN's bf16 weight-gradient kernel (A/B comparisons)


def _conv_wgrad(dpre, x, dst):
    """(dW, db) of a 3x3 convolution, dW into dst (fp32 [9, cout, cin]): the split-K tcgen05
    kernel of csrc/wgrad.cu (the bias gradient comes out of the same launch)."""
    _, cout, cin = dst.shape
    if OWN_WGRAD3X3:
        return T.wgrad3x3(dpre, x, out=dst, want_bias=True)
    dw = th.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (cout, cin, 3, 3),
                                  dpre.permute(0, 3, 1, 2), padding=1)
    dst.copy_(dw.permute(2, 3, 0, 1).reshape(9, cout, cin))
    return dst, T.colsum(dpre.view(-1, cout))


class UNetStage(th.autograd.Function):
    """y = Autoencoder(x) on bf16 [n, h, w, c] tensors.  `params` = (bank token, bias) of
    every convolution in `plan` order (level by level, left then right); ks: their bank
    indices."""

    @staticmethod
    def forward(ctx, x, plan, bank, ks, *params):
        tape = []            # one record per level, finest first

        # parameters come level by level, left then right; the recursion visits left(k),
        # the deeper levels, then right(k)
        order, idx = {}, 0
        for k, (left, right) in enumerate(plan):
            order[(k, "left")] = idx
            idx += 2 * len(left)
            if right is not None:
                order[(k, "right")] = idx
                idx += 2 * len(right)

        def conv_chain(layers, a, start):
            recs = []
            for j, (_, act) in enumerate(layers):
                i = start + 2 * j
                y = T.conv3x3(a, bank.fwd(ks[i // 2]), params[i + 1].detach(), act)
                recs.append((i, a, y, act))
                a = y
            return a, recs

        def level(k, a):
            left, right = plan[k]
            a, lrec = conv_chain(left, a, order[(k, "left")])
            rec = {"left": lrec}
            tape.append(rec)
            if right is None:
                return a
            coarse = level(k + 1, _maxpool(a))
            cat = _upsample_concat(coarse, a)
            rec["coarse"] = coarse
            out, rec["right"] = conv_chain(right, cat, order[(k, "right")])
            return out

        y = level(0, x.contiguous())
        ctx.tape = tape
        ctx.bank, ctx.ks = bank, ks
        ctx.nparams = len(params)
        return y

    @staticmethod
    def backward(ctx, dy):
        tape, bank, ks = ctx.tape, ctx.bank, ctx.ks
        grads = [None] * ctx.nparams
        zeros = {}

        def zero_bias(c, dev):
            if c not in zeros:
                zeros[c] = th.zeros(c, device=dev, dtype=th.float32)
            return zeros[c]

        def chain_bwd(recs, dpre):
            """dpre: gradient w.r.t. the pre-activation of the chain's last conv.  Returns the
            gradient w.r.t. the chain's input (no activation derivative applied)."""
            for j in range(len(recs) - 1, -1, -1):
                i, a, y, act = recs[j]
                k = ks[i // 2]
                wd = bank.dgrad(k)                           # [9, cin, cout]
                cin, cout = wd.shape[1], wd.shape[2]
                grads[i], grads[i + 1] = _conv_wgrad(dpre, a, bank.dw(k))
                if j > 0:
                    # the input of this conv is the activated output of the previous one:
                    # its derivative goes into the epilogue
                    dpre = T.conv3x3(dpre, wd, zero_bias(cin, dpre.device), 0,
                                     mask=a, mask_act=recs[j - 1][3])
                else:
                    dpre = T.conv3x3(dpre, wd, zero_bias(cin, dpre.device), 0)
            return dpre

        def level_bwd(k, dpre):
            rec = tape[k]
            if "right" not in rec:
                return chain_bwd(rec["left"], dpre)
            dcat = chain_bwd(rec["right"], dpre)
            coarse = rec["coarse"]
            cu = coarse.shape[3]
            nxt = tape[k + 1]
            coarse_act = (nxt["right"] if "right" in nxt else nxt["left"])[-1][3]
            dcoarse = T.upsample_bwd(dcat[..., :cu], coarse.shape[1:3], coarse, coarse_act)
            dpooled = level_bwd(k + 1, dcoarse)
            left_out, left_act = rec["left"][-1][2], rec["left"][-1][3]
            dleft = T.maxpool2x2_bwd(left_out, dpooled.contiguous(), dcat[..., cu:], left_act)
            return chain_bwd(rec["left"], dleft)

        top = tape[0]
        last = (top["right"] if "right" in top else top["left"])[-1]
        dpre = T.dact(last[2], dy.contiguous(), last[3]) if last[3] else dy.contiguous()
        dx = level_bwd(0, dpre)
        ctx.tape = None
        return (dx, None, None, None) + tuple(grads)


def unet_stage(bank, plan, ks, toks, x):
    """Differentiable bf16 forward of `modules.Autoencoder` on x [n, h, w, c] bf16."""
    params = []
    convs = [conv for left, right in plan for conv, _ in left + (right or [])]
    for k, conv in zip(ks, convs):
        params.append(toks[k])
        params.append(conv.bias)
    return UNetStage.apply(x, plan, bank, ks, *params)


# -- the model --------------------------------------------------------------------------
def supported(model, nf, ngf, h, w):
    """Whether `forward_train` serves this model / input shape."""
    from . import conv1x1, unet_fast
    if model.width != 128 or model.embedding_width != 128 or not model.splat:
        return False
    if nf + ngf > 128 or (h * w) % 256 != 0:
        return False
    chains = [getattr(model, "embedding_{:02d}".format(i)) for i in range(model.nsteps)]
    if not all(conv1x1.supports(c) for c in chains + [model.kernel_regressor]):
        return False
    if not all(unet_fast.supports_training(getattr(model, "propagation_{:02d}".format(i)))
               for i in range(model.nsteps)):
        return False
    # the weight bank normalizes every convolution itself
    return all(hasattr(m, "weight_v") and hasattr(m, "weight_g") and m.weight_v.is_cuda
               and m.weight_v.dtype == th.float32
               for m in model.modules() if isinstance(m, th.nn.Conv2d))


def forward_train(model, radiance, features, gfeatures):
    """The train branch of `Multisteps.forward` (sbmc/models.py:171-209).  Keeps the
    reference's pairing of samples and global features (sample-major tiling against a
    batch-major flattening: sample (b, s) sees global_features[(b spp + s) % bs])."""
    bs, spp, nf, h, w = features.shape
    hw = h * w
    ngf = gfeatures.shape[1]
    feats = features.contiguous().float()
    x = T.planes_to_rows(feats.view(bs * spp, nf, hw), 128)            # [bs spp, hw, 128]
    gidx = th.arange(bs * spp, device=feats.device) % bs
    x[:, :, nf:nf + ngf] = gfeatures.reshape(bs, ngf)[gidx].unsqueeze(1).to(BF)
    x = x.view(bs * spp * hw, 128)
    bank, index = _model_bank(model)
    toks = bank.apply()                      # one launch: every bf16 operand of the step
    prop = None
    for step in range(model.nsteps):
        embed = getattr(model, "embedding_{:02d}".format(step))
        x, reduced = embed_stage(bank, index["embed"][step], toks, embed, x, prop, bs, spp, hw)
        plan, ks = index["unet"][step]
        prop = unet_stage(bank, plan, ks, toks, reduced.view(bs, h, w, 128)).view(bs * hw, 128)
    logits = regress_stage(bank, index["reg"], toks, model.kernel_regressor, x, prop, bs, spp, hw)
    k2 = logits[0].shape[1]
    sum_r = sum_w = max_w = None
    for sp in range(spp):
        kernels = logits[sp].view(bs, k2, h, w)
        sum_r, sum_w, max_w = model.kernel_update(
            crop_like(radiance[:, sp], kernels), kernels, sum_r, sum_w, max_w)
    output = sum_r / (sum_w + model.eps)
    crop = (model.ksize - 1) // 2
    return {"radiance": output[..., crop:-crop, crop:-crop]}
