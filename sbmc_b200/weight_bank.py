"""Weight normalization of all convolutions of a model in ONE launch per direction
(csrc/weight_bank.cu), for the training pipeline (`sbmc_b200/train_pipeline.py`).

Every convolution of the reference model is `nn.utils.weight_norm(nn.Conv2d(...))`
(sbmc/modules.py:84-87,176-179).  `WeightBank` owns persistent device buffers with, per
convolution k,

  F[k]   bf16 forward operand  [T, cout_pad, cin_pad]   (T = 9 taps or 1)
  D[k]   bf16 data-gradient operand: [T, cin, cout] with flipped taps (3x3) or the
         transpose [cin_pad, cout_pad] (1x1)
  dW[k]  fp32 weight gradient  [T, cout, cin] -- the weight-gradient kernels write here
  dv, dg fp32 gradients of weight_v / weight_g

`bank.apply()` is one autograd node: forward = `wn_prepare_kernel` over all
convolutions (returns one zero-stride fp32 token per convolution: the edge through which
the stages hand back dW), backward = `wn_backward_kernel`.  All addresses are static, so
the node is CUDA-graph friendly; the buffers are overwritten by the next `apply()` (run
backward before the next forward).
"""
import torch as th

from . import _lib

__all__ = ["WeightBank"]


def _is_weight_normed(conv):
    return hasattr(conv, "weight_g") and hasattr(conv, "weight_v")


class _BankFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, bank, *vg):
        ctx.bank = bank
        bank._run(False)
        return tuple(th.empty(1, device=bank.device).expand(e["dw_shape"]) for e in bank.entries)

    @staticmethod
    def backward(ctx, *dws):
        bank = ctx.bank
        for k, g in enumerate(dws):
            dst = bank.dw(k)
            if g is None:
                dst.zero_()
            elif g.data_ptr() != dst.data_ptr() or not g.is_contiguous():
                dst.copy_(g)            # a stage that did not write in place
        bank._run(True)
        out = [None]
        for e in bank.entries:
            out.append(e["dv"])
            out.append(e["dg"])
        return tuple(out)


class WeightBank(object):
    """specs: [(conv module, cin_pad, cout_pad)]; every conv must be weight-normalized,
    CUDA, fp32, with a 1x1 or 3x3 kernel."""

    def __init__(self, specs):
        self.entries = []
        convs = [s[0] for s in specs]
        if not convs:
            raise ValueError("WeightBank: no convolutions")
        self.device = convs[0].weight_v.device
        f_off = d_off = w_off = r_off = 0
        for conv, cinp, coutp in specs:
            if not _is_weight_normed(conv):
                raise ValueError("WeightBank: convolution without weight normalization")
            v, g = conv.weight_v, conv.weight_g
            cout, cin, kh, kw = v.shape
            t = kh * kw
            if t not in (1, 9) or kh != kw or v.dtype != th.float32 or not v.is_cuda \
                    or not v.is_contiguous() or not g.is_contiguous():
                raise ValueError("WeightBank: unsupported convolution %s" % (tuple(v.shape),))
            cinp = max(cinp or cin, cin)
            coutp = max(coutp or cout, cout)
            if t == 9 and (cinp != cin or coutp != cout):
                raise ValueError("WeightBank: 3x3 operands are not padded")
            e = dict(conv=conv, cout=cout, cin=cin, T=t, coutp=coutp, cinp=cinp,
                     f_off=f_off, d_off=d_off, w_off=w_off, r_off=r_off,
                     dw_shape=(t, cout, cin) if t == 9 else (cout, cin))
            f_off += t * coutp * cinp
            d_off += t * coutp * cinp
            w_off += t * cout * cin
            r_off += cout
            self.entries.append(e)
        dev = self.device
        self.F = th.zeros(f_off, device=dev, dtype=th.bfloat16)
        self.D = th.zeros(d_off, device=dev, dtype=th.bfloat16)
        self.dW = th.zeros(w_off, device=dev, dtype=th.float32)
        self.dV = th.zeros(w_off, device=dev, dtype=th.float32)
        self.dG = th.zeros(r_off, device=dev, dtype=th.float32)
        self.rn = th.zeros(r_off, device=dev, dtype=th.float32)
        rows, blocks = [], []
        for k, e in enumerate(self.entries):
            n = e["T"] * e["cout"] * e["cin"]
            e["dv"] = self.dV[e["w_off"]:e["w_off"] + n].view_as(e["conv"].weight_v)
            e["dg"] = self.dG[e["r_off"]:e["r_off"] + e["cout"]].view_as(e["conv"].weight_g)
            rows.append([e["conv"].weight_v.data_ptr(), e["conv"].weight_g.data_ptr(),
                         self.F.data_ptr() + 2 * e["f_off"], self.D.data_ptr() + 2 * e["d_off"],
                         self.rn.data_ptr() + 4 * e["r_off"], self.dW.data_ptr() + 4 * e["w_off"],
                         self.dV.data_ptr() + 4 * e["w_off"], self.dG.data_ptr() + 4 * e["r_off"],
                         e["cout"], e["cin"], e["T"], e["coutp"], e["cinp"], 0, 0, 0])
            blocks.extend([k, co0] for co0 in range(0, e["cout"], 8))
        self._ptrs = tuple((e["conv"].weight_v.data_ptr(), e["conv"].weight_g.data_ptr())
                           for e in self.entries)
        self._entries_dev = th.tensor(rows, dtype=th.int64).to(dev)
        self._blocks_dev = th.tensor(blocks, dtype=th.int64).to(dev)
        self._nblocks = len(blocks)

    # -- validity: the tables hold raw parameter addresses -------------------------------
    def matches(self, specs):
        return len(specs) == len(self.entries) and all(
            s[0] is e["conv"] and (s[0].weight_v.data_ptr(), s[0].weight_g.data_ptr()) == p
            for s, e, p in zip(specs, self.entries, self._ptrs))

    def _run(self, backward):
        lib = _lib.load()
        with th.cuda.device(self.device):
            rc = lib.sbmc_weight_bank_run(self._entries_dev.data_ptr(), self._blocks_dev.data_ptr(),
                                          self._nblocks, 1 if backward else 0,
                                          th.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "weight_bank")

    # -- the autograd node ------------------------------------------------------------------
    def apply(self):
        """Prepare every operand from the current parameters; returns one token per
        convolution (fp32, the shape of its dW; content-free)."""
        vg = []
        for e in self.entries:
            vg.append(e["conv"].weight_v)
            vg.append(e["conv"].weight_g)
        return _BankFn.apply(self, *vg)

    # -- views ----------------------------------------------------------------------------------
    def fwd(self, k):
        e = self.entries[k]
        n = e["T"] * e["coutp"] * e["cinp"]
        shape = (9, e["coutp"], e["cinp"]) if e["T"] == 9 else (e["coutp"], e["cinp"])
        return self.F[e["f_off"]:e["f_off"] + n].view(shape)

    def dgrad(self, k):
        e = self.entries[k]
        n = e["T"] * e["coutp"] * e["cinp"]
        shape = (9, e["cin"], e["cout"]) if e["T"] == 9 else (e["cinp"], e["coutp"])
        return self.D[e["d_off"]:e["d_off"] + n].view(shape)

    def dw(self, k):
        e = self.entries[k]
        n = e["T"] * e["cout"] * e["cin"]
        return self.dW[e["w_off"]:e["w_off"] + n].view(e["dw_shape"])
