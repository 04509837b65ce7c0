"""Kernels of the mixed-precision TRAINING pipeline (`sbmc_b200/train_pipeline.py`):
thin bindings of the C ABI (include/sbmc_b200.h; csrc/linear.cu, csrc/wgrad.cu,
csrc/conv3x3.cu, csrc/train_ops.cu).  All activations are bf16 channels-innermost
rows; gradients of weights / biases are fp32.  No autograd here, no fallbacks.
"""
import torch as th

from . import _lib

__all__ = ["linear", "wgrad", "wgrad3x3", "conv3x3", "spp_reduce", "bcast_add", "maxpool2x2_bwd",
           "upsample_bwd", "dact", "colsum", "planes_to_rows"]

_BF16 = th.bfloat16


def _stream(t):
    return th.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


_SMS = {}


def _num_sms(dev):
    key = (dev.type, dev.index)
    if key not in _SMS:
        _SMS[key] = th.cuda.get_device_properties(dev).multi_processor_count
    return _SMS[key]


def linear(x, w, bias=None, act=0, xb=None, hw=0, spp=0, mask=None, mask_act=0,
           out_mode=0, out=None, out_img_stride=0, out_smp_stride=0, cout_valid=0):
    """y = act([x | xb] . w^T + bias) [* act'(mask)]  (sbmc_linear2_nhwc_bf16).

    x bf16 [rows, cin_a]; xb bf16 [rows / spp, cin_b] (one row per pixel, shared by the
    samples) or None; w bf16 [cout, cin_a + cin_b]; bias fp32 [cout] or None; mask bf16
    [rows, cout] or None.  out_mode 0: bf16 [rows, cout]; 1: fp32 [rows, cout]; 2: fp32
    channel planes written into `out` (caller-allocated) at
    out[b * out_img_stride + s * out_smp_stride + c * hw + p] for c < cout_valid."""
    rows, cin_a = x.shape
    cin_b = 0 if xb is None else xb.shape[1]
    cout = w.shape[0]
    if x.dtype != _BF16 or w.dtype != _BF16 or not x.is_contiguous() or not w.is_contiguous() \
            or w.shape[1] != cin_a + cin_b:
        raise RuntimeError("linear: expected contiguous bf16 [rows, cin] and [cout, cin]")
    if xb is not None and (xb.dtype != _BF16 or not xb.is_contiguous()
                           or xb.shape[0] * spp != rows):
        raise RuntimeError("linear: second source must be contiguous bf16 [rows / spp, cin_b]")
    if bias is not None and (bias.dtype != th.float32 or bias.numel() != cout
                             or not bias.is_contiguous()):
        raise RuntimeError("linear: bias must be contiguous float32 [cout]")
    if mask is not None and (mask.dtype != _BF16 or tuple(mask.shape) != (rows, cout)
                             or not mask.is_contiguous()):
        raise RuntimeError("linear: mask must be contiguous bf16 [rows, cout]")
    if out_mode == 2:
        if out is None or out.dtype != th.float32:
            raise RuntimeError("linear: plane mode writes into a caller-allocated fp32 tensor")
        y = out
    else:
        y = th.empty(rows, cout, device=x.device, dtype=th.float32 if out_mode == 1 else _BF16)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_linear2_nhwc_bf16(
            x.data_ptr(), cin_a, _ptr(xb), cin_b, hw, spp, w.data_ptr(), _ptr(bias), _ptr(mask),
            mask_act, y.data_ptr(), out_mode, out_img_stride, out_smp_stride, cout_valid, rows,
            cout, act, _stream(x))
    _lib.check(rc, "linear")
    return y


def wgrad(dy, x, dw=None, cout_valid=0, cin_valid=0, want_bias=True):
    """(dw, db): dw[co, ci] = sum_r dy[r, co] x[r, ci] (fp32), db[co] = sum_r dy[r, co].

    dy bf16 [rows, cout] contiguous; x bf16 [rows, cin] with any row pitch (a column slice of a
    wider tensor is fine); cout, cin multiples of 128.  `dw`: optional fp32 destination
    (a [cout_valid, cin_valid] view, unit column stride); otherwise allocated."""
    rows, cout = dy.shape
    cin = x.shape[1]
    if dy.dtype != _BF16 or x.dtype != _BF16 or not dy.is_contiguous() or x.stride(1) != 1 \
            or x.shape[0] != rows:
        raise RuntimeError("wgrad: expected bf16 [rows, cout] (contiguous) and [rows, cin]")
    cv = cout_valid or cout
    civ = cin_valid or cin
    if dw is None:
        dw = th.empty(cv, civ, device=dy.device, dtype=th.float32)
    if dw.dtype != th.float32 or tuple(dw.shape) != (cv, civ) or dw.stride(1) != 1:
        raise RuntimeError("wgrad: dw must be float32 [cout_valid, cin_valid]")
    db = th.empty(cv, device=dy.device, dtype=th.float32) if want_bias else None
    blocks = (cout // 128) * (cin // 128)
    nsplit = max(1, min((rows + 127) // 128, _num_sms(dy.device) // max(blocks, 1)))
    ws = th.empty(nsplit * cout * (cin + 1), device=dy.device, dtype=th.float32)
    lib = _lib.load()
    with th.cuda.device(dy.device):
        rc = lib.sbmc_wgrad_nhwc_bf16(dy.data_ptr(), x.data_ptr(), x.stride(0), rows, cout, cin,
                                      nsplit, ws.data_ptr(), dw.data_ptr(), dw.stride(0), cv, civ,
                                      _ptr(db), _stream(dy))
    _lib.check(rc, "wgrad")
    return dw, db


def wgrad3x3(dp, x, out=None, want_bias=False):
    """fp32 [9, cout, cin] weight gradient of a 3x3 / pad 1 convolution, tap = 3 dy + dx:
    sum over pixels of dp[n, y, x, co] * x[n, y + dy - 1, x + dx - 1, ci] (csrc/wgrad.cu);
    with want_bias: (dw9, db), db[co] = sum over pixels of dp (from the same kernel)."""
    n, h, w, cout = dp.shape
    cin = x.shape[3]
    if dp.dtype != _BF16 or x.dtype != _BF16 or not dp.is_contiguous() or not x.is_contiguous() \
            or tuple(x.shape[:3]) != (n, h, w):
        raise RuntimeError("wgrad3x3: expected contiguous bf16 [n,h,w,cout] and [n,h,w,cin]")
    blocks = (cout // 128) * (cin // 128) * 3
    px = 32 if w <= 32 else 64
    chunks = n * ((h + 128 // px - 1) // (128 // px)) * ((w + px - 1) // px)
    nsplit = max(1, min(chunks, _num_sms(dp.device) // max(blocks, 1)))
    ws = th.empty(nsplit * cout * (9 * cin + 1), device=dp.device, dtype=th.float32)
    db = th.empty(cout, device=dp.device, dtype=th.float32) if want_bias else None
    dw9 = out if out is not None else th.empty(9, cout, cin, device=dp.device, dtype=th.float32)
    if dw9.dtype != th.float32 or tuple(dw9.shape) != (9, cout, cin) or not dw9.is_contiguous():
        raise RuntimeError("wgrad3x3: out must be contiguous float32 [9, cout, cin]")
    lib = _lib.load()
    with th.cuda.device(dp.device):
        rc = lib.sbmc_wgrad3x3_nhwc_bf16(dp.data_ptr(), x.data_ptr(), n, h, w, cout, cin, nsplit,
                                         ws.data_ptr(), dw9.data_ptr(), _ptr(db), _stream(dp))
    _lib.check(rc, "wgrad3x3")
    return (dw9, db) if want_bias else dw9


def conv3x3(x, w9, bias, act=0, mask=None, mask_act=0):
    """act(conv3x3(x) + bias) [* act'(mask)] on bf16 [n, h, w, c] (csrc/conv3x3.cu)."""
    n, h, w, cin = x.shape
    cout = w9.shape[1]
    if x.dtype != _BF16 or not x.is_contiguous() or w9.dtype != _BF16 \
            or tuple(w9.shape) != (9, cout, cin) or not w9.is_contiguous():
        raise RuntimeError("conv3x3: expected contiguous bf16 [n,h,w,cin] and [9,cout,cin]")
    if bias.dtype != th.float32 or bias.numel() != cout:
        raise RuntimeError("conv3x3: bias must be float32 [cout]")
    if mask is not None and (mask.dtype != _BF16 or tuple(mask.shape) != (n, h, w, cout)
                             or not mask.is_contiguous()):
        raise RuntimeError("conv3x3: mask must be contiguous bf16 [n,h,w,cout]")
    y = th.empty(n, h, w, cout, device=x.device, dtype=_BF16)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_conv3x3_masked_nhwc_bf16(x.data_ptr(), w9.data_ptr(), bias.data_ptr(),
                                               _ptr(mask), mask_act, y.data_ptr(), n, h, w, cin,
                                               cout, act, _stream(x))
    _lib.check(rc, "conv3x3")
    return y


def spp_reduce(x, n_img, spp, scale=1.0, out_f32=False):
    """x bf16 [n_img * spp * hw, c] (image, sample, pixel) -> [n_img * hw, c] = scale * sum_s."""
    rows, c = x.shape
    hw = rows // (n_img * spp)
    if x.dtype != _BF16 or not x.is_contiguous() or hw * n_img * spp != rows:
        raise RuntimeError("spp_reduce: expected contiguous bf16 [n_img * spp * hw, c]")
    out = th.empty(n_img * hw, c, device=x.device, dtype=th.float32 if out_f32 else _BF16)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_spp_reduce_nhwc_bf16(x.data_ptr(), out.data_ptr(), 1 if out_f32 else 0,
                                           n_img, spp, hw, c, float(scale), _stream(x))
    _lib.check(rc, "spp_reduce")
    return out


def bcast_add(a, r, n_img, spp, scale=1.0):
    """a bf16 [n_img * spp * hw, c] or None, r bf16 [n_img * hw, c] -> a + scale * r (broadcast
    over the samples)."""
    rows_r, c = r.shape
    hw = rows_r // n_img
    if r.dtype != _BF16 or not r.is_contiguous() or (a is not None and (
            a.dtype != _BF16 or not a.is_contiguous() or tuple(a.shape) != (rows_r * spp, c))):
        raise RuntimeError("bcast_add: expected contiguous bf16 [rows, c] tensors")
    out = th.empty(rows_r * spp, c, device=r.device, dtype=_BF16)
    lib = _lib.load()
    with th.cuda.device(r.device):
        rc = lib.sbmc_bcast_add_nhwc_bf16(_ptr(a), r.data_ptr(), out.data_ptr(), n_img, spp, hw, c,
                                          float(scale), _stream(r))
    _lib.check(rc, "bcast_add")
    return out


def maxpool2x2_bwd(x, dpool, dskip, act):
    """Gradient w.r.t. the PRE-activation of x [n, h, w, c] (x = act(pre)) given the gradient
    of maxpool2x2(x) and (optionally) of a second use of x (`dskip`: [n, h, w, c] view with
    unit channel stride, e.g. the skip half of a concatenation's gradient)."""
    n, h, w, c = x.shape
    if x.dtype != _BF16 or not x.is_contiguous() or dpool.dtype != _BF16 \
            or not dpool.is_contiguous() or tuple(dpool.shape) != (n, h // 2, w // 2, c):
        raise RuntimeError("maxpool2x2_bwd: shape / layout mismatch")
    pitch = 0
    if dskip is not None:
        if dskip.dtype != _BF16 or tuple(dskip.shape) != (n, h, w, c) or dskip.stride(3) != 1 \
                or dskip.stride(1) != w * dskip.stride(2) or dskip.stride(0) != h * dskip.stride(1):
            raise RuntimeError("maxpool2x2_bwd: dskip must be a channel slice of an [n,h,w,C] tensor")
        pitch = dskip.stride(2)
    out = th.empty_like(x)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_maxpool2x2_bwd_nhwc_bf16(x.data_ptr(), dpool.data_ptr(), _ptr(dskip), pitch,
                                               out.data_ptr(), n, h, w, c, act, _stream(x))
    _lib.check(rc, "maxpool2x2_bwd")
    return out


def upsample_bwd(dup, coarse_shape, coarse=None, act=0):
    """Transpose of F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False):
    dup [n, h, w, c] (channel slice of a wider tensor allowed) -> [n, hl, wl, c], multiplied by
    act'(coarse) when act != 0."""
    n, h, w, c = dup.shape
    hl, wl = coarse_shape
    if dup.dtype != _BF16 or dup.stride(3) != 1 or dup.stride(1) != w * dup.stride(2) \
            or dup.stride(0) != h * dup.stride(1):
        raise RuntimeError("upsample_bwd: dup must be a channel slice of an [n,h,w,C] tensor")
    if act and (coarse is None or coarse.dtype != _BF16 or not coarse.is_contiguous()
                or tuple(coarse.shape) != (n, hl, wl, c)):
        raise RuntimeError("upsample_bwd: coarse must be contiguous bf16 [n,hl,wl,c]")
    out = th.empty(n, hl, wl, c, device=dup.device, dtype=_BF16)
    lib = _lib.load()
    with th.cuda.device(dup.device):
        rc = lib.sbmc_upsample_bwd_nhwc_bf16(dup.data_ptr(), dup.stride(2), _ptr(coarse),
                                             out.data_ptr(), n, hl, wl, h, w, c, act, _stream(dup))
    _lib.check(rc, "upsample_bwd")
    return out


def dact(y, g, act):
    """g * act'(y) (bf16, same shape, contiguous)."""
    if y.dtype != _BF16 or g.dtype != _BF16 or y.shape != g.shape or not y.is_contiguous() \
            or not g.is_contiguous():
        raise RuntimeError("dact: expected two contiguous bf16 tensors of one shape")
    out = th.empty_like(g)
    lib = _lib.load()
    with th.cuda.device(y.device):
        rc = lib.sbmc_dact_bf16(y.data_ptr(), g.data_ptr(), out.data_ptr(), y.numel(), act,
                                _stream(y))
    _lib.check(rc, "dact")
    return out


def colsum(x):
    """fp32 column sums of a bf16 matrix [rows, c] (any row pitch)."""
    rows, c = x.shape
    if x.dtype != _BF16 or x.stride(1) != 1:
        raise RuntimeError("colsum: expected bf16 [rows, c] with unit column stride")
    nblk = max(1, min(_num_sms(x.device), (rows + 63) // 64))
    ws = th.empty(nblk * c, device=x.device, dtype=th.float32)
    out = th.empty(c, device=x.device, dtype=th.float32)
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_colsum_bf16(x.data_ptr(), x.stride(0), rows, c, ws.data_ptr(), nblk,
                                  out.data_ptr(), _stream(x))
    _lib.check(rc, "colsum")
    return out


def planes_to_rows(x, cpad, out=None, out_img_stride=None):
    """fp32 channel planes x [n, c, hw] -> bf16 rows [n, hw, cpad] (zero-padded channels);
    `out` / `out_img_stride` (elements) place the images inside a larger row buffer."""
    n, c, hw = x.shape
    if x.dtype != th.float32 or x.stride(2) != 1 or x.stride(1) != hw:
        raise RuntimeError("planes_to_rows: expected float32 [n, c, hw] planes")
    if out is None:
        out = th.empty(n, hw, cpad, device=x.device, dtype=_BF16)
        out_img_stride = hw * cpad
    lib = _lib.load()
    with th.cuda.device(x.device):
        rc = lib.sbmc_nchw_to_nhwc_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out_img_stride,
                                        n, c, hw, cpad, _stream(x))
    _lib.check(rc, "planes_to_rows")
    return out
