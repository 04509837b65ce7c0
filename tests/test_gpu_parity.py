"""GPU suite: parity of the sm_100a kernels, called through the C ABI
(ctypes -> libsbmc_b200.so), against the CPU oracle and the golden vectors.

Bars (BASELINE.json north_star): KernelWeighting fwd/bwd fp32 within 1e-5
relative (tests/util.py states the exact form); Scatter2Gather bit-exact.
"""
import glob
import os

# small host bands so that the multi-band streaming pipeline is exercised at test
# sizes (read once by the library, before its first host-buffer call)
os.environ.setdefault("SBMC_HOST_BAND_MB", "16")

import numpy as np
import pytest
import torch as th

import oracle
from sbmc_b200 import _lib, halide_ops
import sbmc_b200.functions as funcs
from tests import kats
from tests.test_oracle import check_against_golden, load_golden
from tests.util import assert_close_sum, kw_magnitudes, make_inputs

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
DEV = "cuda"


def run_cuda(data, weights, d_output, d_sum_w):
    """All three ops through the drop-in module (raw C-ABI calls underneath)."""
    data, weights, d_output, d_sum_w = [t.to(DEV).contiguous() for t in
                                        (data, weights, d_output, d_sum_w)]
    n, c, h, w = data.shape
    # outputs start as garbage: the ops must overwrite every element
    out = th.full_like(data, float("nan"))
    sum_w = th.full((n, h, w), float("nan"), device=DEV)
    halide_ops.kernel_weighting_cuda_float32(data, weights, out, sum_w)
    d_data = th.full_like(data, float("nan"))
    d_weights = th.full_like(weights, float("nan"))
    halide_ops.kernel_weighting_grad_cuda_float32(
        data, weights, sum_w, d_output, d_sum_w, d_data, d_weights)
    gather = th.full_like(weights, float("nan"))
    halide_ops.scatter2gather_cuda_float32(weights, gather)
    th.cuda.synchronize()
    return out, sum_w, d_data, d_weights, gather


@pytest.fixture(params=[0, 1], ids=["tuned", "generic"])
def path(request):
    prev = _lib.force_generic(request.param)
    yield request.param
    _lib.force_generic(prev)


def test_library_is_the_cuda_one():
    before = _lib.launch_count()
    w = th.randn(1, 3, 3, 8, 8, device=DEV)
    funcs.Scatter2Gather.apply(w)
    assert _lib.launch_count() > before


@pytest.mark.parametrize("gpath", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden(gpath, path):
    g, t = load_golden(gpath)
    out, sum_w, d_data, d_weights, gather = run_cuda(*t)
    check_against_golden(g, t, out, sum_w, d_data, d_weights, gather,
                         "cuda[%s]" % ("generic" if path else "tuned"))


SHAPES = [
    # n, c, h, w, kh, kw
    (2, 3, 64, 64, 5, 5),        # BASELINE config 1
    (1, 3, 40, 260, 21, 21),     # model kernel, W spans 3 x-tiles, ragged
    (2, 3, 33, 132, 21, 21),
    (1, 3, 7, 1280, 21, 21),     # full 720p width, few rows
    (3, 5, 16, 16, 5, 5),        # reference test shape
    (2, 5, 19, 36, 3, 3),
    (1, 3, 30, 128, 7, 7),
    (1, 3, 21, 50, 21, 21),      # W % 4 != 0 -> generic
    (2, 2, 13, 20, 4, 6),        # even, non-square
    (1, 1, 5, 4, 9, 9),          # kernel larger than the image
    (1, 7, 9, 12, 3, 5),         # odd channel count
    (1, 3, 1, 4, 1, 1),
    (1, 3, 24, 132, 9, 9),       # the other odd kernel sizes of Multisteps(ksize=...):
    (2, 3, 17, 260, 13, 13),     # tuned instantiations added in round 2
    (1, 3, 30, 128, 19, 19),
    (1, 3, 12, 256, 11, 15),     # non-square: kw selects the instantiation, kh is a runtime loop
]


@pytest.mark.parametrize("shape", SHAPES, ids=["x".join(map(str, s)) for s in SHAPES])
def test_random_vs_oracle(shape, path):
    n, c, h, w, kh, kw = shape
    data, weights, d_output, d_sum_w = make_inputs(n, c, h, w, kh, kw, seed=sum(shape))
    mo, ms, mdd, mdw = kw_magnitudes(data, weights, d_output, d_sum_w)
    ro, rs = oracle.kernel_weighting(data, weights)
    rdd, rdw = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
    rg = oracle.scatter2gather(weights)
    out, sum_w, d_data, d_weights, gather = run_cuda(data, weights, d_output, d_sum_w)
    assert_close_sum(out, ro, mo, "output")
    assert_close_sum(sum_w, rs, ms, "sum_w")
    assert_close_sum(d_data, rdd, mdd, "d_data")
    assert_close_sum(d_weights, rdw, mdw, "d_weights")
    assert np.array_equal(gather.cpu().numpy().view(np.uint32), rg.numpy().view(np.uint32))


def test_tuned_path_is_taken_for_model_shapes():
    data, weights, d_output, d_sum_w = make_inputs(1, 3, 16, 128, 21, 21)
    run_cuda(data, weights, d_output, d_sum_w)
    assert _lib.last_path() == 1
    for k in (9, 11, 13, 15, 17, 19):            # every odd ksize up to the model's 21
        data, weights, d_output, d_sum_w = make_inputs(1, 3, 16, 128, k, k)
        halide = run_cuda(data, weights, d_output, d_sum_w)
        assert _lib.last_path() in (1, 2) and halide is not None
        from sbmc_b200 import halide_ops
        dev = "cuda"
        out = th.empty(1, 3, 16, 128, device=dev); sw = th.empty(1, 16, 128, device=dev)
        halide_ops.kernel_weighting_cuda_float32(data.to(dev), weights.to(dev), out, sw)
        assert _lib.last_path() == 1, k
    data, weights, d_output, d_sum_w = make_inputs(1, 3, 16, 50, 21, 21)
    run_cuda(data, weights, d_output, d_sum_w)
    assert _lib.last_path() == 2


def test_empty_inputs():
    for shape in [(0, 3, 4, 4, 3, 3), (2, 3, 0, 4, 3, 3), (2, 3, 4, 0, 3, 3)]:
        n, c, h, w, kh, kw = shape
        d = th.zeros(n, c, h, w, device=DEV)
        k = th.zeros(n, kh, kw, h, w, device=DEV)
        o, s = funcs.KernelWeighting.apply(d, k)
        assert o.shape == d.shape and s.shape == (n, h, w)
        assert funcs.Scatter2Gather.apply(k).shape == k.shape


def test_special_values_move_bit_exactly():
    """Scatter2Gather must move NaN payloads, infinities, -0.0 and denormals."""
    s = th.randn(1, 5, 5, 16, 32)
    bits = s.numpy().view(np.uint32)
    bits[0, 1, 2, 5, 7] = 0x7FC12345      # NaN with payload
    bits[0, 3, 3, 8, 9] = 0x80000000      # -0.0
    bits[0, 0, 4, 9, 20] = 0x00000001     # denormal
    bits[0, 2, 2, 3, 3] = 0xFF800000      # -inf
    for flag in (0, 1):
        prev = _lib.force_generic(flag)
        try:
            g = funcs.Scatter2Gather.apply(s.to(DEV)).cpu()
        finally:
            _lib.force_generic(prev)
        assert np.array_equal(g.numpy().view(np.uint32),
                              oracle.scatter2gather(s).numpy().view(np.uint32))


# -- the reference's known-answer tests on the CUDA Functions -------------------
def test_kat_forward_impulse():
    kats.kw_forward_impulse(funcs.KernelWeighting, DEV)


def test_kat_backward_impulse():
    kats.kw_backward_impulse(funcs.KernelWeighting, DEV)


def test_kat_gradcheck():
    kats.kw_gradcheck(funcs.KernelWeighting, DEV)


def test_kat_scatter2gather_index_map():
    kats.s2g_index_map(funcs.Scatter2Gather, DEV, stride=5)


def test_kat_scatter2gather_gradcheck():
    kats.s2g_gradcheck(funcs.Scatter2Gather, DEV)


# -- row bands (the multi-GPU / host-streaming decomposition) -------------------
@pytest.mark.parametrize("shape,bands", [
    ((1, 3, 48, 256, 21, 21), [0, 7, 20, 48]),
    ((2, 3, 30, 64, 5, 5), [0, 15, 30]),
    ((1, 2, 20, 18, 4, 6), [0, 3, 11, 20]),
])
def test_band_entry_points(shape, bands, path):
    n, c, h, w, kh, kw = shape
    data, weights, d_output, d_sum_w = make_inputs(n, c, h, w, kh, kw, seed=3)
    mo, ms, mdd, mdw = kw_magnitudes(data, weights, d_output, d_sum_w)
    ro, rs = oracle.kernel_weighting(data, weights)
    rdd, rdw = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
    lib = _lib.load()
    st = th.cuda.current_stream().cuda_stream
    D = data.to(DEV)
    out = th.empty(n, c, h, w, device=DEV)
    sum_w = th.empty(n, h, w, device=DEV)
    d_data = th.zeros(n, c, h, w, device=DEV)
    d_weights = th.empty(n, kh, kw, h, w, device=DEV)
    rt, rb = kh - 1 - (kh - 1) // 2, (kh - 1) // 2   # rows a band scatters into
    for y0, y1 in zip(bands[:-1], bands[1:]):
        hb = y1 - y0
        wb = weights[:, :, :, y0:y1].contiguous().to(DEV)
        dob = d_output[:, :, y0:y1].contiguous().to(DEV)
        dsb = d_sum_w[:, y0:y1].contiguous().to(DEV)
        ob = th.empty(n, c, hb, w, device=DEV)
        sb = th.empty(n, hb, w, device=DEV)
        # forward / d_weights see the whole image as the band's extension
        _lib.check(lib.sbmc_kernel_weighting_fwd_band_f32(
            D.data_ptr(), wb.data_ptr(), ob.data_ptr(), sb.data_ptr(), n, c, hb, w,
            kh, kw, y0, h - y1, st), "fwd_band")
        out[:, :, y0:y1] = ob
        sum_w[:, y0:y1] = sb
        # backward with small halos (the same on data_ext and d_data_ext): they
        # must cover both the rows d_weights reads and the rows d_data reaches
        top, bot = min(max(rt, rb), y0), min(max(rt, rb), h - y1)
        dext = D[:, :, y0 - top:y1 + bot].contiguous()
        dde = th.full((n, c, top + hb + bot, w), float("nan"), device=DEV)
        dwb = th.empty(n, kh, kw, hb, w, device=DEV)
        _lib.check(lib.sbmc_kernel_weighting_bwd_band_f32(
            dext.data_ptr(), wb.data_ptr(), dob.data_ptr(), dsb.data_ptr(),
            dde.data_ptr(), dwb.data_ptr(), n, c, hb, w, kh, kw, top, bot, st),
            "bwd_band")
        d_data[:, :, y0 - top:y1 + bot] += dde
        d_weights[:, :, :, y0:y1] = dwb
    th.cuda.synchronize()
    assert_close_sum(out, ro, mo, "band output")
    assert_close_sum(sum_w, rs, ms, "band sum_w")
    assert_close_sum(d_data, rdd, mdd, "band d_data")
    assert_close_sum(d_weights, rdw, mdw, "band d_weights")


# -- host-buffer entry points (the *_cpu_float32 drop-ins) ----------------------
@pytest.mark.parametrize("shape", [(2, 3, 64, 64, 5, 5), (1, 3, 300, 256, 21, 21),
                                   (2, 2, 13, 20, 4, 6)],
                         ids=["cfg1", "k21_multiband", "even"])
def test_host_entry_points(shape):
    n, c, h, w, kh, kw = shape
    data, weights, d_output, d_sum_w = make_inputs(n, c, h, w, kh, kw, seed=9)
    mo, ms, mdd, mdw = kw_magnitudes(data, weights, d_output, d_sum_w)
    ro, rs = oracle.kernel_weighting(data, weights)
    rdd, rdw = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
    out = th.full_like(data, float("nan"))
    sum_w = th.full((n, h, w), float("nan"))
    halide_ops.kernel_weighting_cpu_float32(data, weights, out, sum_w)
    d_data = th.full_like(data, float("nan"))
    d_weights = th.full_like(weights, float("nan"))
    halide_ops.kernel_weighting_grad_cpu_float32(
        data, weights, sum_w, d_output, d_sum_w, d_data, d_weights)
    gather = th.full_like(weights, float("nan"))
    halide_ops.scatter2gather_cpu_float32(weights, gather)
    assert_close_sum(out, ro, mo, "host output")
    assert_close_sum(sum_w, rs, ms, "host sum_w")
    assert_close_sum(d_data, rdd, mdd, "host d_data")
    assert_close_sum(d_weights, rdw, mdw, "host d_weights")
    assert np.array_equal(gather.numpy().view(np.uint32),
                          oracle.scatter2gather(weights).numpy().view(np.uint32))
    # the autograd Functions dispatch host tensors to the same entry points
    o2, s2 = funcs.KernelWeighting.apply(data, weights)
    assert th.equal(o2, out) and th.equal(s2, sum_w)


# -- BASELINE.json config 2 size (per call: N=4, 720p, K=21) --------------------
def test_full_size_720p_vs_oracle():
    """One 720p image (N=1, K=21) against the oracle directly."""
    n, c, h, w, k = 1, 3, 720, 1280, 21
    data, weights, d_output, d_sum_w = make_inputs(n, c, h, w, k, k, seed=42)
    out, sum_w, d_data, d_weights, gather = run_cuda(data, weights, d_output, d_sum_w)
    assert _lib.last_path() == 1
    mo, ms, mdd, mdw = kw_magnitudes(data, weights, d_output, d_sum_w)
    ro, rs = oracle.kernel_weighting(data, weights)
    assert_close_sum(out, ro, mo, "720p output")
    assert_close_sum(sum_w, rs, ms, "720p sum_w")
    del ro, rs
    rdd, rdw = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
    assert_close_sum(d_data, rdd, mdd, "720p d_data")
    assert_close_sum(d_weights, rdw, mdw, "720p d_weights")
    del rdd, rdw
    rg = oracle.scatter2gather(weights)
    assert th.equal(gather.cpu().view(th.int32), rg.view(th.int32))


def test_config2_call_properties():
    """The full config-2 call (N=4: 6.5 GB of weights, byte offsets > 2^32)
    through size-independent properties computed with plain torch ops."""
    n, c, h, w, k = 4, 3, 720, 1280, 21
    c0 = (k - 1) // 2
    g = th.Generator(device=DEV).manual_seed(7)
    data = 2 * th.randn(n, c, h, w, device=DEV, generator=g)
    weights = th.randn(n, k, k, h, w, device=DEV, generator=g)
    d_output = th.randn(n, c, h, w, device=DEV, generator=g)
    d_sum_w = th.randn(n, h, w, device=DEV, generator=g)
    out, sum_w = funcs.KernelWeighting.apply(data, weights)
    assert _lib.last_path() == 1
    # (1) sum_w is the plain tap sum
    ref_sw = weights.double().sum(dim=(1, 2))
    mag_sw = weights.abs().double().sum(dim=(1, 2))
    assert ((sum_w.double() - ref_sw).abs() <= 1e-5 * (ref_sw.abs() + mag_sw)).all()
    del ref_sw, mag_sw
    # (2) output at 4096 random pixels of the LAST image (largest offsets), in float64
    gi = th.Generator().manual_seed(1)
    ys = th.randint(0, h, (4096,), generator=gi).to(DEV)
    xs = th.randint(0, w, (4096,), generator=gi).to(DEV)
    # include the corners / borders explicitly
    ys[:4] = th.tensor([0, 0, h - 1, h - 1]); xs[:4] = th.tensor([0, w - 1, 0, w - 1])
    pad = th.nn.functional.pad(data[n - 1].double(), (c0, c0, c0, c0))
    dy = th.arange(k, device=DEV).view(k, 1, 1)
    dx = th.arange(k, device=DEV).view(1, k, 1)
    patch = pad[:, ys.view(1, 1, -1) + dy, xs.view(1, 1, -1) + dx]     # [c,k,k,P]
    wv = weights[n - 1][:, :, ys, xs].double()                          # [k,k,P]
    ref = (patch * wv).sum(dim=(1, 2))
    mag = (patch.abs() * wv.abs()).sum(dim=(1, 2))
    got = out[n - 1][:, ys, xs].double()
    assert ((got - ref).abs() <= 1e-5 * (ref.abs() + mag)).all()
    del pad, patch, wv
    # (3) backward: adjointness and the tap-sum identity of d_weights
    d_data = th.empty_like(data)
    d_weights = th.empty_like(weights)
    halide_ops.kernel_weighting_grad_cuda_float32(
        data, weights, sum_w, d_output, d_sum_w, d_data, d_weights)
    lhs = (out.double() * d_output.double()).sum().item()
    rhs = (data.double() * d_data.double()).sum().item()
    scale = (out.double().abs() * d_output.double().abs()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * scale, (lhs, rhs, scale)
    # sum over taps of dW = K^2 dSw + sum_c dO_c * boxsum(D_c)
    box = th.nn.functional.avg_pool2d(
        data.double(), k, stride=1, padding=c0, count_include_pad=True) * (k * k)
    ref = k * k * d_sum_w.double() + (box * d_output.double()).sum(dim=1)
    mag = k * k * d_sum_w.double().abs() + (
        th.nn.functional.avg_pool2d(data.double().abs(), k, stride=1, padding=c0) * (k * k)
        * d_output.double().abs()).sum(dim=1)
    got = d_weights.double().sum(dim=(1, 2))
    assert ((got - ref).abs() <= 1e-5 * (ref.abs() + mag)).all()
    del box, ref, mag, got, d_weights
    # (4) scatter2gather: exact index map on the full volume
    gather = funcs.Scatter2Gather.apply(weights)
    for (dy_, dx_) in [(0, 0), (20, 20), (10, 10), (3, 17), (20, 0)]:
        sy, sx = dy_ - c0, dx_ - c0
        src = weights[:, k - 1 - dy_, k - 1 - dx_]
        want = th.zeros_like(src)
        y0, y1 = max(0, -sy), min(h, h - sy)
        x0, x1 = max(0, -sx), min(w, w - sx)
        want[:, y0:y1, x0:x1] = src[:, y0 + sy:y1 + sy, x0 + sx:x1 + sx]
        assert th.equal(gather[:, dy_, dx_].view(th.int32), want.view(th.int32))


def test_random_small_shapes_vs_oracle():
    """Seeded random shapes beyond the fixed list: ragged sizes, even / non-square
    kernels, 1..6 channels; both device implementations against the oracle."""
    import random
    rng = random.Random(1)
    for it in range(16):
        kh, kw = rng.randint(1, 8), rng.randint(1, 8)
        c = rng.randint(1, 6)
        h, w = rng.randint(1, 24), rng.choice([4, 8, 12, 20, 36, 132, rng.randint(1, 50)])
        n = rng.randint(1, 3)
        data, weights, d_output, d_sum_w = make_inputs(n, c, h, w, kh, kw, seed=100 + it)
        mo, ms, mdd, mdw = kw_magnitudes(data, weights, d_output, d_sum_w)
        ro, rs = oracle.kernel_weighting(data, weights)
        rdd, rdw = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
        rg = oracle.scatter2gather(weights)
        for flag in (0, 1):
            prev = _lib.force_generic(flag)
            try:
                out, sum_w, d_data, d_weights, gather = run_cuda(data, weights, d_output, d_sum_w)
            finally:
                _lib.force_generic(prev)
            what = "shape %s path %d" % ((n, c, h, w, kh, kw), flag)
            assert_close_sum(out, ro, mo, what + " output")
            assert_close_sum(sum_w, rs, ms, what + " sum_w")
            assert_close_sum(d_data, rdd, mdd, what + " d_data")
            assert_close_sum(d_weights, rdw, mdw, what + " d_weights")
            assert np.array_equal(gather.cpu().numpy().view(np.uint32), rg.numpy().view(np.uint32)), what
