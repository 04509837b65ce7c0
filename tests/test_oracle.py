"""CPU suite: pins the oracle (oracle/sbmc_oracle.c) against

* the committed golden vectors (tests/golden/*.npz, exact-math float64 values
  of the reference formulas from the independent restatement numpy_ref.py);
* the reference's own analytic known-answer tests (tests/kats.py);
* size-independent properties (adjointness, linearity, S2G index properties).
"""
import glob
import os

import numpy as np
import pytest
import torch as th

import oracle
from oracle import numpy_ref
from tests import kats
from tests.golden import make_golden
from tests.util import RTOL, assert_close_sum, kw_magnitudes, make_inputs

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_golden(path):
    g = np.load(path)
    shape = tuple(int(v) for v in g["shape"])
    data, weights, d_output, d_sum_w = make_golden.inputs(*shape, int(g["seed"]))
    digest = np.asarray([a.astype(np.float64).sum()
                         for a in (data, weights, d_output, d_sum_w)])
    assert np.array_equal(digest, g["input_digest"]), "input generator drifted"
    t = [th.from_numpy(a) for a in (data, weights, d_output, d_sum_w)]
    return g, t


def check_against_golden(g, t, out, sum_w, d_data, d_weights, gather, what):
    data, weights, d_output, d_sum_w = t
    mo, ms, mdd, mdw = kw_magnitudes(data, weights, d_output, d_sum_w)
    assert_close_sum(out, th.from_numpy(g["output"]), mo, what + " output")
    assert_close_sum(sum_w, th.from_numpy(g["sum_w"]), ms, what + " sum_w")
    assert_close_sum(d_data, th.from_numpy(g["d_data"]), mdd, what + " d_data")
    idx = th.from_numpy(g["sample_index"])
    assert_close_sum(d_weights.cpu().reshape(-1)[idx],
                     th.from_numpy(g["d_weights_sample"]),
                     mdw.reshape(-1)[idx], what + " d_weights")
    s = d_weights.double().sum().item()
    assert abs(s - g["d_weights_sum"][0]) <= RTOL * g["d_weights_sum"][1] + 1e-12
    # scatter2gather: bit-exact
    gs = gather.cpu().reshape(-1)[idx].numpy()
    assert np.array_equal(gs.view(np.uint32), g["gather_sample"].view(np.uint32)), \
        what + " gather sample differs"
    assert np.array_equal(make_golden.bit_checksum(gather.cpu().numpy()),
                          g["gather_bits"]), what + " gather checksum differs"


def test_golden_files_present():
    assert len(GOLDEN) == len(make_golden.CASES)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path):
    g, t = load_golden(path)
    data, weights, d_output, d_sum_w = t
    out, sum_w = oracle.kernel_weighting(data, weights)
    d_data, d_weights = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
    gather = oracle.scatter2gather(weights)
    check_against_golden(g, t, out, sum_w, d_data, d_weights, gather, "oracle")


@pytest.mark.parametrize("shape", [
    (1, 1, 1, 1, 1, 1), (2, 3, 5, 7, 3, 5), (1, 4, 11, 19, 6, 2),
    (3, 2, 9, 8, 9, 9), (1, 3, 4, 40, 21, 21)])
def test_oracle_matches_numpy_ref_random(shape):
    n, c, h, w, kh, kw = shape
    data, weights, d_output, d_sum_w = make_inputs(n, c, h, w, kh, kw, seed=11)
    mo, ms, mdd, mdw = kw_magnitudes(data, weights, d_output, d_sum_w)
    out, sum_w = oracle.kernel_weighting(data, weights)
    ro, rs = numpy_ref.kernel_weighting(data.numpy(), weights.numpy())
    assert_close_sum(out, th.from_numpy(ro), mo, "output")
    assert_close_sum(sum_w, th.from_numpy(rs), ms, "sum_w")
    d_data, d_weights = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
    rdd, rdw = numpy_ref.kernel_weighting_grad(
        data.numpy(), weights.numpy(), d_output.numpy(), d_sum_w.numpy())
    assert_close_sum(d_data, th.from_numpy(rdd), mdd, "d_data")
    assert_close_sum(d_weights, th.from_numpy(rdw), mdw, "d_weights")
    g = oracle.scatter2gather(weights)
    assert np.array_equal(g.numpy(), numpy_ref.scatter2gather(weights.numpy()))


def test_oracle_empty_inputs():
    for shape in [(0, 3, 4, 4, 3, 3), (2, 3, 0, 4, 3, 3), (2, 3, 4, 0, 3, 3)]:
        n, c, h, w, kh, kw = shape
        out, sum_w = oracle.kernel_weighting(th.zeros(n, c, h, w), th.zeros(n, kh, kw, h, w))
        assert out.shape == (n, c, h, w) and sum_w.shape == (n, h, w)
        assert oracle.scatter2gather(th.zeros(n, kh, kw, h, w)).numel() == 0


# -- the reference's known-answer tests, run on the oracle ---------------------
def test_kat_forward_impulse():
    KW, _ = kats.oracle_functions()
    kats.kw_forward_impulse(KW, "cpu")


def test_kat_backward_impulse():
    KW, _ = kats.oracle_functions()
    kats.kw_backward_impulse(KW, "cpu")


def test_kat_gradcheck():
    KW, _ = kats.oracle_functions()
    kats.kw_gradcheck(KW, "cpu")


def test_kat_scatter2gather_index_map():
    _, S2G = kats.oracle_functions()
    kats.s2g_index_map(S2G, "cpu", stride=7)


def test_kat_scatter2gather_gradcheck():
    _, S2G = kats.oracle_functions()
    kats.s2g_gradcheck(S2G, "cpu")


# -- properties ----------------------------------------------------------------
def test_sum_w_counts_out_of_image_taps():
    """sum_w sums ALL taps, including those whose data tap is outside the image
    (homogeneous 1.0f channel, src/kernel_weighting.cpp:49)."""
    weights = th.ones(1, 5, 5, 6, 6)
    out, sum_w = oracle.kernel_weighting(th.ones(1, 1, 6, 6), weights)
    assert (sum_w == 25).all()
    assert out[0, 0, 0, 0].item() == 9 and out[0, 0, 3, 3].item() == 25


def test_adjointness_odd_kernels():
    """<KW(D, W), dO> == <D, dD>  and  dW = dSw + D (x) dO  for odd kernels."""
    data, weights, d_output, d_sum_w = make_inputs(2, 3, 12, 16, 5, 7, seed=5)
    out, _ = oracle.kernel_weighting(data, weights)
    d_data, _ = oracle.kernel_weighting_grad(data, weights, d_output, d_sum_w)
    lhs = (out.double() * d_output.double()).sum().item()
    rhs = (data.double() * d_data.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * (abs(lhs) + 1e3)


def test_scatter2gather_self_adjoint_and_interior_involution():
    s = th.randn(2, 5, 5, 12, 14)
    t = th.randn(2, 5, 5, 12, 14)
    lhs = (oracle.scatter2gather(s).double() * t.double()).sum().item()
    rhs = (s.double() * oracle.scatter2gather(t).double()).sum().item()
    assert abs(lhs - rhs) < 1e-9 * (abs(lhs) + 1)
    # applying it twice restores every tap whose partner pixel is in the image
    gg = oracle.scatter2gather(oracle.scatter2gather(s))
    assert np.array_equal(gg[:, 2, 2].numpy(), s[:, 2, 2].numpy())
    assert np.array_equal(gg[:, :, :, 2:-2, 2:-2].numpy(), s[:, :, :, 2:-2, 2:-2].numpy())
    assert (gg[:, 0, 0, :2, :] == 0).all()
