"""Generates the golden vectors under tests/golden/ (run from the repo root:
``python tests/golden/make_golden.py``).

The reference's native ops cannot be built or imported here (they need a Halide
v8.0.0 distribution), so the golden OUTPUTS are produced
by the independent float64 restatement ``oracle/numpy_ref.py`` (exact-math
values of the reference formulas, rounded once to float32), and the golden
INPUTS come from numpy's stream-stable legacy ``RandomState``.  Both the fp32 C
oracle and the CUDA kernels are checked against these files; Scatter2Gather is
a pure copy, so its golden output is bit-exact.

Cases: BASELINE.json config 1 (B*spp=2, C=3, 64x64, K=5), the reference's own
test shapes (tests/test_functions.py: C=5 K=5 16x16; C=3 K=3 16x16), the model
kernel size K=21 on a small ragged image, even / non-square kernels and a width
that is not a multiple of 4 (generic path).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import numpy_ref  # noqa: E402

CASES = {
    # name: (n, c, h, w, kh, kw, seed)
    "cfg1_n2_c3_64x64_k5": (2, 3, 64, 64, 5, 5, 1),
    "ref_c5_16x16_k5": (4, 5, 16, 16, 5, 5, 2),
    "ref_c3_16x16_k3": (2, 3, 16, 16, 3, 3, 3),
    "model_k21_c3_24x136": (1, 3, 24, 136, 21, 21, 4),
    "even_k4x6_c2_13x20": (2, 2, 13, 20, 4, 6, 5),
    "ragged_w_c3_9x17_k7": (1, 3, 9, 17, 7, 7, 6),
    "k1_c1_8x8": (1, 1, 8, 8, 1, 1, 7),
}


def inputs(n, c, h, w, kh, kw, seed):
    rs = np.random.RandomState(seed)
    data = (2 * rs.standard_normal((n, c, h, w))).astype(np.float32)
    weights = rs.standard_normal((n, kh, kw, h, w)).astype(np.float32)
    d_output = rs.standard_normal((n, c, h, w)).astype(np.float32)
    d_sum_w = rs.standard_normal((n, h, w)).astype(np.float32)
    return data, weights, d_output, d_sum_w


def sample_index(size, seed, count=16384):
    """All indices if the tensor is small, else a fixed pseudo-random subset."""
    if size <= count:
        return np.arange(size, dtype=np.int64)
    rs = np.random.RandomState(seed + 1000)
    return np.sort(rs.choice(size, size=count, replace=False)).astype(np.int64)


def bit_checksum(a):
    """(sum mod 2^64, xor) of the uint32 bit patterns: exact, order-free."""
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel()
    total = int(bits.astype(np.uint64).sum(dtype=np.uint64))
    return np.asarray([total, int(np.bitwise_xor.reduce(bits))], dtype=np.uint64)


def main():
    for name, shape in CASES.items():
        data, weights, d_output, d_sum_w = inputs(*shape)
        out, sum_w = numpy_ref.kernel_weighting(data, weights)
        d_data, d_weights = numpy_ref.kernel_weighting_grad(
            data, weights, d_output, d_sum_w)
        gather = numpy_ref.scatter2gather(weights)
        assert gather.dtype == np.float32
        idx = sample_index(d_weights.size, shape[6])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            shape=np.asarray(shape[:6], dtype=np.int64), seed=shape[6],
            # inputs are regenerated from the seed (inputs() above); a digest
            # guards against a drifting generator
            input_digest=np.asarray([data.astype(np.float64).sum(),
                                     weights.astype(np.float64).sum(),
                                     d_output.astype(np.float64).sum(),
                                     d_sum_w.astype(np.float64).sum()]),
            output=out.astype(np.float32), sum_w=sum_w.astype(np.float32),
            d_data=d_data.astype(np.float32),
            # the K*K-sized outputs: a fixed pseudo-random sample + checksums
            sample_index=idx,
            d_weights_sample=d_weights.ravel()[idx].astype(np.float32),
            d_weights_sum=np.asarray([d_weights.sum(), np.abs(d_weights).sum()]),
            gather_sample=gather.ravel()[idx],
            # order-independent exact checksum of the bit patterns
            gather_bits=bit_checksum(gather))
        print(name, os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    main()
