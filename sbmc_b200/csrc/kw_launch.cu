// kw_launch.cu -- shape dispatch for KernelWeighting forward / backward.
#include "kw_launch.cuh"

namespace sbmc {

// Tuned configurations (see profiles/ for the sweep they come from).
constexpr int kRows = 8;      // warps (rows) per CTA, forward
constexpr int kRowsDw = 16;   // warps (rows) per CTA, d_weights
constexpr int kMinB = 2;      // resident CTAs per SM the register budget allows
constexpr int kChunk = 7;     // taps of one dx-chunk kept in flight per thread

static inline bool aligned16(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

template <int C, int KW>
static int fwd_tuned(const float *data_ext, const float *weights, float *output,
                     float *sum_w, i64 n, i64 h, i64 w, int kh, int halo_top,
                     int halo_bot, cudaStream_t st) {
  constexpr int CH = KW < kChunk ? KW : kChunk;
  return run_fwd<C, KW, kRows, kMinB, CH>(data_ext, weights, output, sum_w, n, h,
                                          w, kh, halo_top, halo_bot, st);
}

template <int C, int KW, int NSEG>
static int ddata_tuned_n(const float *weights, const float *d_output,
                         float *d_data_ext, i64 n, i64 h, i64 w, int kh,
                         int halo_top, int halo_bot, cudaStream_t st) {
  constexpr int CH = KW < kChunk ? KW : kChunk;
  constexpr int MINB = 4;
  return run_bwd_ddata<C, KW, NSEG, MINB, CH>(weights, d_output, d_data_ext, n, h,
                                              w, kh, halo_top, halo_bot, st);
}

template <int C, int KW>
static int ddata_tuned(const float *weights, const float *d_output,
                       float *d_data_ext, i64 n, i64 h, i64 w, int kh,
                       int halo_top, int halo_bot, cudaStream_t st) {
  const i64 segs = ceil_div(w, kTileW);
  // Two 128-pixel segments per CTA (64 threads, 4 CTAs per SM) measured fastest
  // on B200 even though wider images then need the seam atomics
  // (profiles/r1d_sweep.txt: 0.951 ms vs 1.010 ms for one 1280-pixel CTA).
#define SBMC_DD(NSEG)                                                         \
  return ddata_tuned_n<C, KW, NSEG>(weights, d_output, d_data_ext, n, h, w, kh, \
                                    halo_top, halo_bot, st)
  if (segs <= 1) SBMC_DD(1);
  SBMC_DD(2);
#undef SBMC_DD
}

template <int C, int KW>
static int dweights_tuned(const float *data_ext, const float *d_output,
                          const float *d_sum_w, float *d_weights, i64 n, i64 h,
                          i64 w, int kh, int halo_top, int halo_bot,
                          cudaStream_t st) {
  constexpr int CH = KW < kChunk ? KW : kChunk;
  // 16 rows per CTA, one CTA per SM, st.global.cs (profiles/r1d_sweep.txt,
  // r1u_sweep.txt: 1.022 ms vs 1.052 ms for 8 rows; L2 eviction hints on the
  // operands / stores did not move the write rate: 1.033 ms)
  return run_bwd_dweights<C, KW, kRowsDw, 1, CH, 0>(
      data_ext, d_output, d_sum_w, d_weights, n, h, w, kh, halo_top, halo_bot, st);
}

// (C, KW) pairs with a tuned instantiation: the SBMC model (3, 21), the
// reference's own test shapes (tests/test_functions.py:43-144: C in {3, 5},
// K in {3, 5, 7}) and config 1 of BASELINE.json (3, 5).
#define SBMC_TUNED_SHAPES(X)                                                  \
  X(3, 21) X(3, 5) X(3, 3) X(3, 7) X(5, 5) X(5, 3)                            \
  X(3, 9) X(3, 11) X(3, 13) X(3, 15) X(3, 17) X(3, 19)

int launch_fwd(const float *data_ext, const float *weights, float *output,
               float *sum_w, i64 n, int c, i64 h, i64 w, int kh, int kw,
               int halo_top, int halo_bot, cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  const bool vec_ok = !force_generic() && aligned16(data_ext) &&
                      aligned16(weights) && aligned16(output) && aligned16(sum_w);
#define X(CC, KK)                                                              \
  if (vec_ok && c == CC && kw == KK && tile_shape_ok<CC, KK, kRows>(n, h, w, kh, hext)) { \
    note_path(1);                                                              \
    return fwd_tuned<CC, KK>(data_ext, weights, output, sum_w, n, h, w, kh,    \
                             halo_top, halo_bot, st);                          \
  }
  SBMC_TUNED_SHAPES(X)
#undef X
  note_path(2);
  warn_generic("kernel_weighting", c, kh, kw, w);
  return generic_fwd(data_ext, weights, output, sum_w, n, c, h, w, kh, kw,
                     halo_top, halo_bot, st);
}

int launch_bwd_dweights(const float *data_ext, const float *d_output,
                        const float *d_sum_w, float *d_weights, i64 n, int c,
                        i64 h, i64 w, int kh, int kw, int halo_top, int halo_bot,
                        cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  const bool vec_ok = !force_generic() && aligned16(data_ext) &&
                      aligned16(d_output) && aligned16(d_sum_w) &&
                      aligned16(d_weights);
#define X(CC, KK)                                                              \
  if (vec_ok && c == CC && kw == KK && tile_shape_ok<CC, KK, kRowsDw>(n, h, w, kh, hext)) { \
    note_path(1);                                                              \
    return dweights_tuned<CC, KK>(data_ext, d_output, d_sum_w, d_weights, n, h, \
                                  w, kh, halo_top, halo_bot, st);              \
  }
  SBMC_TUNED_SHAPES(X)
#undef X
  note_path(2);
  warn_generic("kernel_weighting_grad (d_weights)", c, kh, kw, w);
  return generic_bwd_dweights(data_ext, d_output, d_sum_w, d_weights, n, c, h, w,
                              kh, kw, halo_top, halo_bot, st);
}

int launch_bwd_ddata(const float *weights, const float *d_output,
                     float *d_data_ext, i64 n, int c, i64 h, i64 w, int kh,
                     int kw, int halo_top, int halo_bot, cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  const bool vec_ok = !force_generic() && aligned16(weights) &&
                      aligned16(d_output) && (w % 4 == 0) &&
                      w < (1ll << 31) - 4096 && hext < (1ll << 31) - 4096;
#define X(CC, KK)                                                              \
  if (vec_ok && c == CC && kw == KK) {                                         \
    note_path(1);                                                              \
    return ddata_tuned<CC, KK>(weights, d_output, d_data_ext, n, h, w, kh,     \
                               halo_top, halo_bot, st);                        \
  }
  SBMC_TUNED_SHAPES(X)
#undef X
  note_path(2);
  warn_generic("kernel_weighting_grad (d_data)", c, kh, kw, w);
  return generic_bwd_ddata(weights, d_output, d_data_ext, n, c, h, w, kh, kw,
                           halo_top, halo_bot, st);
}

}  // namespace sbmc
