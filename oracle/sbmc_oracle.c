/*
 * sbmc_oracle.c -- CPU restatement of the SBMC kernel-splatting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sbmc_b200/ may import, link or
 * call this file.  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker and as the
 * timed CPU baseline, never as the product.
 *
 * What it restates (reference = adobe/sbmc, paths relative to /root/reference):
 *   kernel_weighting        src/kernel_weighting.cpp:27-64   (algorithm)
 *                           src/kernel_weighting.cpp:164-187 (CPU schedule)
 *   kernel_weighting_grad   src/kernel_weighting.cpp:67-124  (algorithm)
 *                           src/kernel_weighting.cpp:219-235 (CPU schedule)
 *   scatter2gather          src/scatter2gather.cpp:28-52     (algorithm)
 *                           src/scatter2gather.cpp:81-89     (CPU schedule)
 * Index order is the torch one used by sbmc/functions.py:42-49,78-86:
 *   data[n,c,y,x]  weights[n,dy,dx,y,x]  (Halide sees them reversed,
 *   src/kernel_weighting.cpp:130-133).
 *
 * The arithmetic lives in the reference tree (Halide algorithm text); the code
 * generator that turns it into machine code is Halide v8.0.0, which is not in
 * /root/reference and cannot be installed here.  PARITY PIN: the reference's
 * own test files run unmodified against this oracle (tests/test_reference_suite.py:
 * tests/test_functions.py, tests/test_modules.py of the reference), plus an
 * independent float64 restatement (oracle/numpy_ref.py).  Dense-random outputs of
 * the Halide binary itself are NOT available: "parity unpinned" for those beyond
 * the known-answer tests.
 *
 * Numerics: fp32 multiply then fp32 add (compile with -ffp-contract=off),
 * reduction order ry outer / rx inner, sequential, exactly the RDom order of
 * src/kernel_weighting.cpp:45-57; d_weights starts from d_sum_w and adds the
 * channels in increasing order (src/kernel_weighting.cpp:113-116).
 *
 * Schedule fidelity (for the timed CPU baseline): the forward computes the
 * C+1 "homogeneous" channels of `summed` one (c, n) at a time, parallel over
 * blocks of 8 rows, x vectorised by 8, then copies into output / sum_w --
 * i.e. the weights are streamed C+1 times, like the reference CPU schedule.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VEC 8

typedef int64_t i64;

int sbmc_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void sbmc_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- forward ------------------------------------------------------------ */
/* summed(x,y,c,n) += w(x,y,rx,ry,n) * homogeneous(x+rx-(kw-1)/2, y+ry-(kh-1)/2, c, n)
 * src/kernel_weighting.cpp:45-57.  c == C is the homogeneous 1.0f channel. */
static void fwd_rows(const float *Dn, const float *Wn, float *Sc, int c, int C,
                     i64 H, i64 W, int KH, int KW, i64 y0, i64 y1) {
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  for (i64 y = y0; y < y1; ++y) {
    for (i64 x = 0; x < W; x += VEC) {
      const int vl = (int)((W - x) < VEC ? (W - x) : VEC);
      float acc[VEC];
      for (int v = 0; v < VEC; ++v) acc[v] = 0.0f;
      for (int ry = 0; ry < KH; ++ry) {
        const i64 yy = y + ry - c0h;
        const int row_in = (yy >= 0 && yy < H);
        for (int rx = 0; rx < KW; ++rx) {
          const float *wp = Wn + (((i64)ry * KW + rx) * H + y) * W + x;
          const i64 xx = x + rx - c0w;
          if (c == C) {
#pragma omp simd
            for (int v = 0; v < vl; ++v) acc[v] = acc[v] + wp[v] * 1.0f;
          } else if (row_in && xx >= 0 && xx + vl <= W) {
            const float *dp = Dn + ((i64)c * H + yy) * W + xx;
#pragma omp simd
            for (int v = 0; v < vl; ++v) acc[v] = acc[v] + wp[v] * dp[v];
          } else {
            for (int v = 0; v < vl; ++v) {
              const i64 xv = xx + v;
              const float h = (row_in && xv >= 0 && xv < W)
                                  ? Dn[((i64)c * H + yy) * W + xv]
                                  : 0.0f; /* constant_exterior, :35-36 */
              acc[v] = acc[v] + wp[v] * h;
            }
          }
        }
      }
      for (int v = 0; v < vl; ++v) Sc[y * W + x + v] = acc[v];
    }
  }
}

/* data[N,C,H,W] weights[N,KH,KW,H,W] -> output[N,C,H,W] sum_w[N,H,W] */
int sbmc_oracle_kernel_weighting_f32(const float *data, const float *weights,
                                     float *output, float *sum_w, i64 N, int C,
                                     i64 H, i64 W, int KH, int KW) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || KH < 0 || KW < 0) return -1;
  if (N == 0 || H == 0 || W == 0) return 0;
  const i64 plane = H * W;
  float *summed = (float *)malloc(sizeof(float) * (size_t)(plane * (C + 1)));
  if (!summed) return -2;
  const i64 nblk = (H + 7) / 8;
  for (i64 n = 0; n < N; ++n) {
    const float *Dn = data + n * C * plane;
    const float *Wn = weights + n * (i64)KH * KW * plane;
    for (int c = 0; c <= C; ++c) { /* serial c, n; parallel(y, 8): :180-187 */
      float *Sc = summed + (i64)c * plane;
#pragma omp parallel for schedule(dynamic, 1)
      for (i64 b = 0; b < nblk; ++b) {
        const i64 y0 = b * 8, y1 = (y0 + 8 < H) ? y0 + 8 : H;
        fwd_rows(Dn, Wn, Sc, c, C, H, W, KH, KW, y0, y1);
      }
    }
    /* output(x,y,c,n) = summed(x,y,c,n); sum_w = summed(x,y,channels,n) :59-60 */
    memcpy(output + n * C * plane, summed, sizeof(float) * (size_t)(plane * C));
    memcpy(sum_w + n * plane, summed + (i64)C * plane,
           sizeof(float) * (size_t)plane);
  }
  free(summed);
  return 0;
}

/* ---- backward ----------------------------------------------------------- */
static inline float at0(const float *p, i64 H, i64 W, i64 y, i64 x) {
  return (y >= 0 && y < H && x >= 0 && x < W) ? p[y * W + x] : 0.0f;
}

/* d_data_tmp(x,y,c,n) += f_weights(x+rx-pw, y+ry-ph, kw-1-rx, kh-1-ry, n)
 *                        * f_d_output(x+rx-pw, y+ry-ph, c, n)
 * src/kernel_weighting.cpp:91-105 */
static void ddata_row(const float *Wn, const float *dOc, float *dDc, i64 H,
                      i64 W, int KH, int KW, i64 y) {
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  for (i64 x = 0; x < W; x += VEC) {
    const int vl = (int)((W - x) < VEC ? (W - x) : VEC);
    float acc[VEC];
    for (int v = 0; v < VEC; ++v) acc[v] = 0.0f;
    for (int ry = 0; ry < KH; ++ry) {
      const i64 yy = y + ry - c0h;
      const int row_in = (yy >= 0 && yy < H);
      for (int rx = 0; rx < KW; ++rx) {
        const i64 xx = x + rx - c0w;
        const float *wp =
            Wn + ((i64)(KH - 1 - ry) * KW + (KW - 1 - rx)) * H * W;
        if (row_in && xx >= 0 && xx + vl <= W) {
          const float *w2 = wp + yy * W + xx;
          const float *o2 = dOc + yy * W + xx;
#pragma omp simd
          for (int v = 0; v < vl; ++v) acc[v] = acc[v] + w2[v] * o2[v];
        } else {
          for (int v = 0; v < vl; ++v)
            acc[v] = acc[v] + at0(wp, H, W, yy, xx + v) * at0(dOc, H, W, yy, xx + v);
        }
      }
    }
    for (int v = 0; v < vl; ++v) dDc[y * W + x + v] = acc[v];
  }
}

/* d_weights_tmp(x,y,dx,dy,n) = d_sum_w(x,y,n);
 * d_weights_tmp += f_data(x+dx-pw, y+dy-ph, rchan, n) * f_d_output(x,y,rchan,n)
 * src/kernel_weighting.cpp:111-117 */
static void dweights_row(const float *Dn, const float *dOn, const float *dSwn,
                         float *dWp, int C, i64 H, i64 W, int dy, int dx,
                         int c0h, int c0w, i64 y) {
  const i64 yy = y + dy - c0h;
  const int row_in = (yy >= 0 && yy < H);
  for (i64 x = 0; x < W; x += VEC) {
    const int vl = (int)((W - x) < VEC ? (W - x) : VEC);
    const i64 xx = x + dx - c0w;
    float acc[VEC];
    for (int v = 0; v < vl; ++v) acc[v] = dSwn[y * W + x + v];
    for (int c = 0; c < C; ++c) {
      const float *oc = dOn + ((i64)c * H + y) * W + x;
      if (row_in && xx >= 0 && xx + vl <= W) {
        const float *dc = Dn + ((i64)c * H + yy) * W + xx;
#pragma omp simd
        for (int v = 0; v < vl; ++v) acc[v] = acc[v] + dc[v] * oc[v];
      } else {
        for (int v = 0; v < vl; ++v)
          acc[v] = acc[v] + at0(Dn + (i64)c * H * W, H, W, yy, xx + v) * oc[v];
      }
    }
    for (int v = 0; v < vl; ++v) dWp[y * W + x + v] = acc[v];
  }
}

/* sum_w is an input of the reference pipeline but is never read
 * (src/kernel_weighting.cpp:67-124); accepted and ignored here as well. */
int sbmc_oracle_kernel_weighting_grad_f32(
    const float *data, const float *weights, const float *sum_w,
    const float *d_output, const float *d_sum_w, float *d_data,
    float *d_weights, i64 N, int C, i64 H, i64 W, int KH, int KW) {
  (void)sum_w;
  if (N < 0 || C < 0 || H < 0 || W < 0 || KH < 0 || KW < 0) return -1;
  if (N == 0 || H == 0 || W == 0) return 0;
  const i64 plane = H * W;
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  /* d_data: fuse(c,n) fuse(y,cn) parallel(.,8) vectorize(x,8)  :220-226 */
  const i64 rows_d = H * C * N;
#pragma omp parallel for schedule(dynamic, 8)
  for (i64 t = 0; t < rows_d; ++t) {
    const i64 y = t % H;
    const i64 cn = t / H;
    const i64 c = cn % C, n = cn / C;
    ddata_row(weights + n * (i64)KH * KW * plane,
              d_output + (n * C + c) * plane, d_data + (n * C + c) * plane, H,
              W, KH, KW, y);
  }
  /* d_weights: fuse(dx,dy) fuse(y,dxdy) fuse(.,n) parallel(.,8) :228-235 */
  const i64 rows_w = H * (i64)KH * KW * N;
#pragma omp parallel for schedule(dynamic, 8)
  for (i64 t = 0; t < rows_w; ++t) {
    const i64 y = t % H;
    const i64 r = t / H;
    const i64 tap = r % ((i64)KH * KW), n = r / ((i64)KH * KW);
    const int dy = (int)(tap / KW), dx = (int)(tap % KW);
    dweights_row(data + n * C * plane, d_output + n * C * plane,
                 d_sum_w + n * plane,
                 d_weights + (n * (i64)KH * KW + tap) * plane, C, H, W, dy, dx,
                 c0h, c0w, y);
  }
  return 0;
}

/* ---- scatter2gather ----------------------------------------------------- */
/* output(x,y,dx,dy,n) = f_weights(x+dx-(kw-1)/2, y+dy-(kh-1)/2, kw-1-dx, kh-1-dy, n)
 * src/scatter2gather.cpp:40-47; zero outside the image (:34-35). */
int sbmc_oracle_scatter2gather_f32(const float *weights, float *output, i64 N,
                                   int KH, int KW, i64 H, i64 W) {
  if (N < 0 || H < 0 || W < 0 || KH < 0 || KW < 0) return -1;
  if (N == 0 || H == 0 || W == 0) return 0;
  const i64 plane = H * W;
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  const i64 rows = H * (i64)KH * KW * N;
#pragma omp parallel for schedule(dynamic, 8)
  for (i64 t = 0; t < rows; ++t) {
    const i64 y = t % H;
    const i64 r = t / H;
    const i64 tap = r % ((i64)KH * KW), n = r / ((i64)KH * KW);
    const int dy = (int)(tap / KW), dx = (int)(tap % KW);
    const float *src =
        weights + (n * (i64)KH * KW + (i64)(KH - 1 - dy) * KW + (KW - 1 - dx)) * plane;
    float *dst = output + (n * (i64)KH * KW + tap) * plane + y * W;
    const i64 yy = y + dy - c0h;
    if (yy < 0 || yy >= H) {
      for (i64 x = 0; x < W; ++x) dst[x] = 0.0f;
      continue;
    }
    const i64 sh = dx - c0w; /* dst[x] = src[yy, x + sh] */
    i64 xa = sh < 0 ? -sh : 0;         /* first x with x+sh >= 0 */
    i64 xb = sh > 0 ? W - sh : W;      /* one past last x with x+sh < W */
    if (xa > W) xa = W;
    if (xb < xa) xb = xa;
    for (i64 x = 0; x < xa; ++x) dst[x] = 0.0f;
    if (xb > xa)
      memcpy(dst + xa, src + yy * W + xa + sh, sizeof(float) * (size_t)(xb - xa));
    for (i64 x = xb; x < W; ++x) dst[x] = 0.0f;
  }
  return 0;
}
