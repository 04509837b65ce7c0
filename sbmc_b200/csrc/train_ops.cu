// train_ops.cu -- the memory-bound passes of the mixed-precision TRAINING pipeline
// (sbmc_b200/train_pipeline.py) on bf16 channels-innermost tensors; each replaces a
// chain of eager autograd kernels of the reference's training step
// (sbmc/models.py:171-209 driven by sbmc/interfaces.py:78-106):
//
//   spp_reduce      out[b][p][:]   = scale * sum_s in[b][s][p][:]          (`features.mean(1)`,
//                   models.py:181, and the sum over the samples in the backward of the
//                   broadcast `propagated.unsqueeze(1).repeat(...)`, models.py:175,193)
//   bcast_add       out[b][s][p][:] = a[b][s][p][:] + scale * r[b][p][:]   (backward of the mean)
//   maxpool2x2_bwd  gradient of MaxPool2d(2, 2) routed to the first maximum of each
//                   window, plus the skip-connection gradient, times the derivative of
//                   the activation that produced the pooled tensor (modules.py:296-319)
//   upsample_bwd    transpose of the bilinear upsampling (align_corners = False) of the
//                   decoder, times the activation derivative of the coarse level's output
//   dact            g * act'(y)
//   colsum          fp32 column sums of a bf16 matrix (bias gradients), deterministic
//
// One thread per 16-byte chunk (8 channels); pure HBM / L2 streams.
#include <cuda_bf16.h>

#include "common.cuh"

namespace sbmc {
namespace tr {

__device__ __forceinline__ void unpack8(const uint4 &q, float (&v)[8]) {
  const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(p[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 q;
  __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return q;
}
// derivative factor of ReLU (act 1) / LeakyReLU(0.01) (act 2) from the sign of the output
__device__ __forceinline__ float dact(float y, int act) {
  return (act == 0 || y > 0.f) ? 1.f : (act == 2 ? 0.01f : 0.f);
}

__global__ void __launch_bounds__(256)
spp_reduce_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, float *__restrict__ out32,
                  i64 n_img, int spp, i64 hwc8, float scale) {
  const i64 total = n_img * hwc8;
  for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < total; i += (i64)gridDim.x * 256) {
    const i64 b = i / hwc8, r = i - b * hwc8;
    const uint4 *p = in + b * spp * hwc8 + r;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < spp; ++s) {
      float v[8];
      unpack8(__ldg(p + (i64)s * hwc8), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= scale;
    if (out32) {
      float4 *o = reinterpret_cast<float4 *>(out32) + 2 * i;
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
      out[i] = pack8(acc);
    }
  }
}

__global__ void __launch_bounds__(256)
bcast_add_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ r, uint4 *__restrict__ out,
                 i64 n_img, int spp, i64 hwc8, float scale) {
  const i64 total = n_img * hwc8;
  for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < total; i += (i64)gridDim.x * 256) {
    const i64 b = i / hwc8, rem = i - b * hwc8;
    float rv[8];
    unpack8(__ldg(r + i), rv);
#pragma unroll
    for (int k = 0; k < 8; ++k) rv[k] *= scale;
    for (int s = 0; s < spp; ++s) {
      const i64 j = (b * spp + s) * hwc8 + rem;
      float v[8];
      if (a) {
        unpack8(__ldg(a + j), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += rv[k];
        out[j] = pack8(v);
      } else {
        out[j] = pack8(rv);
      }
    }
  }
}

// x: pooled tensor's source [n][h][w][c] (the activation output of `left`); dpool
// [n][h/2][w/2][c]; dskip: rows of `pitch8` chunks starting at the skip channels (or null).
__global__ void __launch_bounds__(256)
maxpool2x2_bwd_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ dpool,
                      const uint4 *__restrict__ dskip, i64 skip_pitch8, uint4 *__restrict__ out,
                      int h, int w, int c8, int act) {
  const int ho = h / 2, wo = w / 2;
  const int hb = (h + 1) / 2, wb = (w + 1) / 2;       // windows incl. the unpooled last row / col
  const int yb = blockIdx.x % hb;
  const i64 img = blockIdx.x / hb;
  const int total = wb * c8;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += gridDim.y * blockDim.x) {
    const int xb = idx / c8, c = idx - xb * c8;
    const bool pooled = yb < ho && xb < wo;
    float xv[4][8], g[8];
    bool inside[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = 2 * yb + (k >> 1), xx = 2 * xb + (k & 1);
      inside[k] = yy < h && xx < w;
      if (inside[k]) unpack8(__ldg(x + ((img * h + yy) * (i64)w + xx) * c8 + c), xv[k]);
    }
    if (pooled) unpack8(__ldg(dpool + ((img * ho + yb) * (i64)wo + xb) * c8 + c), g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!inside[k]) continue;
      const int yy = 2 * yb + (k >> 1), xx = 2 * xb + (k & 1);
      const i64 pix = (img * h + yy) * (i64)w + xx;
      float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (dskip) unpack8(__ldg(dskip + pix * skip_pitch8 + c), d);
      if (pooled) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // torch's MaxPool2d keeps the FIRST maximum in row-major order (strict `>`)
          bool first = true;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < k) first = first && (xv[j][e] < xv[k][e]);
            else if (j > k) first = first && (xv[j][e] <= xv[k][e]);
          }
          if (first) d[e] += g[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] *= dact(xv[k][e], act);
      out[pix * c8 + c] = pack8(d);
    }
  }
}

// dup: gradient of the upsampled half, rows of `pitch8` chunks [n][h][w]; coarse: the
// tensor that was upsampled [n][hl][wl][c] (its sign selects the activation derivative).
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const uint4 *__restrict__ dup, i64 pitch8, const uint4 *__restrict__ coarse,
                    uint4 *__restrict__ out, int hl, int wl, int h, int w, int c8, float sy_scale,
                    float sx_scale, int ry, int rx, int act) {
  const int yl = blockIdx.x % hl;
  const i64 img = blockIdx.x / hl;
  const int total = wl * c8;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += gridDim.y * blockDim.x) {
    const int xl = idx / c8, c = idx - xl * c8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // fine rows / columns whose two source taps can include (yl, xl)
    const int ylo = max(0, (int)floorf((yl - 1) / sy_scale) - 1), yhi = min(h - 1, ylo + ry);
    const int xlo = max(0, (int)floorf((xl - 1) / sx_scale) - 1), xhi = min(w - 1, xlo + rx);
    for (int y = ylo; y <= yhi; ++y) {
      float sy = sy_scale * (y + 0.5f) - 0.5f;
      sy = sy < 0.f ? 0.f : sy;
      const int y0 = (int)sy, y1 = y0 + (y0 < hl - 1 ? 1 : 0);
      const float ly = sy - y0;
      const float wy = (y0 == yl ? 1.f - ly : 0.f) + (y1 == yl ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int x = xlo; x <= xhi; ++x) {
        float sx = sx_scale * (x + 0.5f) - 0.5f;
        sx = sx < 0.f ? 0.f : sx;
        const int x0 = (int)sx, x1 = x0 + (x0 < wl - 1 ? 1 : 0);
        const float lx = sx - x0;
        const float wx = (x0 == xl ? 1.f - lx : 0.f) + (x1 == xl ? lx : 0.f);
        if (wx == 0.f) continue;
        float v[8];
        unpack8(__ldg(dup + ((img * h + y) * (i64)w + x) * pitch8 + c), v);
        const float wgt = wy * wx;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += wgt * v[e];
      }
    }
    const i64 o = ((img * hl + yl) * (i64)wl + xl) * c8 + c;
    if (act != 0) {
      float m[8];
      unpack8(__ldg(coarse + o), m);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= dact(m[e], act);
    }
    out[o] = pack8(acc);
  }
}

__global__ void __launch_bounds__(256)
dact_kernel(const uint4 *__restrict__ y, const uint4 *__restrict__ g, uint4 *__restrict__ out,
            i64 total8, int act) {
  for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < total8; i += (i64)gridDim.x * 256) {
    float a[8], b[8];
    unpack8(__ldg(y + i), a);
    unpack8(__ldg(g + i), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) b[e] *= dact(a[e], act);
    out[i] = pack8(b);
  }
}

// partial[blk][c] = sum of the block's rows; 256 threads = (256 / c8) row lanes x c8 chunks
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const uint4 *__restrict__ x, i64 pitch8, i64 rows, int c8,
                      float *__restrict__ partial) {
  extern __shared__ float red[];            // [lanes][c8 * 8]
  const int lanes = 256 / c8;
  const int rl = threadIdx.x / c8, c = threadIdx.x - rl * c8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    const i64 per = (rows + gridDim.x - 1) / gridDim.x;
    const i64 lo = blockIdx.x * per, hi = (lo + per < rows) ? lo + per : rows;
    for (i64 r = lo + rl; r < hi; r += lanes) {
      float v[8];
      unpack8(__ldg(x + r * pitch8 + c), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[(rl * c8 + c) * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c8 * 8; i += 256) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * c8 * 8 + i];
    partial[(i64)blockIdx.x * c8 * 8 + i] = s;
  }
}
// 32 consecutive channels x 8 warps; warp w adds partials w, w + 8, ...
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float *__restrict__ partial, int nblk, int c, float *__restrict__ out) {
  __shared__ float red[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f;
  if (i < c) {
    int b = warp;
    for (; b + 8 < nblk; b += 16) {
      a0 += partial[(i64)b * c + i];
      a1 += partial[(i64)(b + 8) * c + i];
    }
    for (; b < nblk; b += 8) a0 += partial[(i64)b * c + i];
  }
  red[warp][lane] = a0 + a1;
  __syncthreads();
  if (warp == 0 && i < c) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[w][lane];
    out[i] = acc;
  }
}

static unsigned grid_for(i64 total, int cap = 148 * 16) {
  i64 b = (total + 255) / 256;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (unsigned)b;
}

}  // namespace tr
}  // namespace sbmc

using namespace sbmc;

static bool aligned16(const void *a, const void *b = nullptr, const void *c = nullptr,
                      const void *d = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
           reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

extern "C" int sbmc_spp_reduce_nhwc_bf16(const void *in, void *out, int out_f32, int64_t n_img,
                                         int spp, int64_t hw, int c, float scale, void *stream) {
  if (n_img < 0 || spp < 1 || hw < 0 || c < 8 || c % 8) {
    set_error("spp_reduce: invalid shape");
    return SBMC_EINVAL;
  }
  if (n_img == 0 || hw == 0) return SBMC_OK;
  if (!in || !out || !aligned16(in, out)) {
    set_error("spp_reduce: null or unaligned pointer");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const i64 hwc8 = hw * (c / 8);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    tr::spp_reduce_kernel<<<tr::grid_for(n_img * hwc8), 256, 0, st>>>(
        static_cast<const uint4 *>(in), out_f32 ? nullptr : static_cast<uint4 *>(out),
        out_f32 ? static_cast<float *>(out) : nullptr, n_img, spp, hwc8, scale);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

extern "C" int sbmc_bcast_add_nhwc_bf16(const void *a, const void *r, void *out, int64_t n_img,
                                        int spp, int64_t hw, int c, float scale, void *stream) {
  if (n_img < 0 || spp < 1 || hw < 0 || c < 8 || c % 8) {
    set_error("bcast_add: invalid shape");
    return SBMC_EINVAL;
  }
  if (n_img == 0 || hw == 0) return SBMC_OK;
  if (!r || !out || !aligned16(a, r, out)) {
    set_error("bcast_add: null or unaligned pointer");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const i64 hwc8 = hw * (c / 8);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    tr::bcast_add_kernel<<<tr::grid_for(n_img * hwc8), 256, 0, st>>>(
        static_cast<const uint4 *>(a), static_cast<const uint4 *>(r), static_cast<uint4 *>(out),
        n_img, spp, hwc8, scale);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

extern "C" int sbmc_maxpool2x2_bwd_nhwc_bf16(const void *x, const void *dpool, const void *dskip,
                                             int64_t skip_pitch, void *out, int64_t n, int h, int w,
                                             int c, int act, void *stream) {
  if (n < 0 || h < 1 || w < 1 || c < 8 || c % 8 || act < 0 || act > 2 ||
      (dskip && (skip_pitch < c || skip_pitch % 8))) {
    set_error("maxpool2x2_bwd: invalid shape");
    return SBMC_EINVAL;
  }
  if (n == 0) return SBMC_OK;
  if (!x || !dpool || !out || !aligned16(x, dpool, dskip, out)) {
    set_error("maxpool2x2_bwd: null or unaligned pointer");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int hb = (h + 1) / 2, wb = (w + 1) / 2;
  int by = (wb * (c / 8) + 255) / 256;
  if (by > 8) by = 8;
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    tr::maxpool2x2_bwd_kernel<<<dim3((unsigned)(n * hb), (unsigned)by), 256, 0, st>>>(
        static_cast<const uint4 *>(x), static_cast<const uint4 *>(dpool),
        static_cast<const uint4 *>(dskip), skip_pitch / 8, static_cast<uint4 *>(out), h, w, c / 8,
        act);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

extern "C" int sbmc_upsample_bwd_nhwc_bf16(const void *dup, int64_t pitch, const void *coarse,
                                           void *out, int64_t n, int hl, int wl, int h, int w, int c,
                                           int act, void *stream) {
  if (n < 0 || hl < 1 || wl < 1 || h < 1 || w < 1 || c < 8 || c % 8 || pitch < c || pitch % 8 ||
      act < 0 || act > 2) {
    set_error("upsample_bwd: invalid shape");
    return SBMC_EINVAL;
  }
  if (n == 0) return SBMC_OK;
  if (!dup || !out || (act && !coarse) || !aligned16(dup, coarse, out)) {
    set_error("upsample_bwd: null or unaligned pointer");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sy = (float)hl / (float)h, sx = (float)wl / (float)w;
  // candidate window: a coarse pixel is touched by the fine pixels of about 2 / scale rows
  const int ry = (int)(2.5f / sy) + 4, rx = (int)(2.5f / sx) + 4;
  int by = (wl * (c / 8) + 255) / 256;
  if (by > 8) by = 8;
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    tr::upsample_bwd_kernel<<<dim3((unsigned)(n * hl), (unsigned)by), 256, 0, st>>>(
        static_cast<const uint4 *>(dup), pitch / 8, static_cast<const uint4 *>(coarse),
        static_cast<uint4 *>(out), hl, wl, h, w, c / 8, sy, sx, ry, rx, act);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

extern "C" int sbmc_dact_bf16(const void *y, const void *g, void *out, int64_t elems, int act,
                              void *stream) {
  if (elems < 0 || elems % 8 || act < 0 || act > 2) {
    set_error("dact: invalid shape");
    return SBMC_EINVAL;
  }
  if (elems == 0) return SBMC_OK;
  if (!y || !g || !out || !aligned16(y, g, out)) {
    set_error("dact: null or unaligned pointer");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    tr::dact_kernel<<<tr::grid_for(elems / 8), 256, 0, st>>>(
        static_cast<const uint4 *>(y), static_cast<const uint4 *>(g), static_cast<uint4 *>(out),
        elems / 8, act);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

extern "C" int sbmc_colsum_bf16(const void *x, int64_t pitch, int64_t rows, int c, float *workspace,
                                int nblk, float *out, void *stream) {
  if (rows < 1 || c < 8 || c % 8 || c > 2048 || pitch < c || pitch % 8 || nblk < 1) {
    set_error("colsum: invalid shape");
    return SBMC_EINVAL;
  }
  if (!x || !workspace || !out || !aligned16(x)) {
    set_error("colsum: null or unaligned pointer");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int c8 = c / 8;
  if (c8 > 256) {
    set_error("colsum: at most 2048 channels");
    return SBMC_EUNSUPPORTED;
  }
  const int lanes = 256 / c8;
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    tr::colsum_partial_kernel<<<(unsigned)nblk, 256, (size_t)lanes * c * sizeof(float), st>>>(
        static_cast<const uint4 *>(x), pitch / 8, rows, c8, workspace);
    tr::colsum_final_kernel<<<(c + 31) / 32, 256, 0, st>>>(workspace, nblk, c, out);
  }
  count_launch(2);
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}
