// unet_ops.cu -- glue of the U-net (sbmc/modules.py:307-320) as one HBM pass.
//
// The decoder of every level does
//     us = F.interpolate(next_level, size=left.shape[-2:], mode="bilinear", align_corners=False)
//     concat = th.cat([us, left], 1)
// i.e. two kernels and an intermediate tensor.  On channels-innermost bf16
// activations (the inference pipeline's layout) this is one elementwise pass:
// every thread produces 8 channels (16 bytes) of one output pixel, either
// interpolated from the 4 neighbouring coarse pixels or copied from the skip
// tensor.  Interpolation weights follow PyTorch's align_corners=False rule
// (src = scale * (dst + 0.5) - 0.5, clamped at 0; fp32 math, bf16 storage).
#include <cuda_bf16.h>

#include "common.cuh"

namespace sbmc {

__device__ __forceinline__ void unpack8(const uint4 &q, float (&f)[8]) {
  const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// One block row (blockIdx.y) per output image row: all index arithmetic is 32-bit
// with one division per 16-byte chunk (the first version decomposed a 64-bit flat
// index with five 64-bit divisions per chunk and was instruction-bound).
__global__ void __launch_bounds__(256)
upsample_concat_kernel(const uint4 *__restrict__ low, const uint4 *__restrict__ skip,
                       uint4 *__restrict__ out, int hl, int wl, int h, int w, int cu8,
                       int cs8, float sy_scale, float sx_scale) {
  const int ct8 = cu8 + cs8;
  const int y = blockIdx.x % h;
  const i64 img = blockIdx.x / h;
  float sy = sy_scale * (y + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  const int y0 = (int)sy;
  const int y1 = y0 + (y0 < hl - 1 ? 1 : 0);
  const float ly = sy - y0, hy = 1.f - ly;
  const uint4 *row0 = low + (img * hl + y0) * (i64)wl * cu8;
  const uint4 *row1 = low + (img * hl + y1) * (i64)wl * cu8;
  const uint4 *srow = skip + (img * h + y) * (i64)w * cs8;
  uint4 *orow = out + (img * h + y) * (i64)w * ct8;
  const int total = w * ct8;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += gridDim.y * blockDim.x) {
    const int x = idx / ct8;
    const int c8 = idx - x * ct8;
    if (c8 >= cu8) {  // skip connection: straight copy
      orow[idx] = __ldg(srow + x * cs8 + (c8 - cu8));
      continue;
    }
    float sx = sx_scale * (x + 0.5f) - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    const int x0 = (int)sx;
    const int x1 = x0 + (x0 < wl - 1 ? 1 : 0);
    const float lx = sx - x0, hx = 1.f - lx;
    float a[8], b[8], c[8], d[8];
    unpack8(__ldg(row0 + x0 * cu8 + c8), a);
    unpack8(__ldg(row0 + x1 * cu8 + c8), b);
    unpack8(__ldg(row1 + x0 * cu8 + c8), c);
    unpack8(__ldg(row1 + x1 * cu8 + c8), d);
    uint4 q;
    __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float v0 = hy * (hx * a[2 * i] + lx * b[2 * i]) + ly * (hx * c[2 * i] + lx * d[2 * i]);
      const float v1 = hy * (hx * a[2 * i + 1] + lx * b[2 * i + 1]) +
                       ly * (hx * c[2 * i + 1] + lx * d[2 * i + 1]);
      o[i] = __floats2bfloat162_rn(v0, v1);
    }
    orow[idx] = q;
  }
}

// 2 x 2 max pooling, stride 2 (the U-net's `downsample`, sbmc/modules.py:296-299:
// nn.MaxPool2d(2, 2), floor mode) on bf16 channels-innermost tensors: one thread per
// 16-byte chunk (8 channels) of an output pixel.
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, int h, int w, int c8) {
  const int ho = h / 2, wo = w / 2;
  const int yo = blockIdx.x % ho;
  const i64 img = blockIdx.x / ho;
  const uint4 *r0 = x + (img * h + 2 * yo) * (i64)w * c8;
  const uint4 *r1 = r0 + (i64)w * c8;
  uint4 *orow = y + (img * ho + yo) * (i64)wo * c8;
  const int total = wo * c8;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += gridDim.y * blockDim.x) {
    const int xo = idx / c8;
    const int c = idx - xo * c8;
    const uint4 a = __ldg(r0 + (2 * xo) * c8 + c), b = __ldg(r0 + (2 * xo + 1) * c8 + c);
    const uint4 e = __ldg(r1 + (2 * xo) * c8 + c), f = __ldg(r1 + (2 * xo + 1) * c8 + c);
    uint4 q;
    const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&a);
    const __nv_bfloat162 *pb = reinterpret_cast<const __nv_bfloat162 *>(&b);
    const __nv_bfloat162 *pe = reinterpret_cast<const __nv_bfloat162 *>(&e);
    const __nv_bfloat162 *pf = reinterpret_cast<const __nv_bfloat162 *>(&f);
    __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __hmax2(__hmax2(pa[i], pb[i]), __hmax2(pe[i], pf[i]));
    orow[idx] = q;
  }
}

}  // namespace sbmc

extern "C" int sbmc_upsample_concat_nhwc_bf16(const void *low, const void *skip, void *out,
                                              int64_t n, int hl, int wl, int h, int w, int cu,
                                              int cs, void *stream) {
  using namespace sbmc;
  if (n < 0 || hl < 1 || wl < 1 || h < 1 || w < 1 || cu < 0 || cs < 0 || cu + cs < 1) {
    set_error("upsample_concat: invalid shape");
    return SBMC_EINVAL;
  }
  if (n == 0) return SBMC_OK;
  if (!out || (cu > 0 && !low) || (cs > 0 && !skip)) {
    set_error("upsample_concat: null pointer argument");
    return SBMC_EINVAL;
  }
  const void *ptrs[] = {low, skip, out};
  for (int i = 0; i < 3; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) {
      set_error("upsample_concat: pointers must be 16-byte aligned");
      return SBMC_EALIGN;
    }
  if (cu % 8 || cs % 8) {
    set_error("upsample_concat: channel counts must be multiples of 8 (got %d, %d)", cu, cs);
    return SBMC_EUNSUPPORTED;
  }
  if (n * h >= (1ll << 31) || (i64)w * ((cu + cs) / 8) >= (1ll << 31)) {
    set_error("upsample_concat: image too large");
    return SBMC_EUNSUPPORTED;
  }
  const int per_row = w * ((cu + cs) / 8);
  int bx = (per_row + 255) / 256;
  if (bx > 8) bx = 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    const dim3 grid((unsigned)(n * h), (unsigned)bx);     // x: image rows, y: chunks of a row
    upsample_concat_kernel<<<grid, 256, 0, st>>>(
        static_cast<const uint4 *>(low), static_cast<const uint4 *>(skip),
        static_cast<uint4 *>(out), hl, wl, h, w, cu / 8, cs / 8, (float)hl / (float)h,
        (float)wl / (float)w);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}

// ---------------------------------------------------------------------------
// y = act(y + bias[c]) in place on a bf16 channels-innermost tensor [pixels][C]:
// the bias add and the activation that follow every U-net convolution
// (sbmc/modules.py:176-181), which eager PyTorch runs as a broadcast add (a
// non-vectorised kernel) plus an activation kernel.  act: 0 none, 1 ReLU,
// 2 LeakyReLU(0.01).  Each thread handles 8 channels (16 bytes).
// ---------------------------------------------------------------------------
namespace sbmc {

__global__ void __launch_bounds__(256)
bias_act_kernel(uint4 *__restrict__ y, const float *__restrict__ bias, i64 total8, int c8,
                int act) {
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total8;
       idx += (i64)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8) * 8;
    float v[8];
    unpack8(y[idx], v);
    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + c + 4));
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint4 q;
    __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a0 = v[2 * i] + b[2 * i], a1 = v[2 * i + 1] + b[2 * i + 1];
      if (act == 1) {
        a0 = fmaxf(a0, 0.f);
        a1 = fmaxf(a1, 0.f);
      } else if (act == 2) {
        a0 = a0 > 0.f ? a0 : 0.01f * a0;
        a1 = a1 > 0.f ? a1 : 0.01f * a1;
      }
      o[i] = __floats2bfloat162_rn(a0, a1);
    }
    y[idx] = q;
  }
}

}  // namespace sbmc

extern "C" int sbmc_bias_act_nhwc_bf16(void *y, const float *bias, int64_t pixels, int c, int act,
                                       void *stream) {
  using namespace sbmc;
  if (pixels < 0 || c < 1 || act < 0 || act > 2) {
    set_error("bias_act: invalid arguments");
    return SBMC_EINVAL;
  }
  if (pixels == 0) return SBMC_OK;
  if (!y || !bias) {
    set_error("bias_act: null pointer argument");
    return SBMC_EINVAL;
  }
  if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15) || c % 8) {
    set_error("bias_act: needs 16-byte aligned pointers and a channel count multiple of 8");
    return SBMC_EUNSUPPORTED;
  }
  const i64 total8 = pixels * (c / 8);
  i64 blocks = ceil_div(total8, 256);
  const i64 cap = (i64)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    bias_act_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<uint4 *>(y), bias, total8, c / 8,
                                                      act);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}

// ---------------------------------------------------------------------------
// fp32 [n][c][hw] (channel planes) -> bf16 [n][hw][cpad] (channels innermost,
// zero-padded to cpad): the entry of the inference pipeline for the raw sample
// features (sbmc/models.py:120-129).  One thread per pixel: per channel a warp
// reads 32 consecutive floats of one plane (coalesced), and every thread writes
// its pixel's cpad channels as contiguous 16-byte chunks.
// ---------------------------------------------------------------------------
namespace sbmc {

// Block = 64 pixels x all channels through shared memory: per channel pair a warp
// reads 2 x 32 consecutive floats of two planes (coalesced) and writes one packed
// bf16x2 word per pixel into a padded [64][cpad / 2 + 1] tile (conflict-free); the
// tile is then stored as 64 x cpad x 2 contiguous bytes, 16 bytes per lane.
// PX pixels per tile: 128, or 64 for wide outputs (cpad > 256: a 128-pixel tile of 512
// channels is 132 KB of shared memory, one CTA per SM; 64 pixels keep three resident).
template <int kT2Px>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_bf16_kernel(const float *__restrict__ x, uint4 *__restrict__ y, i64 n, int c,
                         i64 hw, i64 x_img, i64 y_img8, int cpad8) {
  extern __shared__ uint32_t tile[];               // [kT2Px][cpad / 2 + 1]
  const int cw = cpad8 * 4;                        // bf16x2 words per pixel
  const int pitch = cw + 1;
  const i64 tiles = (hw + kT2Px - 1) / kT2Px;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cw_real = (c + 1) / 2;                 // channel pairs that hold data
  for (i64 t = blockIdx.x; t < n * tiles; t += gridDim.x) {
    const i64 img = t / tiles;
    const i64 p0 = (t - img * tiles) * kT2Px;
    const float *src = x + img * x_img + p0;
    const bool full = p0 + kT2Px <= hw;
    // two channel pairs per iteration: 16 independent 128-byte row reads in flight per warp
    for (int cp = 2 * warp; cp < cw; cp += 16) {
      float a[2][kT2Px / 32], b[2][kT2Px / 32];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c0 = 2 * (cp + u);
#pragma unroll
        for (int h4 = 0; h4 < kT2Px / 32; ++h4) {
          const int px = h4 * 32 + lane;
          const bool ok = (cp + u) < cw_real && (full || p0 + px < hw);
          a[u][h4] = (ok && c0 < c) ? __ldg(src + (i64)c0 * hw + px) : 0.f;
          b[u][h4] = (ok && c0 + 1 < c) ? __ldg(src + (i64)(c0 + 1) * hw + px) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int h4 = 0; h4 < kT2Px / 32; ++h4) {
          if (cp + u < cw) {
            const __nv_bfloat162 v = __floats2bfloat162_rn(a[u][h4], b[u][h4]);
            tile[(h4 * 32 + lane) * pitch + cp + u] = *reinterpret_cast<const uint32_t *>(&v);
          }
        }
    }
    __syncthreads();
    uint4 *dst = y + img * y_img8 + p0 * cpad8;
    const int nchunks = kT2Px * cpad8;
    for (int i = threadIdx.x; i < nchunks; i += 256) {
      const int px = i / cpad8, c8 = i - px * cpad8;
      if (full || p0 + px < hw) {
        const uint32_t *r = tile + px * pitch + c8 * 4;
        dst[i] = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
    __syncthreads();
  }
}

}  // namespace sbmc

extern "C" int sbmc_nchw_to_nhwc_bf16(const float *x, int64_t x_img_stride, void *y,
                                      int64_t y_img_stride, int64_t n, int c, int64_t hw,
                                      int cpad, void *stream) {
  using namespace sbmc;
  if (n < 0 || c < 1 || hw < 0 || cpad < c || cpad % 8) {
    set_error("nchw_to_nhwc: invalid arguments");
    return SBMC_EINVAL;
  }
  if (n == 0 || hw == 0) return SBMC_OK;
  if (!x || !y || (reinterpret_cast<uintptr_t>(y) & 15) || y_img_stride % 8) {
    set_error("nchw_to_nhwc: null or misaligned pointer");
    return SBMC_EINVAL;
  }
  const int tpx = cpad > 256 ? 64 : 128;
  const i64 tiles = n * ((hw + tpx - 1) / tpx);
  i64 blocks = tiles < (i64)num_sms() * 8 ? tiles : (i64)num_sms() * 8;
  const size_t smem = (size_t)tpx * (cpad / 2 + 1) * sizeof(uint32_t);
  if (smem > 96 * 1024) {
    set_error("nchw_to_nhwc: cpad %d too large", cpad);
    return SBMC_EUNSUPPORTED;
  }
  SBMC_CUDA_OK(cudaFuncSetAttribute(nchw_to_nhwc_bf16_kernel<128>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  SBMC_CUDA_OK(cudaFuncSetAttribute(nchw_to_nhwc_bf16_kernel<64>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    if (tpx == 64)
      nchw_to_nhwc_bf16_kernel<64><<<(unsigned)blocks, 256, smem, st>>>(
          x, static_cast<uint4 *>(y), n, c, hw, x_img_stride, y_img_stride / 8, cpad / 8);
    else
      nchw_to_nhwc_bf16_kernel<128><<<(unsigned)blocks, 256, smem, st>>>(
          x, static_cast<uint4 *>(y), n, c, hw, x_img_stride, y_img_stride / 8, cpad / 8);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}

// 2 x 2 / stride 2 max pooling on bf16 channels-innermost x [n][h][w][c] ->
// y [n][h/2][w/2][c] (floor mode, like nn.MaxPool2d(2, 2)); c multiple of 8.
extern "C" int sbmc_maxpool2x2_nhwc_bf16(const void *x, void *y, int64_t n, int h, int w, int c,
                                         void *stream) {
  using namespace sbmc;
  if (n < 0 || h < 2 || w < 2 || c < 8 || c % 8) {
    set_error("maxpool2x2: invalid shape (c must be a multiple of 8, h, w >= 2)");
    return SBMC_EINVAL;
  }
  if (n == 0) return SBMC_OK;
  if (!x || !y || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15)) {
    set_error("maxpool2x2: null or misaligned pointer");
    return SBMC_EINVAL;
  }
  const int ho = h / 2, wo = w / 2;
  if (n * ho >= (1ll << 31) || (i64)w * (c / 8) >= (1ll << 31)) {
    set_error("maxpool2x2: image too large");
    return SBMC_EUNSUPPORTED;
  }
  int bx = (wo * (c / 8) + 255) / 256;
  if (bx > 8) bx = 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    const dim3 grid((unsigned)(n * ho), (unsigned)bx);
    maxpool2x2_kernel<<<grid, 256, 0, st>>>(static_cast<const uint4 *>(x),
                                            static_cast<uint4 *>(y), h, w, c / 8);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}
