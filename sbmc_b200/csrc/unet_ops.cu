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

__global__ void __launch_bounds__(256)
upsample_concat_kernel(const uint4 *__restrict__ low, const uint4 *__restrict__ skip,
                       uint4 *__restrict__ out, i64 n, int hl, int wl, int h, int w, int cu8,
                       int cs8, float sy_scale, float sx_scale) {
  const int ct8 = cu8 + cs8;
  const i64 total = n * h * w * ct8;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (i64)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % ct8);
    const i64 pix = idx / ct8;
    if (c8 >= cu8) {  // skip connection: straight copy
      out[idx] = __ldg(skip + pix * cs8 + (c8 - cu8));
      continue;
    }
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const i64 img = pix / ((i64)w * h);
    float sy = sy_scale * (y + 0.5f) - 0.5f;
    float sx = sx_scale * (x + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < hl - 1 ? 1 : 0), x1 = x0 + (x0 < wl - 1 ? 1 : 0);
    const float ly = sy - y0, lx = sx - x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const uint4 *base = low + img * hl * wl * cu8 + c8;
    float a[8], b[8], c[8], d[8];
    unpack8(__ldg(base + ((i64)y0 * wl + x0) * cu8), a);
    unpack8(__ldg(base + ((i64)y0 * wl + x1) * cu8), b);
    unpack8(__ldg(base + ((i64)y1 * wl + x0) * cu8), c);
    unpack8(__ldg(base + ((i64)y1 * wl + x1) * cu8), d);
    uint4 q;
    __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float v0 = hy * (hx * a[2 * i] + lx * b[2 * i]) + ly * (hx * c[2 * i] + lx * d[2 * i]);
      const float v1 = hy * (hx * a[2 * i + 1] + lx * b[2 * i + 1]) +
                       ly * (hx * c[2 * i + 1] + lx * d[2 * i + 1]);
      o[i] = __floats2bfloat162_rn(v0, v1);
    }
    out[idx] = q;
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
  const i64 total = n * h * w * ((cu + cs) / 8);
  i64 blocks = ceil_div(total, 256);
  const i64 cap = (i64)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    upsample_concat_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        static_cast<const uint4 *>(low), static_cast<const uint4 *>(skip),
        static_cast<uint4 *>(out), n, hl, wl, h, w, cu / 8, cs / 8, (float)hl / (float)h,
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

__global__ void __launch_bounds__(128)
nchw_to_nhwc_bf16_kernel(const float *__restrict__ x, uint4 *__restrict__ y, i64 n, int c,
                         i64 hw, i64 x_img, i64 y_img8, int cpad8) {
  const i64 tiles = (hw + 127) / 128;
  for (i64 t = blockIdx.x; t < n * tiles; t += gridDim.x) {
    const i64 img = t / tiles;
    const i64 p = (t - img * tiles) * 128 + threadIdx.x;
    if (p >= hw) continue;
    const float *src = x + img * x_img + p;
    uint4 *dst = y + img * y_img8 + p * cpad8;
#pragma unroll 2
    for (int c8 = 0; c8 < cpad8; ++c8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = c8 * 8 + j;
        v[j] = ch < c ? __ldg(src + (i64)ch * hw) : 0.f;
      }
      uint4 q;
      __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      dst[c8] = q;
    }
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
  const i64 tiles = n * ((hw + 127) / 128);
  i64 blocks = tiles < (i64)num_sms() * 16 ? tiles : (i64)num_sms() * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    nchw_to_nhwc_bf16_kernel<<<(unsigned)blocks, 128, 0, st>>>(
        x, static_cast<uint4 *>(y), n, c, hw, x_img_stride, y_img_stride / 8, cpad / 8);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}
