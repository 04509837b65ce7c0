// generic.cu -- shape-generic CUDA kernels (any C, KH, KW, W; 64-bit indexing).
//
// Used when the tuned kernels' preconditions do not hold (W % 4 != 0, unusual
// kernel sizes / channel counts, unaligned pointers) and by the parity tests as
// a second, independent device implementation.  Straight per-element
// restatements of the reference formulas (src/kernel_weighting.cpp:45-57,
// 91-117; src/scatter2gather.cpp:37-47) with the reference's accumulation order.
#include "common.cuh"

namespace sbmc {

static constexpr int kThreads = 256;

static inline unsigned grid_for(i64 total) {
  i64 blocks = ceil_div(total, kThreads);
  const i64 cap = (i64)num_sms() * 32;  // grid-stride beyond this
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

__global__ void __launch_bounds__(kThreads)
generic_fwd_kernel(const float *__restrict__ D, const float *__restrict__ Wt,
                   float *__restrict__ out, float *__restrict__ sum_w, i64 N,
                   int C, i64 H, i64 W, int KH, int KW, int halo_top, i64 Hext) {
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  const i64 plane = H * W, total = N * plane;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (i64)gridDim.x * blockDim.x) {
    const i64 x = idx % W, y = (idx / W) % H, n = idx / plane;
    const float *wn = Wt + n * KH * KW * plane + y * W + x;
    for (int cb = 0; cb < C; cb += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      float sw = 0.f;
      for (int dy = 0; dy < KH; ++dy) {
        const i64 ye = y + dy - c0h + halo_top;
        const bool rin = ye >= 0 && ye < Hext;
        for (int dx = 0; dx < KW; ++dx) {
          const float wv = wn[((i64)dy * KW + dx) * plane];
          sw += wv;
          const i64 xx = x + dx - c0w;
          const bool in = rin && xx >= 0 && xx < W;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (cb + k < C) {
              const float d = in ? D[((n * C + cb + k) * Hext + ye) * W + xx] : 0.f;
              acc[k] = fmaf(wv, d, acc[k]);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (cb + k < C) out[(n * C + cb + k) * plane + y * W + x] = acc[k];
      if (cb == 0) sum_w[idx] = sw;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
generic_dweights_kernel(const float *__restrict__ D, const float *__restrict__ dO,
                        const float *__restrict__ dSw, float *__restrict__ dW,
                        i64 N, int C, i64 H, i64 W, int KH, int KW, int halo_top,
                        i64 Hext) {
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  const i64 plane = H * W, taps = (i64)KH * KW, total = N * taps * plane;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (i64)gridDim.x * blockDim.x) {
    const i64 x = idx % W, y = (idx / W) % H;
    const i64 tap = (idx / plane) % taps, n = idx / (plane * taps);
    const int dy = (int)(tap / KW), dx = (int)(tap % KW);
    const i64 ye = y + dy - c0h + halo_top, xx = x + dx - c0w;
    const bool in = ye >= 0 && ye < Hext && xx >= 0 && xx < W;
    float v = dSw[n * plane + y * W + x];
    for (int c = 0; c < C; ++c) {
      const float d = in ? D[((n * C + c) * Hext + ye) * W + xx] : 0.f;
      v = fmaf(d, dO[(n * C + c) * plane + y * W + x], v);
    }
    dW[idx] = v;
  }
}

__global__ void __launch_bounds__(kThreads)
generic_ddata_kernel(const float *__restrict__ Wt, const float *__restrict__ dO,
                     float *__restrict__ dD, i64 N, int C, i64 H, i64 W, int KH,
                     int KW, int halo_top, i64 Hext) {
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  const i64 plane = H * W, total = N * C * Hext * W;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (i64)gridDim.x * blockDim.x) {
    const i64 x = idx % W, qe = (idx / W) % Hext;
    const i64 c = (idx / (W * Hext)) % C, n = idx / (W * Hext * C);
    const i64 q = qe - halo_top;
    float acc = 0.f;
    for (int ry = 0; ry < KH; ++ry) {
      const i64 py = q + ry - c0h;
      if (py < 0 || py >= H) continue;
      for (int rx = 0; rx < KW; ++rx) {
        const i64 px = x + rx - c0w;
        if (px < 0 || px >= W) continue;
        const float wv =
            Wt[((n * KH + (KH - 1 - ry)) * KW + (KW - 1 - rx)) * plane + py * W + px];
        acc = fmaf(wv, dO[(n * C + c) * plane + py * W + px], acc);
      }
    }
    dD[idx] = acc;
  }
}

__global__ void __launch_bounds__(kThreads)
generic_s2g_kernel(const float *__restrict__ S, float *__restrict__ G, i64 N,
                   int KH, int KW, i64 H, i64 W) {
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  const i64 plane = H * W, taps = (i64)KH * KW, total = N * taps * plane;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (i64)gridDim.x * blockDim.x) {
    const i64 x = idx % W, y = (idx / W) % H;
    const i64 tap = (idx / plane) % taps, n = idx / (plane * taps);
    const int dy = (int)(tap / KW), dx = (int)(tap % KW);
    const i64 yy = y + dy - c0h, xx = x + dx - c0w;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      v = S[((n * KH + (KH - 1 - dy)) * KW + (KW - 1 - dx)) * plane + yy * W + xx];
    G[idx] = v;
  }
}

int generic_fwd(const float *data_ext, const float *weights, float *output,
                float *sum_w, i64 n, int c, i64 h, i64 w, int kh, int kw,
                int halo_top, int halo_bot, cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  KernelTimer timer(SBMC_KERNEL_KW_FWD, st);
  generic_fwd_kernel<<<grid_for(n * h * w), kThreads, 0, st>>>(
      data_ext, weights, output, sum_w, n, c, h, w, kh, kw, halo_top, hext);
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

int generic_bwd_dweights(const float *data_ext, const float *d_output,
                         const float *d_sum_w, float *d_weights, i64 n, int c,
                         i64 h, i64 w, int kh, int kw, int halo_top,
                         int halo_bot, cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  KernelTimer timer(SBMC_KERNEL_KW_DWEIGHTS, st);
  generic_dweights_kernel<<<grid_for(n * kh * kw * h * w), kThreads, 0, st>>>(
      data_ext, d_output, d_sum_w, d_weights, n, c, h, w, kh, kw, halo_top, hext);
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

int generic_bwd_ddata(const float *weights, const float *d_output,
                      float *d_data_ext, i64 n, int c, i64 h, i64 w, int kh,
                      int kw, int halo_top, int halo_bot, cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  KernelTimer timer(SBMC_KERNEL_KW_DDATA, st);
  generic_ddata_kernel<<<grid_for(n * c * hext * w), kThreads, 0, st>>>(
      weights, d_output, d_data_ext, n, c, h, w, kh, kw, halo_top, hext);
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

int generic_s2g(const float *scatter, float *gather, i64 n, int kh, int kw,
                i64 h, i64 w, cudaStream_t st) {
  KernelTimer timer(SBMC_KERNEL_S2G, st);
  generic_s2g_kernel<<<grid_for(n * kh * kw * h * w), kThreads, 0, st>>>(
      scatter, gather, n, kh, kw, h, w);
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace sbmc
