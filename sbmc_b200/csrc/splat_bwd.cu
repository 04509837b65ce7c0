// splat_bwd.cu -- fused backward of one ProgressiveKernelApply update (splat mode).
//
// Forward (sbmc/modules.py:419-473, splat=True), with G = Scatter2Gather(S):
//   m' = max(max_taps G, m);  a = exp(m - m')
//   sum_r' = a sum_r + sum_taps e r(tap);  sum_w' = a sum_w + sum_taps e;  e = exp(G - m')
// Given the upstream gradients g_r = dL/dsum_r', g_w = dL/dsum_w', g_m = dL/dm',
// the reference's autograd walks KernelWeighting.backward, exp_, sub_, max and
// Scatter2Gather.backward: ~10 passes over K*K-channel tensors.  In SCATTER space
// all of it is local to the source sample p and its tap t (target q = p + off(t)):
//   dS[t,p]  = e (g_w[q] + sum_c g_r[c,q] r[c,p])  +  [S[t,p] == m'[q]] T_k[q]
//   d r[c,p] = sum_t e g_r[c,q]                     with e = exp(S[t,p] - m'[q])
// where T_k[q] is the gradient that reaches the running max through the tap
// maximum (computed per pixel by the caller from image-sized planes, see
// sbmc_b200/splat.py; it is zero where the max came from the previous samples).
// So one kernel reads S once and writes dS once (2 x 4 K^2 bytes per sample);
// out-of-image targets read zero-filled planes and therefore get dS = 0, which
// is what Scatter2Gather.backward does with them.
//
// Layout follows kw_fwd_kernel: a thread owns 4 consecutive source pixels, every
// tap is one aligned 128-bit load of S and one aligned 128-bit store of dS, and
// the per-target planes (g_r[C], g_w, m', T_k packed as one [N, C+3, H, W]
// tensor) sit in shared memory (one TMA box, zero outside the image) and are
// read through a sliding register window.
#include "kw_launch.cuh"

namespace sbmc {

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int C, int KW, int ROWS, int MINB, int CH>
__global__ void __launch_bounds__(ROWS * 32, MINB)
splat_bwd_kernel(const __grid_constant__ CUtensorMap pmap,   // planes [N][C+3][H][W]
                 const float *__restrict__ S, const float *__restrict__ R,
                 float *__restrict__ dS, float *__restrict__ dR, int H, int W, int KH,
                 int xtiles, int ytiles) {
  using G = TileGeom<KW>;
  constexpr int P = C + 3;               // g_r[0..C), g_w, m', T_k
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile = reinterpret_cast<float *>(smem_raw);
  const int trows = ROWS + KH - 1;
  uint64_t *bar = reinterpret_cast<uint64_t *>(
      smem_raw + (((size_t)P * trows * G::TWS * sizeof(float) + 15) & ~(size_t)15));

  const TileCoord tc = decode_tile(blockIdx.x, xtiles, ytiles);
  const int X0 = tc.xt * kTileW, Y0 = tc.yt * ROWS;
  const int sh = KH - 1 - (KH - 1) / 2;  // target row of tap ky: y + ky - sh
  load_image_tile<KW>(&pmap, tile, bar, P, trows, X0, Y0 - sh, tc.n);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = Y0 + warp, x0 = X0 + 4 * lane;
  const bool valid = (y < H) && (x0 < W);
  const i64 plane = (i64)H * W;
  const i64 pix = (i64)y * W + x0;

  float r[C][4], dr[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dr[c][i] = 0.f;
    if (valid) {
      const float4 t = ldg_cached(R + ((i64)tc.n * C + c) * plane + pix);
      r[c][0] = t.x; r[c][1] = t.y; r[c][2] = t.z; r[c][3] = t.w;
    }
  }

  mbar_wait(bar, 0);

  if (valid) {
    const float *sp = S + (i64)tc.n * KH * KW * plane + pix;
    float *dp = dS + (i64)tc.n * KH * KW * plane + pix;
    const float *srow = tile + (size_t)warp * G::TWS + 4 * lane;
    const int cstride = trows * G::TWS;
    for (int ky = 0; ky < KH; ++ky) {
#pragma unroll
      for (int cs = 0; cs < KW; cs += CH) {
        float sv[CH][4];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (cs + j < KW) {
            const float4 t = ldg_stream(sp + (i64)(cs + j) * plane);
            sv[j][0] = t.x; sv[j][1] = t.y; sv[j][2] = t.z; sv[j][3] = t.w;
          }
        }
        constexpr int WMAX = (CH + 3 + 3 + 3) / 4 * 4;
        const int lo = (G::LEFT + cs) & ~3;
        const int last = (cs + CH < KW ? cs + CH : KW) - 1;
        const int hi = G::LEFT + last + 4;
        // e = exp(S - m') and the rare "this tap is the max" term, from m' and T_k
        float e[CH][4], out[CH][4];
        {
          float wm[WMAX];
#pragma unroll
          for (int q = 0; q < WMAX; q += 4) {
            if (lo + q < hi) {
              const float4 t = *reinterpret_cast<const float4 *>(
                  srow + (C + 1) * cstride + lo + q);
              wm[q] = t.x; wm[q + 1] = t.y; wm[q + 2] = t.z; wm[q + 3] = t.w;
            }
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (cs + j < KW) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int wi = G::LEFT + cs + j + i - lo;
                const float d = sv[j][i] - wm[wi];
                // d <= 0 for in-image targets; the clamp keeps out-of-image ones
                // (m' read as 0) from overflowing before they are multiplied by 0
                e[j][i] = ex2_approx_b(fminf(d, 0.f) * 1.4426950408889634f);
                out[j][i] = (d == 0.f) ? srow[(C + 2) * cstride + lo + wi] : 0.f;
              }
            }
          }
        }
        // de = g_w + sum_c g_r[c] r[c];  dS = e de + out;  dr[c] += e g_r[c]
        float de[CH][4];
        {
          float wg[WMAX];
#pragma unroll
          for (int q = 0; q < WMAX; q += 4) {
            if (lo + q < hi) {
              const float4 t = *reinterpret_cast<const float4 *>(srow + C * cstride + lo + q);
              wg[q] = t.x; wg[q + 1] = t.y; wg[q + 2] = t.z; wg[q + 3] = t.w;
            }
          }
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (cs + j < KW) {
#pragma unroll
              for (int i = 0; i < 4; ++i) de[j][i] = wg[G::LEFT + cs + j + i - lo];
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float wr[WMAX];
#pragma unroll
          for (int q = 0; q < WMAX; q += 4) {
            if (lo + q < hi) {
              const float4 t = *reinterpret_cast<const float4 *>(srow + c * cstride + lo + q);
              wr[q] = t.x; wr[q + 1] = t.y; wr[q + 2] = t.z; wr[q + 3] = t.w;
            }
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (cs + j < KW) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float gr = wr[G::LEFT + cs + j + i - lo];
                de[j][i] = fmaf(gr, r[c][i], de[j][i]);
                dr[c][i] = fmaf(e[j][i], gr, dr[c][i]);
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (cs + j < KW) {
            float4 v;
            v.x = fmaf(e[j][0], de[j][0], out[j][0]);
            v.y = fmaf(e[j][1], de[j][1], out[j][1]);
            v.z = fmaf(e[j][2], de[j][2], out[j][2]);
            v.w = fmaf(e[j][3], de[j][3], out[j][3]);
            stg_policy<2>(dp + (i64)(cs + j) * plane, v);
          }
        }
      }
      sp += (i64)KW * plane;
      dp += (i64)KW * plane;
      srow += G::TWS;
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
      *reinterpret_cast<float4 *>(dR + ((i64)tc.n * C + c) * plane + pix) =
          make_float4(dr[c][0], dr[c][1], dr[c][2], dr[c][3]);
  }
}

template <int C, int KW, int ROWS, int MINB, int CH>
int run_splat_bwd(const float *planes, const float *kernels, const float *data,
                  float *d_kernels, float *d_data, i64 n, i64 h, i64 w, int kh,
                  cudaStream_t st) {
  constexpr int P = C + 3;
  CUtensorMap pmap;
  if (!make_image_map<KW>(&pmap, planes, n, P, h, w, ROWS + kh - 1)) return SBMC_ECUDA;
  const size_t smem = tile_smem_bytes<P, KW, ROWS>(kh);
  auto kern = splat_bwd_kernel<C, KW, ROWS, MINB, CH>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int xt = (int)ceil_div(w, kTileW), yt = (int)ceil_div(h, ROWS);
  const unsigned grid = (unsigned)((i64)xt * yt * n);
  {
    KernelTimer timer(SBMC_KERNEL_SPLAT_BWD, st);
    kern<<<grid, ROWS * 32, smem, st>>>(pmap, kernels, data, d_kernels, d_data, (int)h,
                                        (int)w, kh, xt, yt);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

constexpr int kSplatBwdRows = 8, kSplatBwdMinB = 2, kSplatBwdCh = 7;

}  // namespace sbmc

extern "C" int sbmc_progressive_splat_bwd_f32(const float *planes, const float *kernels,
                                              const float *data, float *d_kernels,
                                              float *d_data, int64_t n, int c, int64_t h,
                                              int64_t w, int kh, int kw, void *stream) {
  using namespace sbmc;
  if (n < 0 || h < 0 || w < 0 || c < 1 || kh < 1 || kw < 1) {
    set_error("invalid shape n=%lld c=%d h=%lld w=%lld kh=%d kw=%d", (long long)n, c,
              (long long)h, (long long)w, kh, kw);
    return SBMC_EINVAL;
  }
  if (n == 0 || h == 0 || w == 0) return SBMC_OK;
  const void *ptrs[] = {planes, kernels, data, d_kernels, d_data};
  for (int i = 0; i < 5; ++i)
    if (!ptrs[i]) {
      set_error("null pointer argument (#%d)", i);
      return SBMC_EINVAL;
    }
  bool ok = (w % 4 == 0) && (kw % 2 == 1) && (kh % 2 == 1);
  for (int i = 0; i < 5; ++i) ok = ok && (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define X(CC, KK)                                                                      \
  if (ok && c == CC && kw == KK &&                                                     \
      tile_shape_ok<CC + 3, KK, kSplatBwdRows>(n, h, w, kh, h)) {                      \
    note_path(1);                                                                      \
    return run_splat_bwd<CC, KK, kSplatBwdRows, kSplatBwdMinB, kSplatBwdCh>(           \
        planes, kernels, data, d_kernels, d_data, n, h, w, kh, st);                    \
  }
  X(3, 21) X(3, 5) X(3, 3) X(3, 7)
#undef X
  set_error("progressive_splat_bwd: no fused kernel for c=%d kh=%d kw=%d w=%lld "
            "(use the composed operators)", c, kh, kw, (long long)w);
  return SBMC_EUNSUPPORTED;
}
