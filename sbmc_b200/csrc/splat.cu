// splat.cu -- fused ProgressiveKernelApply forward (one streaming pass).
//
// Reference chain (sbmc/modules.py:419-473), per sample index:
//   G = Scatter2Gather(S)            (skipped in gather mode)
//   kmax = max_taps G ; new_max = max(kmax, max_w) ; scaler = exp(max_w - new_max)
//   sum_r = sum_r * scaler + KW_out(data, exp(G - new_max))
//   sum_w = sum_w * scaler + KW_sum_w(exp(G - new_max)) ; max_w = new_max
// which streams the K*K-channel kernel tensor through HBM about eight times.
//
// Here every thread owns 4 target pixels and walks all K*K taps once with an
// online softmax (running max m, running sums rescaled by exp(m_old - m_new)
// whenever the max grows) -- the running state (sum_r, sum_w, max_w) of the
// previous samples is simply the initial value of that recurrence, so the
// progressive update and the per-sample max / exp / weighting collapse into ONE
// pass over the logits: 4*K*K bytes per sample instead of ~32*K*K.
//
// The gather-space logits G[dy,dx,y,x] = S[KH-1-dy, KW-1-dx, y+dy-c0h, x+dx-c0w]
// are never materialised: a producer warp pulls, for every tap, the box of the
// scatter-kernel plane at the SHIFTED coordinates with TMA straight into a
// shared-memory ring (out-of-bounds elements arrive as 0.0f, which is exactly
// the reference's zero logit for taps whose source pixel is outside the image,
// src/scatter2gather.cpp:34-35).  As in s2g.cu the box start is kept 16-byte
// aligned and the residue of the x shift is applied by the consumer warps when
// they read their 4 pixels from shared memory (lds_shifted4).  The radiance tile (+ K-1 halo, zero
// outside the image as in src/kernel_weighting.cpp:35-36) is one more TMA box.
#include <cfloat>

#include "kw_launch.cuh"

namespace sbmc {

template <int C, int KW, int ROWS, int STAGES, int TPS>
struct SplatSmem {
  using G = TileGeom<KW>;
  static constexpr int kBoxFloats = tap_slot_floats(ROWS);  // slot of one tap
  static constexpr int kStageFloats = TPS * kBoxFloats;
  static size_t tile_floats(int kh) { return (size_t)C * (ROWS + kh - 1) * G::TWS; }
  static size_t bytes(int kh) {
    size_t b = (tile_floats(kh) * 4 + 127) & ~(size_t)127;
    b += (size_t)STAGES * kStageFloats * 4;
    b += (2 * STAGES + 1) * 8;
    return b + 128;  // slack for the manual 128-byte alignment
  }
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp(d) for d <= 0 as 2^(d * log2 e): one multiply + one MUFU.  The rounding of
// the product costs |d| * 6e-8 relative, i.e. only the terms that are already
// negligible (large |d|) lose digits; the sums stay within ~1e-6 of expf.
__device__ __forceinline__ float exp_neg(float d) {
  return ex2_approx(d * 1.4426950408889634f);
}

// first != 0: the running state is initialised by this call (its input content
// is ignored); otherwise it is updated in place.
//
// Pipeline unit ("stage") = up to TPS consecutive dx taps of one dy row.  The
// consumer handles a stage in two phases so that the exponentials of a stage
// are independent of each other: (1) stage maximum -> the running max grows at
// most once per stage and the running sums are rescaled branch-free; (2) the
// TPS x 4 exponentials and their multiply-adds against a register window of the
// radiance row (read with aligned 128-bit shared loads, as in kw_fwd_kernel).
template <int C, int KW, int ROWS, int STAGES, int TPS>
__global__ void __launch_bounds__((ROWS + 1) * 32)
splat_fwd_kernel(const __grid_constant__ CUtensorMap dmap,
                 const __grid_constant__ CUtensorMap kmap,     // box 128 x ROWS
                 const __grid_constant__ CUtensorMap kmap_pad, // box 132 x ROWS
                 float *__restrict__ sum_r, float *__restrict__ sum_w,
                 float *__restrict__ max_w, int H, int W, int KH, int splat,
                 int first, int xtiles, int ytiles) {
  using G = TileGeom<KW>;
  using L = SplatSmem<C, KW, ROWS, STAGES, TPS>;
  constexpr int NCH = (KW + TPS - 1) / TPS;      // stages per dy row
  constexpr int C0W = (KW - 1) / 2;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem_raw = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  const int trows = ROWS + KH - 1;
  float *tile = reinterpret_cast<float *>(smem_raw);
  float *ring = reinterpret_cast<float *>(
      smem_raw + (((size_t)C * trows * G::TWS * 4 + 127) & ~(size_t)127));
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)STAGES * L::kStageFloats);
  uint64_t *empty = full + STAGES;
  uint64_t *tbar = empty + STAGES;

  const TileCoord tc = decode_tile(blockIdx.x, xtiles, ytiles);
  const int X0 = tc.xt * kTileW, Y0 = tc.yt * ROWS;
  const int c0h = (KH - 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstages = KH * NCH;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ROWS);
    }
    mbar_init(tbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == ROWS) {
    // ---------------- producer warp: one lane drives the copy engine ----------
    if (lane == 0) {
      prefetch_tensormap(&dmap);
      prefetch_tensormap(&kmap);
      prefetch_tensormap(&kmap_pad);
      mbar_expect_tx(tbar, (uint32_t)(C * trows * G::TWS * sizeof(float)));
      tma_load_4d(tile, &dmap, tbar, X0 - G::A, Y0 - c0h, 0, tc.n);
      for (int it = 0; it < nstages; ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
        const int dy = it / NCH, cs = (it - dy * NCH) * TPS;
        const int cnt = (KW - cs < TPS) ? (KW - cs) : TPS;
        uint32_t bytes = 0;
        for (int j = 0; j < cnt; ++j) {
          const int r = splat ? ((cs + j - C0W) & 3) : 0;
          bytes += (uint32_t)(ROWS * (r ? kBoxW : kTileW) * sizeof(float));
        }
        mbar_expect_tx(&full[s], bytes);
        for (int j = 0; j < cnt; ++j) {
          const int dx = cs + j;
          int plane = dy * KW + dx, sx = X0, sy = Y0, r = 0;
          if (splat) {  // the transposed tap at the shifted position
            const int sh = dx - C0W;
            plane = (KH - 1 - dy) * KW + (KW - 1 - dx);
            sx = X0 + (sh & ~3);
            sy = Y0 + dy - c0h;
            r = sh & 3;
          }
          tma_load_4d(ring + (size_t)s * L::kStageFloats + (size_t)j * L::kBoxFloats,
                      r ? &kmap_pad : &kmap, &full[s], sx, sy, plane, tc.n);
        }
      }
    }
    return;
  }

  // ---------------- consumer warps: warp r owns row Y0 + r -------------------
  const int y = Y0 + warp, x0 = X0 + 4 * lane;
  const bool valid = (y < H) && (x0 < W);
  const long long plane_sz = (long long)H * W;
  const long long pix = (long long)y * W + x0;

  float m[4], aw[4], ar[C][4];
  if (valid && !first) {
    const float4 tm = ldg_cached(max_w + (long long)tc.n * plane_sz + pix);
    const float4 tw = ldg_cached(sum_w + (long long)tc.n * plane_sz + pix);
    m[0] = tm.x; m[1] = tm.y; m[2] = tm.z; m[3] = tm.w;
    aw[0] = tw.x; aw[1] = tw.y; aw[2] = tw.z; aw[3] = tw.w;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 t = ldg_cached(sum_r + ((long long)tc.n * C + c) * plane_sz + pix);
      ar[c][0] = t.x; ar[c][1] = t.y; ar[c][2] = t.z; ar[c][3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m[i] = -FLT_MAX;
      aw[i] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) ar[c][i] = 0.f;
    }
  }

  mbar_wait(tbar, 0);
  const float *srow = tile + (size_t)warp * G::TWS + 4 * lane;
  const int cstride = trows * G::TWS;

  // The sums are built hierarchically -- taps into a per-row partial (rw, rr),
  // rows into the running totals (aw, ar) -- so that hundreds of near-identical
  // tiny terms (e.g. the zero logits of out-of-image taps at the border) are not
  // added one by one to a large accumulator: the fp32 error stays ~20x below
  // that of the plain sequential sum.
  int it = 0;
  for (int dy = 0; dy < KH; ++dy) {
    float rw[4] = {0.f, 0.f, 0.f, 0.f};
    float rr[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) rr[c][i] = 0.f;
#pragma unroll
    for (int cs = 0; cs < KW; cs += TPS, ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
      const float *slot = ring + (size_t)s * L::kStageFloats;
      // ---- phase 1: the logits of this stage and their maximum ----
      float g[TPS][4];
      float smax[4] = {m[0], m[1], m[2], m[3]};
#pragma unroll
      for (int j = 0; j < TPS; ++j) {
        if (cs + j < KW) {
          const int r = splat ? ((cs + j - C0W) & 3) : 0;
          const float4 gv = lds_shifted4(
              slot + (size_t)j * L::kBoxFloats + (size_t)warp * (r ? kBoxW : kTileW), lane, r);
          g[j][0] = gv.x; g[j][1] = gv.y; g[j][2] = gv.z; g[j][3] = gv.w;
#pragma unroll
          for (int i = 0; i < 4; ++i) smax[i] = fmaxf(smax[i], g[j][i]);
        }
      }
      const bool grew = (smax[0] != m[0]) | (smax[1] != m[1]) | (smax[2] != m[2]) |
                        (smax[3] != m[3]);
      if (__any_sync(0xffffffffu, grew)) {  // rare after the first rows: rescale
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = (smax[i] == m[i]) ? 1.f : exp_neg(m[i] - smax[i]);
          aw[i] *= a;
          rw[i] *= a;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            ar[c][i] *= a;
            rr[c][i] *= a;
          }
          m[i] = smax[i];
        }
      }
      // ---- phase 2: exponentials and multiply-adds ----
      constexpr int WMAX = (TPS + 3 + 3 + 3) / 4 * 4;
      const int lo = (G::LEFT + cs) & ~3;
      const int last = (cs + TPS < KW ? cs + TPS : KW) - 1;
      const int hi = G::LEFT + last + 4;
      float e[TPS][4];
#pragma unroll
      for (int j = 0; j < TPS; ++j) {
        if (cs + j < KW) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            e[j][i] = exp_neg(g[j][i] - m[i]);
            rw[i] += e[j][i];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float win[WMAX];
#pragma unroll
        for (int q = 0; q < WMAX; q += 4) {
          if (lo + q < hi) {
            const float4 t = *reinterpret_cast<const float4 *>(srow + c * cstride + lo + q);
            win[q] = t.x; win[q + 1] = t.y; win[q + 2] = t.z; win[q + 3] = t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < TPS; ++j) {
          if (cs + j < KW) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              rr[c][i] = fmaf(e[j][i], win[G::LEFT + cs + j + i - lo], rr[c][i]);
          }
        }
      }
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s]))
                     : "memory");
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      aw[i] += rw[i];
#pragma unroll
      for (int c = 0; c < C; ++c) ar[c][i] += rr[c][i];
    }
    srow += G::TWS;
  }

  if (valid) {
    *reinterpret_cast<float4 *>(max_w + (long long)tc.n * plane_sz + pix) =
        make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4 *>(sum_w + (long long)tc.n * plane_sz + pix) =
        make_float4(aw[0], aw[1], aw[2], aw[3]);
#pragma unroll
    for (int c = 0; c < C; ++c)
      *reinterpret_cast<float4 *>(sum_r + ((long long)tc.n * C + c) * plane_sz + pix) =
          make_float4(ar[c][0], ar[c][1], ar[c][2], ar[c][3]);
  }
}

// Shape-generic fallback: one thread per target pixel, same recurrence.
__global__ void __launch_bounds__(256)
splat_fwd_generic_kernel(const float *__restrict__ K, const float *__restrict__ D,
                         float *__restrict__ sum_r, float *__restrict__ sum_w,
                         float *__restrict__ max_w, i64 N, int C, i64 H, i64 W, int KH,
                         int KW, int splat, int first) {
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  const i64 plane = H * W, total = N * plane;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (i64)gridDim.x * blockDim.x) {
    const i64 x = idx % W, y = (idx / W) % H, n = idx / plane;
    float m = first ? -FLT_MAX : max_w[idx];
    float aw = first ? 0.f : sum_w[idx];
    for (int cb = 0; cb < C; cb += 4) {
      // channels are processed 4 at a time; the max / weight recurrence is
      // repeated identically for each group (only the first group stores it)
      float mm = m, ww = aw;
      float ar[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ar[k] = (first || cb + k >= C) ? 0.f : sum_r[(n * C + cb + k) * plane + y * W + x];
      for (int dy = 0; dy < KH; ++dy) {
        const i64 yy = y + dy - c0h;
        float rw = 0.f, rr[4] = {0.f, 0.f, 0.f, 0.f};   // per-row partial sums
        for (int dx = 0; dx < KW; ++dx) {
          const i64 xx = x + dx - c0w;
          const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
          float g;
          if (splat)
            g = in ? K[((n * KH + (KH - 1 - dy)) * KW + (KW - 1 - dx)) * plane + yy * W + xx]
                   : 0.f;
          else
            g = K[((n * KH + dy) * KW + dx) * plane + y * W + x];
          if (g > mm) {
            const float a = expf(mm - g);
            ww *= a;
            rw *= a;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ar[k] *= a;
              rr[k] *= a;
            }
            mm = g;
          }
          const float e = expf(g - mm);
          rw += e;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (cb + k < C && in)
              rr[k] = fmaf(e, D[((n * C + cb + k) * H + yy) * W + xx], rr[k]);
        }
        ww += rw;
#pragma unroll
        for (int k = 0; k < 4; ++k) ar[k] += rr[k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (cb + k < C) sum_r[(n * C + cb + k) * plane + y * W + x] = ar[k];
      if (cb + 4 >= C) {
        max_w[idx] = mm;
        sum_w[idx] = ww;
      }
    }
  }
}

template <int C, int KW, int ROWS, int STAGES, int TPS>
static int run_splat(const float *kernels, const float *data, float *sum_r, float *sum_w,
                     float *max_w, i64 n, i64 h, i64 w, int kh, int splat, int first,
                     cudaStream_t st) {
  using L = SplatSmem<C, KW, ROWS, STAGES, TPS>;
  CUtensorMap dmap, kmap, kmap_pad;
  if (!make_image_map<KW>(&dmap, data, n, C, h, w, ROWS + kh - 1)) return SBMC_ECUDA;
  {
    const uint64_t dims[4] = {(uint64_t)w, (uint64_t)h, (uint64_t)(kh * KW), (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)w * 4, (uint64_t)w * h * 4,
                                 (uint64_t)w * h * kh * KW * 4};
    const uint32_t box[4] = {(uint32_t)kTileW, (uint32_t)ROWS, 1u, 1u};
    const uint32_t box_pad[4] = {(uint32_t)kBoxW, (uint32_t)ROWS, 1u, 1u};
    if (!encode_tensor_map_f32(&kmap, kernels, 4, dims, strides, box) ||
        !encode_tensor_map_f32(&kmap_pad, kernels, 4, dims, strides, box_pad))
      return SBMC_ECUDA;
  }
  const size_t smem = L::bytes(kh);
  auto kern = splat_fwd_kernel<C, KW, ROWS, STAGES, TPS>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int xt = (int)ceil_div(w, kTileW), yt = (int)ceil_div(h, ROWS);
  const unsigned grid = (unsigned)((i64)xt * yt * n);
  {
    KernelTimer timer(SBMC_KERNEL_SPLAT_FWD, st);
    kern<<<grid, (ROWS + 1) * 32, smem, st>>>(dmap, kmap, kmap_pad, sum_r, sum_w, max_w, (int)h,
                                              (int)w, kh, splat, first, xt, yt);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

// tuning knobs of the fused splat (see profiles/)
constexpr int kSplatRows = 8, kSplatStages = 4, kSplatTps = 3;

int launch_splat_fwd(const float *kernels, const float *data, float *sum_r, float *sum_w,
                     float *max_w, i64 n, int c, i64 h, i64 w, int kh, int kw, int splat,
                     int first, cudaStream_t st) {
  auto a16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool ok = !force_generic() && w % 4 == 0 && a16(kernels) && a16(data) &&
                  a16(sum_r) && a16(sum_w) && a16(max_w) && w < (1ll << 31) - 512 &&
                  h < (1ll << 31) - 512 &&
                  ceil_div(w, kTileW) * ceil_div(h, kSplatRows) * n < 0x7fffffffll &&
                  (unsigned long long)w * h * kh * kw * 4ull < (1ull << 40);
#define X(CC, KK)                                                                     \
  if (ok && c == CC && kw == KK && kSplatRows + kh - 1 <= 256 &&                       \
      SplatSmem<CC, KK, kSplatRows, kSplatStages, kSplatTps>::bytes(kh) <= 220 * 1024) { \
    note_path(1);                                                                     \
    return run_splat<CC, KK, kSplatRows, kSplatStages, kSplatTps>(                     \
        kernels, data, sum_r, sum_w, max_w, n, h, w, kh, splat, first, st);           \
  }
  X(3, 21) X(3, 5) X(3, 3) X(5, 3) X(3, 7)
#undef X
  note_path(2);
  warn_generic("progressive_splat", c, kh, kw, w);
  i64 blocks = ceil_div(n * h * w, 256);
  const i64 cap = (i64)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  {
    KernelTimer timer(SBMC_KERNEL_SPLAT_FWD, st);
    splat_fwd_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        kernels, data, sum_r, sum_w, max_w, n, c, h, w, kh, kw, splat, first);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace sbmc

extern "C" int sbmc_progressive_splat_fwd_f32(const float *kernels, const float *data,
                                              float *sum_r, float *sum_w, float *max_w,
                                              int64_t n, int c, int64_t h, int64_t w,
                                              int kh, int kw, int splat, int first,
                                              void *stream) {
  if (n < 0 || h < 0 || w < 0 || c < 1 || kh < 1 || kw < 1) {
    sbmc::set_error("invalid shape n=%lld c=%d h=%lld w=%lld kh=%d kw=%d", (long long)n, c,
                    (long long)h, (long long)w, kh, kw);
    return SBMC_EINVAL;
  }
  if (n == 0 || h == 0 || w == 0) return SBMC_OK;
  const void *ptrs[] = {kernels, data, sum_r, sum_w, max_w};
  for (int i = 0; i < 5; ++i) {
    if (!ptrs[i]) {
      sbmc::set_error("null pointer argument (#%d)", i);
      return SBMC_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 3) {
      sbmc::set_error("pointer argument #%d is not 4-byte aligned", i);
      return SBMC_EALIGN;
    }
  }
  return sbmc::launch_splat_fwd(kernels, data, sum_r, sum_w, max_w, n, c, h, w, kh, kw,
                                splat ? 1 : 0, first ? 1 : 0,
                                static_cast<cudaStream_t>(stream));
}
