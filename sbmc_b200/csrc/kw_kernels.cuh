// kw_kernels.cuh -- tuned sm_100a kernels for KernelWeighting forward/backward.
//
// What is computed (reference: src/kernel_weighting.cpp:27-124 of adobe/sbmc,
// torch index order as in sbmc/functions.py:78-86):
//   fwd : out[n,c,y,x] = sum_{dy,dx} Wt[n,dy,dx,y,x] * D[n,c,y+dy-c0h,x+dx-c0w]
//         sum_w[n,y,x] = sum_{dy,dx} Wt[n,dy,dx,y,x]
//   dW  : dW[n,dy,dx,y,x] = dSw[n,y,x] + sum_c D[n,c,y+dy-c0h,x+dx-c0w]*dO[n,c,y,x]
//   dD  : dD[n,c,y,x] = sum_{ry,rx} Wt[n,KH-1-ry,KW-1-rx,y+ry-c0h,x+rx-c0w]
//                                   * dO[n,c,y+ry-c0h,x+rx-c0w]
//
// All three are pure HBM streams of the K*K weight volume (1764 B per sample at
// K=21) against a few bytes of reused image data, so the design is:
//   * one thread owns 4 consecutive pixels of one row; every tap of the weight
//     volume is ONE coalesced 128-bit load/store per thread (512 B per warp),
//     issued with streaming cache hints, a whole dx-chunk in flight at a time;
//   * fwd / dW: the reused image tile (+ K-1 halo) is fetched by a single TMA
//     box load per CTA; TMA's out-of-bounds zero fill is the reference's
//     constant_exterior(data, 0) boundary condition, so the inner loop has no
//     bounds checks.  Each thread reads its sliding window from shared memory
//     with aligned 128-bit loads and keeps it in registers across a dx-chunk;
//   * dD: written as gather-in-y / scatter-in-x so the weight loads stay
//     16-byte aligned: a thread accumulates a (4 + KW-1)-wide register window
//     per channel over all taps and the overlapping windows of neighbouring
//     lanes are merged once at the end with warp shuffles, then across warps
//     through shared memory; only tile seams (image wider than one CTA) use
//     global atomics, with exactly two contributors per address, which keeps
//     the result deterministic.
//   * accumulation order per output is dy-outer / dx-inner with fused
//     multiply-add, i.e. the reference's RDom order (src/kernel_weighting.cpp:45).
#pragma once
#include "common.cuh"

namespace sbmc {

constexpr int kTileW = 128;  // pixels per warp-row: 32 lanes x float4

template <int KW>
struct TileGeom {
  static constexpr int C0W = (KW - 1) / 2;        // centre tap, floor
  static constexpr int A = (C0W + 3) / 4 * 4;     // smem col 0 <-> x = X0 - A
  static constexpr int LEFT = A - C0W;            // window index of (dx=0, px 0)
  static constexpr int WIN = (LEFT + KW - 1 + 4 + 3) / 4 * 4;  // floats/thread
  static constexpr int TWS = kTileW - 4 + WIN;    // smem row pitch (floats)
};

// Elements [4*lane + r, 4*lane + r + 4) of a 16-byte aligned shared-memory row;
// r is warp-uniform.
__device__ __forceinline__ float4 lds_shifted4(const float *row, int lane, int r) {
  const float4 a = *reinterpret_cast<const float4 *>(row + 4 * lane);
  if (r == 0) return a;
  const float4 b = *reinterpret_cast<const float4 *>(row + 4 * lane + 4);
  if (r == 1) return make_float4(a.y, a.z, a.w, b.x);
  if (r == 2) return make_float4(a.z, a.w, b.x, b.y);
  return make_float4(a.w, b.x, b.y, b.z);
}

constexpr int kBoxPad = 4;                       // extra columns of a shifted box
constexpr int kBoxW = kTileW + kBoxPad;          // 132 floats = 528 B
// shared-memory slot of one tap's box: TMA destinations must be 128-byte aligned
__host__ __device__ constexpr int tap_slot_floats(int rows) { return (rows * kBoxW + 31) / 32 * 32; }

struct TileCoord {
  int xt, yt, n;
};
__device__ __forceinline__ TileCoord decode_tile(unsigned b, int xtiles,
                                                 int ytiles) {
  TileCoord t;
  t.xt = b % xtiles;
  unsigned r = b / xtiles;
  t.yt = r % ytiles;
  t.n = r / ytiles;
  return t;
}

// One CTA-wide TMA box load of the image tile (all C channels, ROWS + KH - 1
// rows, TWS columns) with zero fill outside the (extended) image.
template <int KW, bool KEEP = false>
__device__ __forceinline__ void load_image_tile(const CUtensorMap *dmap,
                                                float *tile, uint64_t *bar,
                                                int C, int rows, int x0tile,
                                                int y0ext, int n) {
  using G = TileGeom<KW>;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)(C * rows * G::TWS * sizeof(float)));
    if (KEEP)   // the image tile is re-read by neighbouring CTAs and the next kernel
      tma_load_4d_hint(tile, dmap, bar, x0tile - G::A, y0ext, 0, n, l2_policy_evict_last());
    else
      tma_load_4d(tile, dmap, bar, x0tile - G::A, y0ext, 0, n);
  }
}

// ---------------------------------------------------------------------------
// Forward.  grid = xtiles * ytiles * N CTAs, ROWS warps each; warp r owns row
// Y0 + r, lane l owns pixels X0 + 4l .. X0 + 4l + 3.
// ---------------------------------------------------------------------------
template <int C, int KW, int ROWS, int MINB, int CH>
__global__ void __launch_bounds__(ROWS * 32, MINB)
kw_fwd_kernel(const __grid_constant__ CUtensorMap dmap,
              const float *__restrict__ Wt, float *__restrict__ out,
              float *__restrict__ sum_w, int H, int W, int KH, int halo_top,
              int xtiles, int ytiles) {
  using G = TileGeom<KW>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile = reinterpret_cast<float *>(smem_raw);
  const int trows = ROWS + KH - 1;
  uint64_t *bar = reinterpret_cast<uint64_t *>(
      smem_raw + (((size_t)C * trows * G::TWS * sizeof(float) + 15) & ~(size_t)15));

  const TileCoord tc = decode_tile(blockIdx.x, xtiles, ytiles);
  const int X0 = tc.xt * kTileW, Y0 = tc.yt * ROWS;
  const int c0h = (KH - 1) / 2;
  load_image_tile<KW>(&dmap, tile, bar, C, trows, X0, Y0 + halo_top - c0h, tc.n);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = Y0 + warp, x0 = X0 + 4 * lane;
  const bool valid = (y < H) && (x0 < W);
  const i64 plane = (i64)H * W;

  float acc[C][4];
  float sw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sw[i] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c][i] = 0.f;
  }

  mbar_wait(bar, 0);

  if (valid) {
    const float *wp = Wt + (i64)tc.n * KH * KW * plane + (i64)y * W + x0;
    const float *srow = tile + (size_t)warp * G::TWS + 4 * lane;
    const int cstride = trows * G::TWS;
    for (int dy = 0; dy < KH; ++dy) {
#pragma unroll
      for (int cs = 0; cs < KW; cs += CH) {
        float wv[CH][4];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (cs + j < KW) {
            const float4 t = ldg_stream(wp + (i64)(cs + j) * plane);
            wv[j][0] = t.x; wv[j][1] = t.y; wv[j][2] = t.z; wv[j][3] = t.w;
          }
        }
        constexpr int WMAX = (CH + 3 + 3 + 3) / 4 * 4;
        const int lo = (G::LEFT + cs) & ~3;
        const int last = (cs + CH < KW ? cs + CH : KW) - 1;  // last dx in chunk
        const int hi = G::LEFT + last + 4;                   // exclusive
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float win[WMAX];
#pragma unroll
          for (int q = 0; q < WMAX; q += 4) {
            if (lo + q < hi) {
              const float4 t = *reinterpret_cast<const float4 *>(
                  srow + c * cstride + lo + q);
              win[q] = t.x; win[q + 1] = t.y; win[q + 2] = t.z; win[q + 3] = t.w;
            }
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (cs + j < KW) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                acc[c][i] = fmaf(wv[j][i], win[G::LEFT + cs + j + i - lo], acc[c][i]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (cs + j < KW) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sw[i] += wv[j][i];
          }
        }
      }
      wp += (i64)KW * plane;
      srow += G::TWS;
    }
    const i64 pix = (i64)y * W + x0;
#pragma unroll
    for (int c = 0; c < C; ++c)
      stg_stream(out + ((i64)tc.n * C + c) * plane + pix,
                 make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]));
    stg_stream(sum_w + (i64)tc.n * plane + pix,
               make_float4(sw[0], sw[1], sw[2], sw[3]));
  }
}

// ---------------------------------------------------------------------------
// Backward, d_weights: pure write stream.  Same tiling as the forward.
// ---------------------------------------------------------------------------
template <int C, int KW, int ROWS, int MINB, int CH, int STORE = 0>
__global__ void __launch_bounds__(ROWS * 32, MINB)
kw_bwd_dweights_kernel(const __grid_constant__ CUtensorMap dmap,
                       const float *__restrict__ dO,
                       const float *__restrict__ dSw, float *__restrict__ dW,
                       int H, int W, int KH, int halo_top, int xtiles,
                       int ytiles) {
  using G = TileGeom<KW>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile = reinterpret_cast<float *>(smem_raw);
  const int trows = ROWS + KH - 1;
  uint64_t *bar = reinterpret_cast<uint64_t *>(
      smem_raw + (((size_t)C * trows * G::TWS * sizeof(float) + 15) & ~(size_t)15));

  const TileCoord tc = decode_tile(blockIdx.x, xtiles, ytiles);
  const int X0 = tc.xt * kTileW, Y0 = tc.yt * ROWS;
  const int c0h = (KH - 1) / 2;
  load_image_tile<KW, STORE == 3>(&dmap, tile, bar, C, trows, X0, Y0 + halo_top - c0h, tc.n);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = Y0 + warp, x0 = X0 + 4 * lane;
  const bool valid = (y < H) && (x0 < W);
  const i64 plane = (i64)H * W;
  const i64 pix = (i64)y * W + x0;
  const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();

  float go[C][4];
  float gs[4];
  if (valid) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float *src = dO + ((i64)tc.n * C + c) * plane + pix;
      const float4 t = (STORE == 3) ? ldg_hint(src, keep) : ldg_cached(src);
      go[c][0] = t.x; go[c][1] = t.y; go[c][2] = t.z; go[c][3] = t.w;
    }
    const float *src = dSw + (i64)tc.n * plane + pix;
    const float4 t = (STORE == 3) ? ldg_hint(src, keep) : ldg_cached(src);
    gs[0] = t.x; gs[1] = t.y; gs[2] = t.z; gs[3] = t.w;
  }

  mbar_wait(bar, 0);

  if (valid) {
    float *wp = dW + (i64)tc.n * KH * KW * plane + pix;
    const float *srow = tile + (size_t)warp * G::TWS + 4 * lane;
    const int cstride = trows * G::TWS;
    for (int dy = 0; dy < KH; ++dy) {
#pragma unroll
      for (int cs = 0; cs < KW; cs += CH) {
        constexpr int WMAX = (CH + 3 + 3 + 3) / 4 * 4;
        const int lo = (G::LEFT + cs) & ~3;
        const int last = (cs + CH < KW ? cs + CH : KW) - 1;
        const int hi = G::LEFT + last + 4;
        float win[C][WMAX];
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
          for (int q = 0; q < WMAX; q += 4) {
            if (lo + q < hi) {
              const float4 t = *reinterpret_cast<const float4 *>(
                  srow + c * cstride + lo + q);
              win[c][q] = t.x; win[c][q + 1] = t.y;
              win[c][q + 2] = t.z; win[c][q + 3] = t.w;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (cs + j < KW) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[i] = gs[i];  // d_weights starts from d_sum_w (:113)
#pragma unroll
              for (int c = 0; c < C; ++c)
                v[i] = fmaf(win[c][G::LEFT + cs + j + i - lo], go[c][i], v[i]);
            }
            if (STORE == 3)
              stg_hint(wp + (i64)(cs + j) * plane, make_float4(v[0], v[1], v[2], v[3]), stream);
            else
              stg_policy<STORE>(wp + (i64)(cs + j) * plane,
                                make_float4(v[0], v[1], v[2], v[3]));
          }
        }
      }
      wp += (i64)KW * plane;
      srow += G::TWS;
    }
  }
}

// ---------------------------------------------------------------------------
// Backward, d_data.  grid = xtiles * Hext * N CTAs of NSEG warps; the CTA owns
// target row `qe` (extended coordinates) and x range [X0, X0 + NSEG*128); warp
// s / lane l stream the weights of source pixels p = (py, X0 + 128 s + 4 l ..+3)
// for py = q - dy + (KH-1-c0h), and accumulate W*dO into a register window of
// targets x = px + dx - (KW-1-c0w).
// ---------------------------------------------------------------------------
template <int C, int KW, int NSEG, int MINB, int CH>
__global__ void __launch_bounds__(NSEG * 32, MINB)
kw_bwd_ddata_kernel(const float *__restrict__ Wt, const float *__restrict__ dO,
                    float *__restrict__ dD, int H, int W, int KH, int halo_top,
                    int Hext, int xtiles) {
  constexpr int C0W = (KW - 1) / 2;
  constexpr int SW = KW - 1 - C0W;        // x shift of the flipped taps
  constexpr int NG = (KW + 3 + 3) / 4;    // float4 groups per register window
  constexpr int ROWBUF = NSEG * kTileW + 4 * (NG - 1);
  __shared__ __align__(16) float rowbuf[C][ROWBUF];

  const int xt = blockIdx.x % xtiles;
  const unsigned r = blockIdx.x / xtiles;
  const int qe = r % Hext;
  const int n = r / Hext;
  const int X0 = xt * NSEG * kTileW;
  const int c0h = (KH - 1) / 2;
  const int sh = KH - 1 - c0h;
  const int q = qe - halo_top;  // band-relative target row

  const int seg = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = X0 + seg * kTileW + 4 * lane;
  const bool valid = x0 < W;
  const i64 plane = (i64)H * W;

  float acc[C][4 * NG];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int j = 0; j < 4 * NG; ++j) acc[c][j] = 0.f;

  if (valid) {
    for (int dy = 0; dy < KH; ++dy) {
      const int py = q - dy + sh;
      if (py < 0 || py >= H) continue;  // CTA-uniform: zero exterior
      const i64 pix = (i64)py * W + x0;
      const float *wp = Wt + ((i64)n * KH + dy) * KW * plane + pix;
      float go[C][4];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float4 t = ldg_cached(dO + ((i64)n * C + c) * plane + pix);
        go[c][0] = t.x; go[c][1] = t.y; go[c][2] = t.z; go[c][3] = t.w;
      }
#pragma unroll
      for (int cs = 0; cs < KW; cs += CH) {
        float wv[CH][4];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (cs + j < KW) {
            const float4 t = ldg_stream(wp + (i64)(cs + j) * plane);
            wv[j][0] = t.x; wv[j][1] = t.y; wv[j][2] = t.z; wv[j][3] = t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (cs + j < KW) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int c = 0; c < C; ++c)
                acc[c][cs + j + i] = fmaf(wv[j][i], go[c][i], acc[c][cs + j + i]);
          }
        }
      }
    }
  }

  // ---- merge the overlapping windows: lane t group m covers the same pixels
  // as lane t+m group 0.  After the rotation lane L holds group G = L of this
  // warp (`own`) and, for L < NG-1, group G = 32 + L (`over`), which belongs to
  // the next warp's pixels.
  float own[C][4], over[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      own[c][i] = acc[c][i];
      over[c][i] = 0.f;
    }
#pragma unroll
  for (int m = 1; m < NG; ++m) {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v =
            __shfl_sync(0xffffffffu, acc[c][4 * m + i], (lane - m) & 31);
        if (lane >= m) own[c][i] += v;
        else over[c][i] += v;
      }
  }
  // rowbuf index j <-> x = X0 - SW + j
  const int g = seg * 32 + lane;
#pragma unroll
  for (int c = 0; c < C; ++c)
    *reinterpret_cast<float4 *>(&rowbuf[c][4 * g]) =
        make_float4(own[c][0], own[c][1], own[c][2], own[c][3]);
  __syncthreads();
  if (lane < NG - 1) {
    const int g2 = seg * 32 + 32 + lane;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float4 *p = reinterpret_cast<float4 *>(&rowbuf[c][4 * g2]);
      float4 v = make_float4(over[c][0], over[c][1], over[c][2], over[c][3]);
      if (seg != NSEG - 1) {  // the last warp's overhang has no owner: store
        const float4 o = *p;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *p = v;
    }
  }
  __syncthreads();
  // ---- write the row out.  Seams between x tiles get two contributions.
  constexpr int JEND = NSEG * kTileW + KW - 1;  // beyond: zero padding only
  const int seam_lo_end = (xt > 0) ? (KW - 1) : 0;  // j < this: shared w/ left
  const int seam_hi_beg = (xt < xtiles - 1) ? NSEG * kTileW : JEND;
  for (int j = threadIdx.x; j < JEND; j += NSEG * 32) {
    const int x = X0 - SW + j;
    if (x < 0 || x >= W) continue;
    const bool seam = (j < seam_lo_end) || (j >= seam_hi_beg);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float *dst = dD + (((i64)n * C + c) * Hext + qe) * W + x;
      if (seam) atomicAdd(dst, rowbuf[c][j]);
      else *dst = rowbuf[c][j];
    }
  }
}

}  // namespace sbmc
