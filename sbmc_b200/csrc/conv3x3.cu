// conv3x3.cu -- 3x3 convolution (stride 1, zero padding 1) as an implicit GEMM on
// tcgen05, bf16 channels-innermost activations, fp32 accumulation in TMEM, bias +
// activation fused into the epilogue.
//
// Reference: the U-net of sbmc/modules.py:195-320 (`Autoencoder`: per level a
// `left` and a `right` ConvChain of three 3x3 convolutions, modules.py:278-305) --
// 88 % of the model's FLOPs.
//
// Formulation.  out[y][x][co] = bias[co] + sum_{dy,dx,ci} W[co][ci][dy][dx] *
// in[y+dy-1][x+dx-1][ci].  For one of the nine taps this is a GEMM whose A operand
// is the input shifted by (dy-1, dx-1).  A CTA owns an output tile of kRows = 2
// image rows x 128 pixels (two M = 128 MMA row blocks) and kNT output channels:
//
//   * A.  For every 64-channel slab of the input ONE TMA box brings the tile plus
//     its one-pixel halo -- 4 rows x 130 pixels x 64 channels, 128-byte swizzle --
//     into shared memory; TMA's out-of-bounds zero fill IS the convolution's zero
//     padding.  Pixels are the rows of the K-major operand (128 bytes each), so the
//     operand of tap (dy, dx) for output row g is simply the 128 consecutive rows
//     that start at halo row (g + dy) * 130 + dx: the nine taps are nine UMMA
//     descriptors into the SAME resident slab (the swizzle pattern is a function of
//     the shared-memory address, so a descriptor may start at any 128-byte row --
//     checked on hardware by tools/umma_probe2.cu `shift`).  Every input element is
//     fetched from L2 once per slab (x 1.07 halo) instead of nine times.
//   * B.  Weights are prepared as [9][Cout][Cin] bf16; a stage is the
//     (tap, 64-channel slab) block of NT output channels (NT = 128: 16 KB, six-deep
//     mbarrier ring; NT = 256: 32 KB, three-deep), streamed by a second producer
//     thread.  NT = 256 whenever Cout allows: one MMA then covers 128 pixels x 256
//     channels and the A rows are read from shared memory once per 256 outputs --
//     with N = 128 the operand reads alone (A 4 KB + B 4 KB per 64-cycle MMA) use the
//     whole 128 B/clk of shared-memory bandwidth (profiles/r2c_ncu.md).
//   * D.  fp32 accumulators in TMEM: NT = 128 -> two row blocks x 128 columns,
//     double-buffered so that the epilogue of tile i overlaps the main loop of tile
//     i + 1; NT = 256 -> two row blocks x 256 columns = all 512 columns.
//   * Epilogue (16 warps): tcgen05.ld -> + bias -> ReLU / LeakyReLU -> bf16 -> 64-byte
//     vector stores, channels innermost.
//
//   * Narrow images ("linear" mode, W <= 85; the coarse U-net levels of a 128 x 128
//     training crop are 64 and 32 pixels wide, so a 128-pixel tile would be 50-75 %
//     padding).  The halo box is {64 ch, W + 1 px, R rows}: the column past the image edge
//     is zero-filled by TMA, so in shared memory the image is ONE sequence of pixels with
//     a single zero between rows -- which serves both as the right padding of row y and
//     as the left padding of row y + 1.  A tile is then any 256 consecutive elements of
//     that sequence (two M = 128 blocks), tap (dy, dx) is the same descriptor moved by
//     (dy - 1) (W + 1) + (dx - 1) rows, and the epilogue drops the one zero-column
//     element per image row: W / (W + 1) of the MMA rows are useful.
//
// Warp roles: 0 = A producer, 1 = MMA issuer, 2 = TMEM allocation, 3 = B producer,
// 4..19 = epilogue.
#include <cuda_bf16.h>

#include <cstdlib>

#include "umma.cuh"

namespace sbmc {
namespace c3 {

constexpr int kSegPx = 128;
constexpr int kRows = 2;
constexpr int kHaloW = kSegPx + 2;
constexpr int kHaloH = kRows + 2;
constexpr int kASlab = kHaloH * kHaloW * 128;      // 66,560 bytes = 65 KB (1024-aligned)
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (4 + kEpiWarps);
constexpr int kMaxStages = 6;
__host__ __device__ constexpr int stages_for(int nt) { return nt == 128 ? 6 : 3; }

struct Args {
  const float *bias;
  __nv_bfloat16 *out;
  const __nv_bfloat16 *mask;   // may be null: [n][H][W][Cout]; out *= act'(.) read off its sign
  int mask_act;            // 1 ReLU, 2 LeakyReLU(0.01) (data-gradient calls of the training path)
  int act;                 // 0 none, 1 ReLU, 2 LeakyReLU(0.01)
  int H, W, Cin, Cout;
  int tiles_x, tiles_y, n_img, n_tiles_n;
  long long ntiles;
  // "linear" mode for narrow images (pitch > 0, single-CTA kernel): see conv3x3_kernel
  int pitch;               // W + 1: image row + one zero column
  int slab_rows;           // image rows per halo slab
  int tiles_img;           // 256-element tiles per image
};

enum { B_AF = 0, B_AE = 2, B_BF = 4, B_BE = 4 + kMaxStages, B_ACCF = 4 + 2 * kMaxStages,
       B_ACCE = 6 + 2 * kMaxStages, B_COUNT = 8 + 2 * kMaxStages };

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void stg128(void *p, const uint4 &v) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// The CTA-pair kernel (below) serves Cout = 128 by default; sbmc_b200_conv3x3_pair(0) or
// SBMC_B200_CONV_PAIR=0 selects the single-CTA kernel (A/B runs).  History of the
// measurement (profiles/r2j_*, r2o_*): with the MMAs issued from a divergent
// `if (lane == 0)` block both kernels were bound by the issuing thread and the pair was
// 11 % SLOWER; once the issue loop compiled to back-to-back UTCHMMA (umma.cuh::elect_one)
// the pair became 4-8 % FASTER (128->128: 0.195 vs 0.204 ms, 384->128: 0.517 vs 0.560 ms).
static int g_pair = -1;
static bool pair_enabled() {
  if (g_pair < 0) {
    const char *e = getenv("SBMC_B200_CONV_PAIR");
    g_pair = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pair != 0;
}

// Widest image served by the linear mode: its slab, ceil((257 + 3 (W + 1)) / (W + 1)) rows
// of W + 1 pixels, must fit the 520 rows of a halo slab.  SBMC_B200_CONV_LINEAR=0 disables
// the mode (A/B runs).
constexpr int kLinearMaxW = 85;
static int g_linear = -1;
static bool linear_enabled() {
  if (g_linear < 0) {
    const char *e = getenv("SBMC_B200_CONV_LINEAR");
    g_linear = (e && e[0] == '0') ? 0 : 1;
  }
  return g_linear != 0;
}

struct TileCoord { int n, y0, x0, n0; };
__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ TileCoord tile_coord(long long tile, const Args &P, int nt) {
  TileCoord t;
  if (P.pitch > 0) {
    // linear mode: x0 = first element of the tile in the image's [H][W + 1] numbering,
    // y0 = first image row of its halo slab (may be negative: zero fill)
    const int tl = (int)(tile % P.tiles_img); tile /= P.tiles_img;
    t.n = (int)(tile % P.n_img); tile /= P.n_img;
    t.n0 = (int)tile * nt;
    t.x0 = tl * 256;
    t.y0 = floor_div(t.x0 - P.pitch - 1, P.pitch);
    return t;
  }
  const int tx = (int)(tile % P.tiles_x); tile /= P.tiles_x;
  const int ty = (int)(tile % P.tiles_y); tile /= P.tiles_y;
  t.n = (int)(tile % P.n_img); tile /= P.n_img;
  t.n0 = (int)tile * nt;
  t.x0 = tx * kSegPx;
  t.y0 = ty * kRows;
  return t;
}

// NT: output channels per CTA tile = MMA N (128 or 256).
template <int NT>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap amap,      // input  {Cin, W, H, N}
               const __grid_constant__ CUtensorMap wmap,      // weights {Cin, Cout, 9}
               const Args P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                               // 2 slabs
  constexpr int kStages = stages_for(NT);
  constexpr int kBStage = NT * 128;                       // NT output channels x 64 bf16
  unsigned char *sB = smem + 2 * kASlab;                  // kStages stages
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + kStages * kBStage);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + B_COUNT);

  // Role index = hardware warp.  (Experiment, -DSBMC_CTRL_WARPS_LAST: the four control
  // warps as the LAST hardware warps, in case the scheduler's warp-id priority starved the
  // single-thread MMA issuer behind sixteen epilogue warps.  Measured: no gain on
  // conv3x3, embeddings 8 % slower -- profiles/r2n_*.jsonl -- so the default stays.)
#ifdef SBMC_CTRL_WARPS_LAST
  const int tid = threadIdx.x, warp = ((tid >> 5) + 4) % (kThreads / 32), lane = tid & 31;
#else
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#endif
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bars + i, (i >= B_ACCE) ? kEpiWarps : 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nslabs = P.Cin / 64;

  if (warp == 0) {
    // ===================== A producer: halo slabs =====================
    if (lane == 0) {
      uint32_t ph = 0;          // bit ab
      int ab = 0;
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(tile, P, NT);
        for (int s = 0; s < nslabs; ++s) {
          mbar_wait(bars + B_AE + ab, ((ph >> ab) & 1) ^ 1); ph ^= 1u << ab;
          if (P.pitch > 0) {
            mbar_expect_tx(bars + B_AF + ab, (uint32_t)(P.slab_rows * P.pitch * 128));
            tma_load_4d(sA + ab * kASlab, &amap, bars + B_AF + ab, s * 64, 0, t.y0, t.n);
          } else {
            mbar_expect_tx(bars + B_AF + ab, (uint32_t)kASlab);
            tma_load_4d(sA + ab * kASlab, &amap, bars + B_AF + ab, s * 64, t.x0 - 1, t.y0 - 1, t.n);
          }
          ab ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: weight stages =====================
    if (lane == 0) {
      uint32_t ph = 0;          // bit st
      int st = 0;
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const TileCoord t = tile_coord(tile, P, NT);
        for (int s = 0; s < nslabs; ++s)
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bars + B_BE + st, ((ph >> st) & 1) ^ 1); ph ^= 1u << st;
            mbar_expect_tx(bars + B_BF + st, (uint32_t)kBStage);
            tma_load_3d(sB + st * kBStage, &wmap, bars + B_BF + st, s * 64, t.n0, tap);
            st = (st + 1 == kStages) ? 0 : st + 1;
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp: uniform control flow + barrier waits; one elected lane issues
    // (umma.cuh::elect_one); descriptors are hoisted, per MMA only an add remains
    {
      const uint32_t idesc = umma_idesc_bf16(128, NT);
      const uint64_t dA = umma_smem_desc_sw128(sA), dB = umma_smem_desc_sw128(sB);
      uint32_t ph_a = 0, ph_b = 0, ph_acc = 0;
      int ab = 0, st = 0, it = 0;
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
        const int buf = (NT == 128) ? (it & 1) : 0;
        // slab row of tap (0, 0) of row block 0, and the row distance between dy taps /
        // row blocks (linear mode: elements of the [H][W + 1] sequence)
        int base_row = 0, dy_rows = kHaloW, g_rows = kHaloW;
        if (P.pitch > 0) {
          const TileCoord t = tile_coord(tile, P, NT);
          base_row = t.x0 - t.y0 * P.pitch - P.pitch - 1;
          dy_rows = P.pitch;
          g_rows = 128;
        }
        mbar_wait(bars + B_ACCE + buf, ((ph_acc >> buf) & 1) ^ 1); ph_acc ^= 1u << buf;
        for (int s = 0; s < nslabs; ++s) {
          mbar_wait(bars + B_AF + ab, (ph_a >> ab) & 1); ph_a ^= 1u << ab;
          const uint64_t da = dA + (uint64_t)ab * (kASlab >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            constexpr int kRowD = 128 >> 4;               // one halo row (pixel) in 16-byte units
            const int dy = tap / 3, dx = tap - 3 * dy;
            mbar_wait(bars + B_BF + st, (ph_b >> st) & 1); ph_b ^= 1u << st;
            tcgen05_fence_after();
            if (elect_one()) {
              const uint64_t bd0 = dB + (uint64_t)st * (kBStage >> 4);
#pragma unroll
              for (int g = 0; g < kRows; ++g) {
                const uint64_t ad0 =
                    da + (uint64_t)((base_row + g * g_rows + dy * dy_rows + dx) * kRowD);
                const uint32_t d = tmem + ((NT == 128) ? buf * 256 + g * 128 : g * 256);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(d, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc,
                            (s | tap | k) > 0);
              }
              umma_commit(bars + B_BE + st);
              if (tap == 8) {
                umma_commit(bars + B_AE + ab);
                if (s == nslabs - 1) umma_commit(bars + B_ACCF + buf);
              }
            }
            __syncwarp();
            st = (st + 1 == kStages) ? 0 : st + 1;
          }
          ab ^= 1;
        }
      }
      // drain: all commits have landed before the CTA exits
      if (elect_one()) umma_commit(bars + B_AF);
      __syncwarp();
      mbar_wait(bars + B_AF, (ph_a & 1));
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // warp -> (lane quadrant, row block g, column half): 4 x 2 x 2 = 16 warps
    const int quad = warp & 3, g = ((warp - 4) >> 2) & 1, part = (warp - 4) >> 3;
    const int px = quad * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const TileCoord t = tile_coord(tile, P, NT);
      const int buf = (NT == 128) ? (it & 1) : 0;
      mbar_wait(bars + B_ACCF + buf, (ph >> buf) & 1); ph ^= 1u << buf;
      tcgen05_fence_after();
      // 64 channels per step: 2 x tcgen05.ld -> bias, activation -> 32 packed words ->
      // transpose inside the lane quad -> 4 x 256-bit stores that each complete 8 lines
      // (umma.cuh::quad_transpose32; one 16-byte store per lane would touch 32 lines)
      int y = t.y0 + g;
      int xq = t.x0 + px - (lane & 3);                  // first pixel of this lane's quad
      int xown = t.x0 + px;
      if (P.pitch > 0) {
        const int e = t.x0 + g * 128 + px;              // element of the [H][W + 1] sequence
        y = e / P.pitch;
        xown = e - y * P.pitch;                         // == W: the zero column, dropped
        xq = xown - (lane & 3);
      }
      __nv_bfloat16 *dst = P.out + (((long long)t.n * P.H + y) * P.W + xq) * P.Cout + t.n0 +
                           (lane & 3) * 16;
      const float *bias = P.bias + t.n0;
      const __nv_bfloat16 *mrow = nullptr;
      if (P.mask && y < P.H && xown < P.W)
        mrow = P.mask + (((long long)t.n * P.H + y) * P.W + xown) * P.Cout + t.n0;
      const float mslope = (P.mask_act == 2) ? 0.01f : 0.f;
#pragma unroll 1
      for (int c0 = part * (NT / 2); c0 < (part + 1) * (NT / 2); c0 += 64) {
        const uint32_t col = (NT == 128) ? buf * 256 + g * 128 + c0 : g * 256 + c0;
        uint32_t q[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[32];
          tmem_ld_32x32b_x32(lane_base + col + 32 * hh, v);
          if (mrow) apply_act_mask32(v, mrow + c0 + 32 * hh, mslope);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 32 * hh + 4 * q4));
            float u0 = v[4 * q4] + b.x, u1 = v[4 * q4 + 1] + b.y;
            float u2 = v[4 * q4 + 2] + b.z, u3 = v[4 * q4 + 3] + b.w;
            if (P.act == 1) {
              u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); u2 = fmaxf(u2, 0.f); u3 = fmaxf(u3, 0.f);
            } else if (P.act == 2) {
              u0 = fmaxf(u0, 0.01f * u0); u1 = fmaxf(u1, 0.01f * u1);
              u2 = fmaxf(u2, 0.01f * u2); u3 = fmaxf(u3, 0.01f * u3);
            }
            q[16 * hh + 2 * q4] = pack_bf16(u0, u1);
            q[16 * hh + 2 * q4 + 1] = pack_bf16(u2, u3);
          }
        }
        if (NT == 256 && P.pitch == 0) {
          // the epilogue runs after the tile's last MMA: the shuffles have the shared-memory
          // crossbar to themselves
          quad_transpose32(q, lane);
          if (y < P.H) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (xq + k < P.W) stg256(dst + (long long)k * P.Cout + c0, q + 8 * k);
          }
        } else {
          // NT = 128: this epilogue overlaps the next tile's MMAs, whose operand reads
          // already saturate the shared-memory data path that shuffles share (measured:
          // the transposed variant made the 128 -> 128 layer 30 % slower,
          // profiles/r2f_convs.jsonl): plain 256-bit stores of the thread's own row
          if (y < P.H && xown < P.W) {
            __nv_bfloat16 *own = dst - (lane & 3) * 16 + (long long)(lane & 3) * P.Cout + c0;
#pragma unroll
            for (int k = 0; k < 4; ++k) stg256(own + 16 * k, q + 8 * k);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_ACCE + buf);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ===========================================================================
// CTA-pair variant for Cout = 128 (the full-resolution layers, 44 % of the U-net's
// FLOPs).  With N = 128 a single-CTA MMA reads A (4 KB) + B (4 KB) from shared memory
// every 64 cycles -- all of the 128 B/clk the SM has -- and the TMA writes of the next
// operands come on top.  tcgen05.mma.cta_group::2 pairs two SMs: M = 256 = the same row
// block of TWO spatial tiles (one per CTA), each CTA holds only HALF of the weight stage
// (64 of the 128 output channels), and the weight traffic from L2 halves.
//   * both CTAs load their own halo slab and their half of every weight stage; all
//     loads signal the LEADER's mbarriers (cp.async.bulk.tensor ... .cta_group::2);
//   * only the leader issues MMAs; tcgen05.commit ... .multicast::cluster releases the
//     operand slots and publishes the accumulators in both CTAs;
//   * both CTAs run their own epilogue on their own TMEM and tell the leader when an
//     accumulator buffer is free (remote mbarrier arrive).
// ===========================================================================
constexpr int kPairStages = 10;                   // 8 KB half stages
constexpr int kPairBStage = 64 * 128;
enum { P_AF = 0, P_AE = 2, P_BF = 4, P_BE = 4 + kPairStages, P_ACCF = 4 + 2 * kPairStages,
       P_ACCE = 6 + 2 * kPairStages, P_COUNT = 8 + 2 * kPairStages };
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;       // shared::cluster address -> the even CTA's copy

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all MMAs issued so far -> arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
// TMA loads whose completion bytes are credited to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_4d_pair(void *smem_dst, const CUtensorMap *map,
                                                 uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void *smem_dst, const CUtensorMap *map,
                                                 uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive on the leader's copy of a barrier (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                   smem_u32(bar) & kPeerMask)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap amap,      // input  {Cin, W, H, N}
                    const __grid_constant__ CUtensorMap wmap,      // weights {Cin, Cout, 9}, box 64 rows
                    const Args P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                               // 2 halo slabs
  unsigned char *sB = smem + 2 * kASlab;                  // kPairStages half stages
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + kPairStages * kPairBStage);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + P_COUNT);

  // Role index = hardware warp.  (Experiment, -DSBMC_CTRL_WARPS_LAST: the four control
  // warps as the LAST hardware warps, in case the scheduler's warp-id priority starved the
  // single-thread MMA issuer behind sixteen epilogue warps.  Measured: no gain on
  // conv3x3, embeddings 8 % slower -- profiles/r2n_*.jsonl -- so the default stays.)
#ifdef SBMC_CTRL_WARPS_LAST
  const int tid = threadIdx.x, warp = ((tid >> 5) + 4) % (kThreads / 32), lane = tid & 31;
#else
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#endif
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (tid == 0) {
    // full barriers: one arrive.expect_tx by the leader's producer (bytes of both CTAs);
    // empty / acc_full: one multicast commit; acc_empty: the 16 epilogue warps of both CTAs
    for (int i = 0; i < P_COUNT; ++i) mbar_init(bars + i, (i >= P_ACCE) ? 2 * kEpiWarps : 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nslabs = P.Cin / 64;
  const long long npairs_total = (P.ntiles + 1) / 2;
  const long long pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== A producer: this CTA's halo slabs =====================
    if (lane == 0) {
      uint32_t ph = 0;
      int ab = 0;
      for (long long pt = pair0; pt < npairs_total; pt += pair_step) {
        const long long tile = 2 * pt + rank;
        TileCoord t = tile_coord(tile < P.ntiles ? tile : 0, P, 128);
        if (tile >= P.ntiles) t.y0 = P.H + 8;             // dummy tile: the box is all zero fill
        for (int s = 0; s < nslabs; ++s) {
          mbar_wait(bars + P_AE + ab, ((ph >> ab) & 1) ^ 1); ph ^= 1u << ab;
          if (leader) mbar_expect_tx(bars + P_AF + ab, (uint32_t)(2 * kASlab));
          tma_load_4d_pair(sA + ab * kASlab, &amap, bars + P_AF + ab, s * 64, t.x0 - 1, t.y0 - 1, t.n);
          ab ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: this CTA's half of every weight stage =====================
    if (lane == 0) {
      uint32_t ph = 0;
      int st = 0;
      for (long long pt = pair0; pt < npairs_total; pt += pair_step)
        for (int s = 0; s < nslabs; ++s)
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bars + P_BE + st, ((ph >> st) & 1) ^ 1); ph ^= 1u << st;
            if (leader) mbar_expect_tx(bars + P_BF + st, (uint32_t)(2 * kPairBStage));
            tma_load_3d_pair(sB + st * kPairBStage, &wmap, bars + P_BF + st, s * 64, (int)rank * 64, tap);
            st = (st + 1 == kPairStages) ? 0 : st + 1;
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, 128);
      const uint64_t dA = umma_smem_desc_sw128(sA), dB = umma_smem_desc_sw128(sB);
      uint32_t ph_a = 0, ph_b = 0, ph_acc = 0;
      int ab = 0, st = 0, it = 0;
      for (long long pt = pair0; pt < npairs_total; pt += pair_step, ++it) {
        const int buf = it & 1;
        mbar_wait(bars + P_ACCE + buf, ((ph_acc >> buf) & 1) ^ 1); ph_acc ^= 1u << buf;
        for (int s = 0; s < nslabs; ++s) {
          mbar_wait(bars + P_AF + ab, (ph_a >> ab) & 1); ph_a ^= 1u << ab;
          const uint64_t da = dA + (uint64_t)ab * (kASlab >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - 3 * dy;
            mbar_wait(bars + P_BF + st, (ph_b >> st) & 1); ph_b ^= 1u << st;
            tcgen05_fence_after();
            if (elect_one()) {
              const uint64_t bd0 = dB + (uint64_t)st * (kPairBStage >> 4);
#pragma unroll
              for (int g = 0; g < kRows; ++g) {
                const uint64_t ad0 = da + (uint64_t)(((g + dy) * kHaloW + dx) * 8);
                const uint32_t d = tmem + buf * 256 + g * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_2sm(d, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc,
                                (s | tap | k) > 0);
              }
              umma_commit_pair(bars + P_BE + st);
              if (tap == 8) {
                umma_commit_pair(bars + P_AE + ab);
                if (s == nslabs - 1) umma_commit_pair(bars + P_ACCF + buf);
              }
            }
            __syncwarp();
            st = (st + 1 == kPairStages) ? 0 : st + 1;
          }
          ab ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own TMEM) =====================
    const int quad = warp & 3, g = ((warp - 4) >> 2) & 1, part = (warp - 4) >> 3;
    const int px = quad * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph = 0;
    int it = 0;
    for (long long pt = pair0; pt < npairs_total; pt += pair_step, ++it) {
      const long long tile = 2 * pt + rank;
      const bool real = tile < P.ntiles;
      const TileCoord t = tile_coord(real ? tile : 0, P, 128);
      const int buf = it & 1;
      mbar_wait(bars + P_ACCF + buf, (ph >> buf) & 1); ph ^= 1u << buf;
      tcgen05_fence_after();
      const int y = t.y0 + g, x = t.x0 + px;
      __nv_bfloat16 *own = P.out + (((long long)t.n * P.H + y) * P.W + x) * P.Cout;
      const bool valid = real && y < P.H && x < P.W;
      const int c0 = part * 64;
      uint32_t q[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tmem_ld_32x32b_x32(lane_base + buf * 256 + g * 128 + c0 + 32 * hh, v);
        if (P.mask && valid)
          apply_act_mask32(v, P.mask + (own - P.out) + c0 + 32 * hh, (P.mask_act == 2) ? 0.01f : 0.f);
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b = __ldg(reinterpret_cast<const float4 *>(P.bias + c0 + 32 * hh + 4 * q4));
          float u0 = v[4 * q4] + b.x, u1 = v[4 * q4 + 1] + b.y;
          float u2 = v[4 * q4 + 2] + b.z, u3 = v[4 * q4 + 3] + b.w;
          if (P.act == 1) {
            u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); u2 = fmaxf(u2, 0.f); u3 = fmaxf(u3, 0.f);
          } else if (P.act == 2) {
            u0 = fmaxf(u0, 0.01f * u0); u1 = fmaxf(u1, 0.01f * u1);
            u2 = fmaxf(u2, 0.01f * u2); u3 = fmaxf(u3, 0.01f * u3);
          }
          q[16 * hh + 2 * q4] = pack_bf16(u0, u1);
          q[16 * hh + 2 * q4 + 1] = pack_bf16(u2, u3);
        }
      }
      if (valid) {
#pragma unroll
        for (int k = 0; k < 4; ++k) stg256(own + c0 + 16 * k, q + 8 * k);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bars + P_ACCE + buf);
    }
  }
  // nobody leaves while the partner may still read this CTA's shared memory / signal its
  // barriers: the epilogues have seen the last accumulators, i.e. every MMA has retired
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)
                 : "memory");
}

static int launch_pair(const Args &a, const CUtensorMap &am, const CUtensorMap &wm,
                       cudaStream_t st) {
  const size_t smem = (size_t)2 * kASlab + (size_t)kPairStages * kPairBStage +
                      P_COUNT * sizeof(uint64_t) + 16;
  SBMC_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long pairs = (a.ntiles + 1) / 2;
  long long grid = 2 * (pairs < num_sms() / 2 ? pairs : num_sms() / 2);
  {
    KernelTimer timer(SBMC_KERNEL_CONV3X3, st);
    conv3x3_pair_kernel<<<(unsigned)grid, kThreads, smem, st>>>(am, wm, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

template <int NT>
static int launch(const Args &a, const CUtensorMap &am, const CUtensorMap &wm, cudaStream_t st) {
  const size_t smem = (size_t)2 * kASlab + (size_t)stages_for(NT) * NT * 128 +
                      B_COUNT * sizeof(uint64_t) + 16;
  auto kern = conv3x3_kernel<NT>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  {
    KernelTimer timer(SBMC_KERNEL_CONV3X3, st);
    kern<<<(unsigned)grid, kThreads, smem, st>>>(am, wm, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace c3
}  // namespace sbmc

extern "C" int sbmc_b200_conv3x3_linear(int flag) {
  const int prev = sbmc::c3::linear_enabled() ? 1 : 0;
  sbmc::c3::g_linear = flag ? 1 : 0;
  return prev;
}

extern "C" int sbmc_b200_conv3x3_pair(int flag) {
  const int prev = sbmc::c3::pair_enabled() ? 1 : 0;
  sbmc::c3::g_pair = flag ? 1 : 0;
  return prev;
}

extern "C" int sbmc_conv3x3_masked_nhwc_bf16(const void *x, const void *w9, const float *bias,
                                             const void *mask, int mask_act, void *y, int64_t n,
                                             int h, int w, int cin, int cout, int act,
                                             void *stream);

extern "C" int sbmc_conv3x3_nhwc_bf16(const void *x, const void *w9, const float *bias, void *y,
                                      int64_t n, int h, int w, int cin, int cout, int act,
                                      void *stream) {
  return sbmc_conv3x3_masked_nhwc_bf16(x, w9, bias, nullptr, 0, y, n, h, w, cin, cout, act, stream);
}

extern "C" int sbmc_conv3x3_masked_nhwc_bf16(const void *x, const void *w9, const float *bias,
                                             const void *mask, int mask_act, void *y, int64_t n,
                                             int h, int w, int cin, int cout, int act,
                                             void *stream) {
  using namespace sbmc;
  if (n < 0 || h < 1 || w < 1 || cin < 1 || cout < 1 || act < 0 || act > 2 ||
      (mask && (mask_act < 1 || mask_act > 2))) {
    set_error("conv3x3: invalid shape");
    return SBMC_EINVAL;
  }
  if (n == 0) return SBMC_OK;
  if (!x || !w9 || !bias || !y) {
    set_error("conv3x3: null pointer argument");
    return SBMC_EINVAL;
  }
  if (cin % 64 != 0 || cout % 128 != 0 || n >= (1ll << 31)) {
    set_error("conv3x3: needs cin %% 64 == 0 and cout %% 128 == 0 (got %d, %d)", cin, cout);
    return SBMC_EUNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
       reinterpret_cast<uintptr_t>(w9) | reinterpret_cast<uintptr_t>(mask)) & 15) {
    set_error("conv3x3: pointers must be 16-byte aligned");
    return SBMC_EALIGN;
  }
  const int nt = (cout % 256 == 0) ? 256 : 128;
  c3::Args a;
  a.bias = bias; a.out = static_cast<__nv_bfloat16 *>(y); a.act = act;
  a.mask = static_cast<const __nv_bfloat16 *>(mask); a.mask_act = mask_act;
  a.H = h; a.W = w; a.Cin = cin; a.Cout = cout;
  a.tiles_x = (w + c3::kSegPx - 1) / c3::kSegPx;
  a.tiles_y = (h + c3::kRows - 1) / c3::kRows;
  a.n_img = (int)n;
  a.n_tiles_n = cout / nt;
  a.ntiles = (long long)a.tiles_x * a.tiles_y * n * a.n_tiles_n;
  a.pitch = 0; a.slab_rows = 0; a.tiles_img = 0;
  if (w <= c3::kLinearMaxW && c3::linear_enabled()) {
    // narrow image: tiles of 256 consecutive elements of the [H][W + 1] sequence
    a.pitch = w + 1;
    a.slab_rows = (257 + 3 * a.pitch + a.pitch - 1) / a.pitch;
    a.tiles_img = (h * a.pitch + 255) / 256;
    a.ntiles = (long long)a.tiles_img * n * a.n_tiles_n;
  }
  CUtensorMap am, wm;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)cin * 2 * w, (uint64_t)cin * 2 * w * h};
    const uint32_t box[4] = {64, a.pitch ? (uint32_t)a.pitch : c3::kHaloW,
                             a.pitch ? (uint32_t)a.slab_rows : c3::kHaloH, 1};
    if (!encode_tensor_map_bf16_sw128(&am, x, 4, dims, str, box)) return SBMC_ECUDA;
  }
  {
    const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cout, 9};
    const uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)cin * 2 * cout};
    const uint32_t box[3] = {64, (uint32_t)nt, 1};
    if (!encode_tensor_map_bf16_sw128(&wm, w9, 3, dims, str, box)) return SBMC_ECUDA;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  note_path(1);
  if (nt == 128 && cout == 128 && a.pitch == 0 && c3::pair_enabled()) {
    // CTA pairs (cta_group::2): each CTA loads 64 of the 128 output channels per stage
    CUtensorMap wh;
    const uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cout, 9};
    const uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)cin * 2 * cout};
    const uint32_t box[3] = {64, 64, 1};
    if (!encode_tensor_map_bf16_sw128(&wh, w9, 3, dims, str, box)) return SBMC_ECUDA;
    return c3::launch_pair(a, am, wh, st);
  }
  return nt == 256 ? c3::launch<256>(a, am, wm, st) : c3::launch<128>(a, am, wm, st);
}
