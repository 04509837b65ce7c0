// wgrad.cu -- weight gradient of a 1x1-convolution layer as a split-K tcgen05 GEMM:
//
//   dW[co][ci] = sum_r dY[r][co] * X[r][ci]        db[co] = sum_r dY[r][co]
//
// on bf16 channels-innermost rows (r = pixel or sample), fp32 accumulation and output.
// Training-path counterpart of csrc/linear.cu for the per-sample 1x1 ConvChains
// (sbmc/modules.py:34-125; embedding_XX / kernel_regressor of sbmc/models.py:86-102; the
// reference gets these gradients from cuDNN through autograd, sbmc/interfaces.py:78-106).
//
// The reduction runs over the ROWS, which are the slow dimension of both operands in
// memory: dY^T (M = co) and X (N = ci) are "MN-major" UMMA operands.  One TMA box
// {64 channels, 128 rows} with 128-byte swizzle is exactly the canonical MN-major
// SWIZZLE_128B layout (cute/atom/mma_traits_sm100.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units): 64 channels contiguous in a 128-byte line, 8 consecutive rows = one
// 1024-byte swizzle atom (SBO), the next 64 channels in the next box (LBO = box size).
// The instruction descriptor's a_major / b_major bits select the transposed read.
//
// Grid = (K splits, cout / 128, cin / 128): every CTA streams its row range through a
// 3-stage TMA ring (64 KB per stage: 2 boxes of dY, 2 of X), accumulates a 128 x 128
// fp32 block in tensor memory and writes it to a partial buffer; `wgrad_reduce_kernel`
// adds the partials in a fixed order (deterministic, no atomics).  The bias gradient
// is produced by the tensor cores as well: one extra N = 16 MMA per K step against a
// constant block of ones.  HBM-bound: 2 * rows * 128 * 2 bytes per 128 x 128 block
// against 64 cycles of MMA per 16 rows.
#include <cuda_bf16.h>

#include "umma.cuh"

namespace sbmc {
namespace wg {

constexpr int kKP = 128;                    // rows per stage
constexpr int kBox = kKP * 128;             // one {64 ch, 128 rows} box: 16 KB
constexpr int kStage = 4 * kBox;            // A lo, A hi, B lo, B hi
constexpr int kStages = 3;
constexpr int kOnes = 2048;                 // 16 rows x 128 B of bf16 1.0
constexpr int kThreads = 256;

struct Args {
  float *partial;          // [nsplit][Cout][Cin]
  float *partial_b;        // [nsplit][Cout] or null
  long long rows;
  long long rows_per_split;    // multiple of kKP
  int Cout, Cin;
};

enum { B_FULL = 0, B_EMPTY = 3, B_ACC = 6, B_COUNT = 7 };

// Shared-memory descriptor of an MN-major SWIZZLE_128B operand: 64-element (128-byte)
// lines, K groups of 8 lines 1024 B apart (SBO), MN blocks of 64 elements `lbo` bytes apart.
__device__ __forceinline__ uint64_t desc_mn_sw128(const void *p, uint32_t lbo) {
  const uint32_t addr = smem_u32(p);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// bf16 x bf16 -> fp32, BOTH operands MN-major (bits 15 / 16), M x N tile.
__device__ __forceinline__ uint32_t idesc_bf16_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap ymap,      // dY {Cout, rows}
             const __grid_constant__ CUtensorMap xmap,      // X  {Cin, rows}
             const Args P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *stages = smem;
  unsigned char *ones = smem + kStages * kStage;
  uint64_t *bars = reinterpret_cast<uint64_t *>(ones + kOnes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int split = blockIdx.x, cob = blockIdx.y, cib = blockIdx.z;
  const bool with_bias = P.partial_b != nullptr && cib == 0;
  const long long r_lo = (long long)split * P.rows_per_split;
  long long r_hi = r_lo + P.rows_per_split;
  if (r_hi > P.rows) r_hi = P.rows;
  const int nchunks = (r_hi > r_lo) ? (int)((r_hi - r_lo + kKP - 1) / kKP) : 0;

  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < kOnes / 4; i += kThreads)
    reinterpret_cast<uint32_t *>(ones)[i] = 0x3F803F80u;      // bf16 1.0 pairs
  fence_proxy_async();
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ph = 0;
      int st = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int r0 = (int)(r_lo + (long long)c * kKP);
        unsigned char *s = stages + st * kStage;
        mbar_wait(bars + B_EMPTY + st, ((ph >> st) & 1) ^ 1); ph ^= 1u << st;
        mbar_expect_tx(bars + B_FULL + st, (uint32_t)kStage);
        tma_load_2d(s, &ymap, bars + B_FULL + st, cob * 128, r0);
        tma_load_2d(s + kBox, &ymap, bars + B_FULL + st, cob * 128 + 64, r0);
        tma_load_2d(s + 2 * kBox, &xmap, bars + B_FULL + st, cib * 128, r0);
        tma_load_2d(s + 3 * kBox, &xmap, bars + B_FULL + st, cib * 128 + 64, r0);
        st = (st + 1 == kStages) ? 0 : st + 1;
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_bf16_mn(128, 128), idesc1 = idesc_bf16_mn(128, 16);
    const uint64_t dA = desc_mn_sw128(stages, kBox), dB = desc_mn_sw128(stages + 2 * kBox, kBox);
    const uint64_t dOnes = desc_mn_sw128(ones, 0);
    uint32_t ph = 0;
    int st = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(bars + B_FULL + st, (ph >> st) & 1); ph ^= 1u << st;
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t a0 = dA + (uint64_t)st * (kStage >> 4);
        const uint64_t b0 = dB + (uint64_t)st * (kStage >> 4);
#pragma unroll
        for (int k = 0; k < kKP / 16; ++k) {
          // 16 rows further along K: 16 lines of 128 B
          umma_bf16(tmem, a0 + (uint64_t)(k * 128), b0 + (uint64_t)(k * 128), idesc, (c | k) > 0);
          if (with_bias) umma_bf16(tmem + 128, a0 + (uint64_t)(k * 128), dOnes, idesc1, (c | k) > 0);
        }
        umma_commit(bars + B_EMPTY + st);
        if (c == nchunks - 1) umma_commit(bars + B_ACC);
      }
      __syncwarp();
      st = (st + 1 == kStages) ? 0 : st + 1;
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int co = cob * 128 + quad * 32 + lane;
    float *dst = P.partial + ((long long)split * P.Cout + co) * P.Cin + cib * 128;
    if (nchunks > 0) {
      mbar_wait(bars + B_ACC, 0);
      tcgen05_fence_after();
      const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld_32x32b_x32(lane_base + c0, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) stg256(dst + c0 + 8 * k, reinterpret_cast<const uint32_t *>(v) + 8 * k);
      }
      if (with_bias) {
        float v[32];
        tmem_ld_32x32b_x32(lane_base + 128, v);      // columns 128..143 hold the sums, 144.. unused
        P.partial_b[(long long)split * P.Cout + co] = v[0];
      }
    } else {
      const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c0 = 0; c0 < 128; c0 += 8) stg256(dst + c0, z);
      if (with_bias) P.partial_b[(long long)split * P.Cout + co] = 0.f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 256);
}

// Few splits, many outputs (the 3x3 layers): one thread per 4 consecutive outputs, the
// splits added in order.  Needs Cin % 4 == 0, contiguous output (ldw == Cin, no trimming).
__global__ void __launch_bounds__(256)
wgrad_reduce_vec_kernel(const float4 *__restrict__ partial, int nsplit, long long total4,
                        float4 *__restrict__ dw, const float *__restrict__ partial_b, int cout,
                        float *__restrict__ db) {
  if (db && blockIdx.x == gridDim.x - 1) {
    for (int co = threadIdx.x; co < cout; co += 256) {
      float acc = 0.f;
      for (int s = 0; s < nsplit; ++s) acc += partial_b[(long long)s * cout + co];
      db[co] = acc;
    }
  }
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total4; i += 256ll * gridDim.x) {
    float4 acc = __ldg(partial + i);
    for (int s = 1; s < nsplit; ++s) {
      const float4 v = __ldg(partial + s * total4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    dw[i] = acc;
  }
}

// dw[co][ci] (leading dimension ldw) = sum over the splits, co < cout_valid, ci < cin_valid.
// Block = 32 consecutive outputs x 8 warps; warp w adds splits w, w + 8, ... (independent
// coalesced 128-byte loads), the eight partial sums are added in a fixed order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ partial_b,
                    int nsplit, int Cout, int Cin, float *__restrict__ dw, long long ldw,
                    int cout_valid, int cin_valid, float *__restrict__ db) {
  __shared__ float red[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)cout_valid * cin_valid;
  const long long nb = (total + 31) / 32;                 // blocks of weight outputs
  const long long stride = (long long)Cout * Cin;
  for (long long blk = blockIdx.x; blk < nb + (db ? (cout_valid + 31) / 32 : 0); blk += gridDim.x) {
    const bool bias = blk >= nb;
    const long long i = (bias ? blk - nb : blk) * 32 + lane;
    const bool ok = i < (bias ? (long long)cout_valid : total);
    int co = 0, ci = 0;
    const float *p = partial_b;
    long long st = Cout;
    if (!bias) {
      co = (int)(i / cin_valid); ci = (int)(i - (long long)co * cin_valid);
      p = partial + (long long)co * Cin + ci;
      st = stride;
    } else {
      p = partial_b + i;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (ok) {
      int s = warp;
      for (; s + 24 < nsplit; s += 32) {
        a0 += p[s * st]; a1 += p[(s + 8) * st]; a2 += p[(s + 16) * st]; a3 += p[(s + 24) * st];
      }
      for (; s < nsplit; s += 8) a0 += p[s * st];
    }
    red[warp][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (warp == 0 && ok) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) acc += red[w][lane];
      if (bias) db[i] = acc;
      else dw[co * ldw + ci] = acc;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// 3x3 convolutions (the U-net, sbmc/modules.py:248-320):
//   dW[3 dy + dx][co][ci] = sum_{n,y,x} dP[n][y][x][co] * X[n][y + dy - 1][x + dx - 1][ci]
// Same MN-major split-K GEMM; a K chunk is 2 image rows x 64 pixels.  A CTA owns one
// (co block, ci block, dy): per chunk ONE halo box of X ({64 ch, 66 px, 2 rows}, zero fill
// = the convolution's padding) serves the three dx taps as three descriptors whose start
// address is moved by dx rows of 128 bytes (the swizzle is a function of the address), and
// the three taps accumulate side by side in tensor memory (3 x 128 columns).
// ---------------------------------------------------------------------------------------
// Chunk shapes: 2 rows x 64 pixels, or 4 rows x 32 pixels for images at most 32 wide (the
// coarsest U-net level of a 128 x 128 crop), so that no half of the box is padding.
constexpr int k3ABox = 128 * 128;                           // 16 KB per 64-channel half
constexpr int k3BBox = 17 * 1024;                           // 132 / 136 box rows, padded
constexpr int k3Stage = 2 * k3ABox + 2 * k3BBox;

struct Args3 {
  float *partial;          // [nsplit][9][Cout][Cin]
  float *partial_b;        // [nsplit][Cout] or null: bias gradient (sum of dP over the pixels)
  int Cout, Cin, H, W;
  int tiles_x, tiles_y;
  long long nchunks;       // n * tiles_y * tiles_x
  long long chunks_per_split;
};

template <int k3Px, int k3Rows>
__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap ymap,      // dP {Cout, W, H, n}
                const __grid_constant__ CUtensorMap xmap,      // X  {Cin, W, H, n}
                const Args3 P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *stages = smem;
  unsigned char *ones = smem + kStages * k3Stage;
  uint64_t *bars = reinterpret_cast<uint64_t *>(ones + kOnes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int split = blockIdx.x, dy = blockIdx.z;
  const int ci_blocks = P.Cin / 128;
  const int cob = blockIdx.y / ci_blocks, cib = blockIdx.y % ci_blocks;
  // the bias gradient rides on the (ci block 0, dy 0) CTAs: one N = 16 MMA per K step
  // against a block of ones (columns 384..399 of tensor memory)
  const bool with_bias = P.partial_b != nullptr && cib == 0 && dy == 0;
  const long long c_lo = (long long)split * P.chunks_per_split;
  long long c_hi = c_lo + P.chunks_per_split;
  if (c_hi > P.nchunks) c_hi = P.nchunks;
  const int nchunks = (c_hi > c_lo) ? (int)(c_hi - c_lo) : 0;

  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < kOnes / 4; i += kThreads)
    reinterpret_cast<uint32_t *>(ones)[i] = 0x3F803F80u;      // bf16 1.0 pairs
  fence_proxy_async();
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ph = 0;
      int st = 0;
      for (int c = 0; c < nchunks; ++c) {
        const long long ch = c_lo + c;
        const int tx = (int)(ch % P.tiles_x);
        const int ty = (int)((ch / P.tiles_x) % P.tiles_y);
        const int n = (int)(ch / ((long long)P.tiles_x * P.tiles_y));
        const int x0 = tx * k3Px, y0 = ty * k3Rows;
        unsigned char *s = stages + st * k3Stage;
        mbar_wait(bars + B_EMPTY + st, ((ph >> st) & 1) ^ 1); ph ^= 1u << st;
        mbar_expect_tx(bars + B_FULL + st, (uint32_t)(2 * k3ABox + 2 * k3Rows * (k3Px + 2) * 128));
        tma_load_4d(s, &ymap, bars + B_FULL + st, cob * 128, x0, y0, n);
        tma_load_4d(s + k3ABox, &ymap, bars + B_FULL + st, cob * 128 + 64, x0, y0, n);
        tma_load_4d(s + 2 * k3ABox, &xmap, bars + B_FULL + st, cib * 128, x0 - 1, y0 + dy - 1, n);
        tma_load_4d(s + 2 * k3ABox + k3BBox, &xmap, bars + B_FULL + st, cib * 128 + 64, x0 - 1,
                    y0 + dy - 1, n);
        st = (st + 1 == kStages) ? 0 : st + 1;
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_bf16_mn(128, 128), idesc1 = idesc_bf16_mn(128, 16);
    const uint64_t dA = desc_mn_sw128(stages, k3ABox);
    const uint64_t dB = desc_mn_sw128(stages + 2 * k3ABox, k3BBox);
    const uint64_t dOnes = desc_mn_sw128(ones, 0);
    uint32_t ph = 0;
    int st = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(bars + B_FULL + st, (ph >> st) & 1); ph ^= 1u << st;
      tcgen05_fence_after();
      if (elect_one()) {
        const uint64_t a0 = dA + (uint64_t)st * (k3Stage >> 4);
        const uint64_t b0 = dB + (uint64_t)st * (k3Stage >> 4);
#pragma unroll
        for (int r = 0; r < k3Rows; ++r)
#pragma unroll
          for (int j = 0; j < k3Px / 16; ++j) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              umma_bf16(tmem + dx * 128, a0 + (uint64_t)((r * k3Px + 16 * j) * 8),
                        b0 + (uint64_t)((r * (k3Px + 2) + 16 * j + dx) * 8), idesc, (c | r | j) > 0);
            if (with_bias)
              umma_bf16(tmem + 384, a0 + (uint64_t)((r * k3Px + 16 * j) * 8), dOnes, idesc1,
                        (c | r | j) > 0);
          }
        umma_commit(bars + B_EMPTY + st);
        if (c == nchunks - 1) umma_commit(bars + B_ACC);
      }
      __syncwarp();
      st = (st + 1 == kStages) ? 0 : st + 1;
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int co = cob * 128 + quad * 32 + lane;
    if (nchunks > 0) {
      mbar_wait(bars + B_ACC, 0);
      tcgen05_fence_after();
    }
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int dx = 0; dx < 3; ++dx) {
      float *dst = P.partial + (((long long)split * 9 + 3 * dy + dx) * P.Cout + co) * P.Cin + cib * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        if (nchunks > 0) {
          tmem_ld_32x32b_x32(lane_base + dx * 128 + c0, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) stg256(dst + c0 + 8 * k, reinterpret_cast<const uint32_t *>(v) + 8 * k);
      }
    }
    if (with_bias) {
      float v[32];
      v[0] = 0.f;
      if (nchunks > 0) tmem_ld_32x32b_x32(lane_base + 384, v);
      P.partial_b[(long long)split * P.Cout + co] = v[0];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace wg
}  // namespace sbmc

extern "C" int sbmc_wgrad_nhwc_bf16(const void *dy, const void *x, int64_t x_row_pitch,
                                    int64_t rows, int cout, int cin, int nsplit,
                                    float *workspace, float *dw, int64_t ldw, int cout_valid,
                                    int cin_valid, float *db, void *stream) {
  using namespace sbmc;
  if (rows < 1 || cout < 1 || cin < 1 || nsplit < 1 || nsplit > 65535 || ldw < 1) {
    set_error("wgrad: invalid shape");
    return SBMC_EINVAL;
  }
  if (!dy || !x || !workspace || !dw) {
    set_error("wgrad: null pointer argument");
    return SBMC_EINVAL;
  }
  if (cout % 128 != 0 || cin % 128 != 0 || rows >= (1ll << 31) || x_row_pitch < cin ||
      x_row_pitch % 8 != 0) {
    set_error("wgrad: needs cout %% 128 == 0 and cin %% 128 == 0 (got %d, %d)", cout, cin);
    return SBMC_EUNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) |
       reinterpret_cast<uintptr_t>(workspace)) & 31) {
    set_error("wgrad: pointers must be 32-byte aligned");
    return SBMC_EALIGN;
  }
  if (cout_valid < 1 || cout_valid > cout) cout_valid = cout;
  if (cin_valid < 1 || cin_valid > cin) cin_valid = cin;
  wg::Args a;
  a.partial = workspace;
  a.partial_b = db ? workspace + (size_t)nsplit * cout * cin : nullptr;
  a.rows = rows;
  a.rows_per_split = ceil_div(ceil_div(rows, nsplit), wg::kKP) * wg::kKP;
  a.Cout = cout; a.Cin = cin;
  CUtensorMap ym, xm;
  const uint64_t ydims[2] = {(uint64_t)cout, (uint64_t)rows}, ystr[1] = {(uint64_t)cout * 2};
  const uint64_t xdims[2] = {(uint64_t)cin, (uint64_t)rows}, xstr[1] = {(uint64_t)x_row_pitch * 2};
  const uint32_t box[2] = {64, (uint32_t)wg::kKP};
  if (!encode_tensor_map_bf16_sw128(&ym, dy, 2, ydims, ystr, box) ||
      !encode_tensor_map_bf16_sw128(&xm, x, 2, xdims, xstr, box))
    return SBMC_ECUDA;
  const size_t smem = (size_t)wg::kStages * wg::kStage + wg::kOnes + wg::B_COUNT * sizeof(uint64_t) + 16;
  static bool configured = false;
  if (!configured) {
    SBMC_CUDA_OK(cudaFuncSetAttribute(wg::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    configured = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_CONV1X1, st);
    wg::wgrad_kernel<<<dim3((unsigned)nsplit, cout / 128, cin / 128), wg::kThreads, smem, st>>>(ym, xm, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  const long long total = (long long)cout_valid * cin_valid;
  const long long rblocks = (total + 31) / 32 + (db ? (cout_valid + 31) / 32 : 0);
  const unsigned blocks = (unsigned)(rblocks > 148 * 8 ? 148 * 8 : rblocks);
  wg::wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(a.partial, a.partial_b, nsplit, cout, cin, dw, ldw,
                                                  cout_valid, cin_valid, db);
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}

extern "C" int sbmc_wgrad3x3_nhwc_bf16(const void *dp, const void *x, int64_t n, int h, int w,
                                       int cout, int cin, int nsplit, float *workspace,
                                       float *dw9, float *db, void *stream) {
  using namespace sbmc;
  if (n < 1 || h < 1 || w < 1 || cout < 1 || cin < 1 || nsplit < 1 || nsplit > 65535) {
    set_error("wgrad3x3: invalid shape");
    return SBMC_EINVAL;
  }
  if (!dp || !x || !workspace || !dw9) {
    set_error("wgrad3x3: null pointer argument");
    return SBMC_EINVAL;
  }
  if (cout % 128 != 0 || cin % 128 != 0 || n >= (1ll << 31)) {
    set_error("wgrad3x3: needs cout %% 128 == 0 and cin %% 128 == 0 (got %d, %d)", cout, cin);
    return SBMC_EUNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(dp) | reinterpret_cast<uintptr_t>(x) |
       reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(dw9)) & 31) {
    set_error("wgrad3x3: pointers must be 32-byte aligned");
    return SBMC_EALIGN;
  }
  wg::Args3 a;
  a.partial = workspace;
  a.partial_b = db ? workspace + (size_t)nsplit * 9 * cout * cin : nullptr;
  a.Cout = cout; a.Cin = cin; a.H = h; a.W = w;
  const int px = (w <= 32) ? 32 : 64, rows = 128 / px;
  a.tiles_x = (w + px - 1) / px;
  a.tiles_y = (h + rows - 1) / rows;
  a.nchunks = (long long)n * a.tiles_x * a.tiles_y;
  a.chunks_per_split = ceil_div(a.nchunks, nsplit);
  CUtensorMap ym, xm;
  {
    const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)cout * 2, (uint64_t)cout * 2 * w, (uint64_t)cout * 2 * w * h};
    const uint32_t box[4] = {64, (uint32_t)px, (uint32_t)rows, 1};
    if (!encode_tensor_map_bf16_sw128(&ym, dp, 4, dims, str, box)) return SBMC_ECUDA;
  }
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)cin * 2 * w, (uint64_t)cin * 2 * w * h};
    const uint32_t box[4] = {64, (uint32_t)px + 2, (uint32_t)rows, 1};
    if (!encode_tensor_map_bf16_sw128(&xm, x, 4, dims, str, box)) return SBMC_ECUDA;
  }
  const size_t smem = (size_t)wg::kStages * wg::k3Stage + wg::kOnes + wg::B_COUNT * sizeof(uint64_t) + 16;
  static bool configured = false;
  if (!configured) {
    SBMC_CUDA_OK(cudaFuncSetAttribute(wg::wgrad3x3_kernel<64, 2>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SBMC_CUDA_OK(cudaFuncSetAttribute(wg::wgrad3x3_kernel<32, 4>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_CONV3X3, st);
    const dim3 grid((unsigned)nsplit, (cout / 128) * (cin / 128), 3);
    if (px == 32)
      wg::wgrad3x3_kernel<32, 4><<<grid, wg::kThreads, smem, st>>>(ym, xm, a);
    else
      wg::wgrad3x3_kernel<64, 2><<<grid, wg::kThreads, smem, st>>>(ym, xm, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  // partial is [nsplit][9 * cout][cin]
  const long long total = 9ll * cout * cin;
  if (nsplit <= 64) {
    const long long total4 = total / 4;
    const long long b = (total4 + 255) / 256;
    wg::wgrad_reduce_vec_kernel<<<(unsigned)(b > 148 * 16 ? 148 * 16 : b), 256, 0, st>>>(
        reinterpret_cast<const float4 *>(workspace), nsplit, total4, reinterpret_cast<float4 *>(dw9),
        a.partial_b, cout, db);
  } else {
    const long long rblocks = (total + 31) / 32;
    const unsigned blocks = (unsigned)(rblocks > 148 * 8 ? 148 * 8 : rblocks);
    wg::wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(workspace, nullptr, nsplit, 9 * cout, cin, dw9,
                                                    cin, 9 * cout, cin, nullptr);
    if (db)      // bias partials: the generic reduction with zero weight outputs (cin_valid = 0)
      wg::wgrad_reduce_kernel<<<(cout + 31) / 32, 256, 0, st>>>(workspace, a.partial_b, nsplit, cout,
                                                               cin, dw9, cin, cout, 0, db);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}
