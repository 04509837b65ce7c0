// conv1x1.cu -- fused per-sample 1x1 ConvChain on the 5th-gen tensor cores.
//
// The reference's per-sample networks (sbmc/models.py:86-102: embedding_XX
// 96|256 -> 128 -> 128 -> 128 and kernel_regressor 256 -> 128 -> 128 -> 441, all
// 1x1 convolutions, modules.py:34-125) are dense contractions over the channel
// axis applied independently to every pixel: a 3-layer MLP per pixel.  cuDNN
// runs them as three convolutions + three activation kernels with fp32 NCHW
// round trips through HBM in between.  Here one kernel does the whole chain for
// a tile of 128 pixels:
//
//   * pixels are the MMA M dimension (128 = one TMEM lane per pixel / thread),
//     output channels the N dimension, input channels K;
//   * the fp32 NCHW input (up to two tensors, concatenated on the fly -- the
//     reference materialises th.cat([features, propagated]) -- the second one
//     optionally a per-image broadcast vector for the global features) is read
//     with coalesced loads, converted to bf16 and written by each thread as the
//     K-major row of its pixel into 128B-swizzled shared memory (the canonical
//     UMMA A-operand layout);
//   * weights (bf16, K-major, weight-norm folded by the host) come in by TMA
//     with a SWIZZLE_128B tensor map; W1 / W2 stay resident, W3 streams through
//     two 32 KB buffers in 128-row chunks;
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N<=128 per instruction, K=16)
//     accumulates in TMEM; the epilogue of a hidden layer reads the accumulator
//     with tcgen05.ld, adds the bias, applies ReLU / LeakyReLU, converts to bf16
//     and writes the next layer's A operand -- the 128-channel activations never
//     leave the SM;
//   * the last layer's accumulator (fp32) + bias is stored straight to the
//     fp32 NCHW output (coalesced: a warp writes 32 consecutive pixels of one
//     channel), so the K*K logits keep fp32 accumulation and fp32 storage
//     (SURVEY.md section 8a note 4).
#include <cuda_bf16.h>

#include "umma.cuh"

namespace sbmc {

constexpr int kHid = 128;            // hidden width of the chain (models.py:56: width=128)
constexpr int kTileP = 128;          // pixels per tile = MMA M
constexpr int kSlab = 128 * 128;     // bytes of a 128-row x 64-bf16 slab (16 KB)

struct ChainArgs {
  const float *xa, *xb;              // input sources (xb may be null)
  long long a_img, b_img;            // elements between consecutive images
  int ca, cb, b_bcast;               // channels; xb is [n][cb] broadcast over pixels
  const float *b1, *b2, *b3;         // biases (fp32; b3 has n3p entries)
  float *y;                          // output [n][cout][hw]
  long long y_img;
  int cout, n3p, act;                // act: 0 ReLU, 1 LeakyReLU(0.01)
  long long hw, tiles_per_img, ntiles;
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&v);
}

__device__ __forceinline__ float activate(float v, int act) {
  return act ? (v > 0.f ? v : 0.01f * v) : fmaxf(v, 0.f);
}

// Hidden-layer epilogue: TMEM accumulator [128 x 128] -> bias, activation, bf16 ->
// K-major swizzled A operand (2 slabs of 64 channels).
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_row, const float *bias, int act,
                                                unsigned char *dstA, int row) {
#pragma unroll 1
  for (int c0 = 0; c0 < kHid; c0 += 32) {
    float v[32];
    tmem_ld_32x32b_x32(tmem_row + c0, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 q;
      const int c = c0 + 8 * j;
      q.x = pack_bf16(activate(v[8 * j + 0] + bias[c + 0], act), activate(v[8 * j + 1] + bias[c + 1], act));
      q.y = pack_bf16(activate(v[8 * j + 2] + bias[c + 2], act), activate(v[8 * j + 3] + bias[c + 3], act));
      q.z = pack_bf16(activate(v[8 * j + 4] + bias[c + 4], act), activate(v[8 * j + 5] + bias[c + 5], act));
      q.w = pack_bf16(activate(v[8 * j + 6] + bias[c + 6], act), activate(v[8 * j + 7] + bias[c + 7], act));
      const int chunk = c >> 3;                  // 16-byte chunk index along K (0..15)
      *reinterpret_cast<uint4 *>(dstA + (chunk >> 3) * kSlab + sw128_offset(row, chunk & 7)) = q;
    }
  }
}

// K1P: padded input channels (128 or 256).
template <int K1P>
__global__ void __launch_bounds__(128, 1)
conv1x1_chain_kernel(const __grid_constant__ CUtensorMap w1map,
                     const __grid_constant__ CUtensorMap w2map,
                     const __grid_constant__ CUtensorMap w3map, const ChainArgs P) {
  constexpr int KS1 = K1P / 64;                  // slabs of the first layer's operands
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *sW1 = smem;
  unsigned char *sW2 = sW1 + KS1 * kSlab;
  unsigned char *sA0 = sW2 + 2 * kSlab;
  unsigned char *sA1 = sA0 + KS1 * kSlab;
  unsigned char *sX = sA1 + 2 * kSlab;           // only carved when K1P == 128
  unsigned char *buf0 = (K1P == 256) ? (sA0 + 2 * kSlab) : sX;   // W3 chunk buffers
  unsigned char *buf1 = sA1;
  float *sB = reinterpret_cast<float *>((K1P == 256) ? sX : (sX + 2 * kSlab));  // biases
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + 2 * kHid + 512);
  uint64_t *bar_w = bars, *bar_mma = bars + 1, *bar_w3 = bars + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_w3, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < kHid; i += 128) {
    sB[i] = P.b1[i];
    sB[kHid + i] = P.b2[i];
  }
  for (int i = tid; i < P.n3p; i += 128) sB[2 * kHid + i] = P.b3[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);

  if (tid == 0) {                                 // resident weights
    mbar_expect_tx(bar_w, (uint32_t)((KS1 + 2) * kSlab));
    for (int kb = 0; kb < KS1; ++kb) tma_load_2d(sW1 + kb * kSlab, &w1map, bar_w, kb * 64, 0);
    for (int kb = 0; kb < 2; ++kb) tma_load_2d(sW2 + kb * kSlab, &w2map, bar_w, kb * 64, 0);
  }
  const int nchunks = (P.n3p + 127) / 128;        // W3 chunks of up to 128 output channels
  const int w3_box_bytes = (P.n3p < 128 ? P.n3p : 128) * 128;   // one 64-channel slab of a chunk
  uint32_t ph_mma = 0, ph_w3 = 0;
  bool weights_ready = false;

  for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long n = tile / P.tiles_per_img;
    const long long p = (tile - n * P.tiles_per_img) * kTileP + tid;
    const bool valid = p < P.hw;

    // ---- prologue: fp32 NCHW -> bf16 K-major swizzled rows (thread = pixel) ----
    {
      const float *pa = P.xa + n * P.a_img + p;
      const float *pb = P.xb ? (P.b_bcast ? P.xb + n * P.b_img : P.xb + n * P.b_img + p) : nullptr;
      const int cin = P.ca + P.cb;
#pragma unroll 4
      for (int c0 = 0; c0 < K1P; c0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j;
          float x = 0.f;
          if (valid && c < cin) {
            if (c < P.ca) x = __ldg(pa + (long long)c * P.hw);
            else if (P.b_bcast) x = __ldg(pb + (c - P.ca));
            else x = __ldg(pb + (long long)(c - P.ca) * P.hw);
          }
          v[j] = x;
        }
        uint4 q;
        q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
        q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
        const int chunk = c0 >> 3;
        *reinterpret_cast<uint4 *>(sA0 + (chunk >> 3) * kSlab + sw128_offset(tid, chunk & 7)) = q;
      }
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();

    // ---- layer 1: D[128 x 128] = A0 . W1^T ----
    if (tid == 0) {
      if (!weights_ready) mbar_wait(bar_w, 0);
      tcgen05_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, kHid);
#pragma unroll 1
      for (int k = 0; k < K1P / 16; ++k) {
        const uint64_t ad = umma_smem_desc_sw128(sA0 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        const uint64_t bd = umma_smem_desc_sw128(sW1 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        umma_bf16(tmem, ad, bd, idesc, k > 0);
      }
      umma_commit(bar_mma);
    }
    weights_ready = true;
    mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
    tcgen05_fence_after();
    hidden_epilogue(tmem_row, sB, P.act, sA1, tid);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();

    // ---- layer 2: D = A1 . W2^T ----
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, kHid);
#pragma unroll 1
      for (int k = 0; k < kHid / 16; ++k) {
        const uint64_t ad = umma_smem_desc_sw128(sA1 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        const uint64_t bd = umma_smem_desc_sw128(sW2 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        umma_bf16(tmem, ad, bd, idesc, k > 0);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
    tcgen05_fence_after();
    // both W3 buffers are free now (A0's upper half / sX since MMA1, A1 since MMA2)
    if (tid == 0) {
      const int nb = nchunks < 2 ? nchunks : 2;
      // (a box that hangs over the last row is zero-filled and still counts in full)
      mbar_expect_tx(bar_w3, (uint32_t)(nb * 2 * w3_box_bytes));
      for (int ch = 0; ch < nb; ++ch)
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d((ch ? buf1 : buf0) + kb * kSlab, &w3map, bar_w3, kb * 64, ch * 128);
    }
    hidden_epilogue(tmem_row, sB + kHid, P.act, sA0, tid);   // A2 aliases A0's first 32 KB
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();

    // ---- layer 3 in chunks of up to 128 output channels; chunk pairs share a commit ----
    float *yp = P.y + n * P.y_img + p;
    for (int c0 = 0; c0 < nchunks; c0 += 2) {
      const int nb = (nchunks - c0 < 2) ? (nchunks - c0) : 2;
      if (tid == 0) {
        mbar_wait(bar_w3, ph_w3);
        tcgen05_fence_after();
        for (int ch = 0; ch < nb; ++ch) {
          const int rows = (P.n3p - (c0 + ch) * 128 < 128) ? (P.n3p - (c0 + ch) * 128) : 128;
          const uint32_t idesc = umma_idesc_bf16(128, rows);
          unsigned char *wb = ch ? buf1 : buf0;
#pragma unroll 1
          for (int k = 0; k < kHid / 16; ++k) {
            const uint64_t ad = umma_smem_desc_sw128(sA0 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
            const uint64_t bd = umma_smem_desc_sw128(wb + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
            umma_bf16(tmem + (uint32_t)((c0 + ch) * 128), ad, bd, idesc, k > 0);
          }
        }
        umma_commit(bar_mma);
      }
      ph_w3 ^= 1;
      mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
      tcgen05_fence_after();
      if (tid == 0 && c0 + 2 < nchunks) {          // next chunk pair streams in during the stores
        const int nn = (nchunks - c0 - 2 < 2) ? (nchunks - c0 - 2) : 2;
        mbar_expect_tx(bar_w3, (uint32_t)(nn * 2 * w3_box_bytes));
        for (int ch = 0; ch < nn; ++ch)
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d((ch ? buf1 : buf0) + kb * kSlab, &w3map, bar_w3, kb * 64, (c0 + 2 + ch) * 128);
      }
      // ---- output epilogue for these chunks: + bias, fp32 NCHW store ----
      const int col_end = (c0 + nb) * 128 < P.n3p ? (c0 + nb) * 128 : P.n3p;
#pragma unroll 1
      for (int col = c0 * 128; col < col_end; col += 32) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_row + col, v);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col + i < P.cout) yp[(long long)(col + i) * P.hw] = v[i] + sB[2 * kHid + col + i];
        }
      }
    }
    tcgen05_fence_before();   // the next tile's MMA1 overwrites the accumulator
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int K1P>
static int run_chain(const ChainArgs &a, const void *w1, const void *w2, const void *w3,
                     cudaStream_t st) {
  CUtensorMap m1, m2, m3;
  if (!encode_tensor_map_bf16_2d_sw128(&m1, w1, K1P, kHid, 64, kHid) ||
      !encode_tensor_map_bf16_2d_sw128(&m2, w2, kHid, kHid, 64, kHid) ||
      !encode_tensor_map_bf16_2d_sw128(&m3, w3, kHid, (uint64_t)a.n3p, 64,
                                       (uint32_t)(a.n3p < 128 ? a.n3p : 128)))
    return SBMC_ECUDA;
  constexpr int KS1 = K1P / 64;
  const size_t smem = (size_t)(KS1 + 2 + KS1 + 2 + (K1P == 256 ? 0 : 2)) * kSlab +
                      (2 * kHid + 512) * sizeof(float) + 64 + 1024;
  auto kern = conv1x1_chain_kernel<K1P>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  {
    KernelTimer timer(SBMC_KERNEL_CONV1X1, st);
    kern<<<(unsigned)grid, 128, smem, st>>>(m1, m2, m3, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

// ===========================================================================
// NHWC-bf16 variant: the inference pipeline keeps every per-sample / per-pixel
// activation as bf16 with channels innermost ([n][pixel][128]), which IS the
// K-major UMMA A-operand order.  One TMA box (64 channels x 128 pixels,
// SWIZZLE_128B) is one operand slab, so the input needs no thread at all: the
// producer thread issues the next tile's loads as soon as the first layer's
// MMAs have retired, and they land while the rest of the chain runs.  Output is
// either bf16 NHWC (embeddings: 256 contiguous bytes per pixel / thread) or fp32
// NCHW (the K*K logits for the splat kernels).
// ===========================================================================
struct ChainArgsV2 {
  const float *b1;  long long b1_img;    // first-layer bias, optionally per image
  const float *b2, *b3;
  void *y;          long long y_img;     // elements between output images
  int out_nhwc_bf16;                     // 1: bf16 [n][hw][128]; 0: fp32 [n][cout][hw]
  int cout, n3p, act;
  long long hw, tiles_per_img, ntiles;
};

// Epilogue helpers of the NHWC kernel: 256 threads, thread t owns pixel (t & 127)
// and the column half (t >> 7) of every 128-column accumulator block; biases are
// read from shared memory (float4 broadcasts).
__device__ __forceinline__ void hidden_epilogue_half(uint32_t tmem_row, const float *sbias,
                                                     int act, unsigned char *dstA, int row,
                                                     int half) {
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int c0 = half * 64 + it * 32;
    float v[32];
    tmem_ld_32x32b_x32(tmem_row + c0, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + 8 * j;
      const float4 ba = *reinterpret_cast<const float4 *>(sbias + c);
      const float4 bb = *reinterpret_cast<const float4 *>(sbias + c + 4);
      uint4 q;
      q.x = pack_bf16(activate(v[8 * j + 0] + ba.x, act), activate(v[8 * j + 1] + ba.y, act));
      q.y = pack_bf16(activate(v[8 * j + 2] + ba.z, act), activate(v[8 * j + 3] + ba.w, act));
      q.z = pack_bf16(activate(v[8 * j + 4] + bb.x, act), activate(v[8 * j + 5] + bb.y, act));
      q.w = pack_bf16(activate(v[8 * j + 6] + bb.z, act), activate(v[8 * j + 7] + bb.w, act));
      const int chunk = c >> 3;
      *reinterpret_cast<uint4 *>(dstA + (chunk >> 3) * kSlab + sw128_offset(row, chunk & 7)) = q;
    }
  }
}

// KS1: 64-channel slabs of the first layer's K (2: one source, 4: two sources).
template <int KS1>
__global__ void __launch_bounds__(256, 1)
conv1x1_chain_nhwc_kernel(const __grid_constant__ CUtensorMap amap,
                          const __grid_constant__ CUtensorMap bmap,
                          const __grid_constant__ CUtensorMap w1map,
                          const __grid_constant__ CUtensorMap w2map,
                          const __grid_constant__ CUtensorMap w3map, const ChainArgsV2 P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sW1 = smem;
  unsigned char *sW2 = sW1 + KS1 * kSlab;
  unsigned char *sA0 = sW2 + 2 * kSlab;          // TMA destination, KS1 slabs
  unsigned char *sA12 = sA0 + KS1 * kSlab;       // hidden activations (layer 1, then 2)
  unsigned char *sW3 = sA12 + 2 * kSlab;         // one chunk of up to 128 output channels
  float *sB1 = reinterpret_cast<float *>(sW3 + 2 * kSlab);
  float *sB2 = sB1 + kHid;
  float *sB3 = sB2 + kHid;                       // up to 448 entries
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB3 + 448);
  uint64_t *bar_w = bars, *bar_mma = bars + 1, *bar_w3 = bars + 2, *bar_a = bars + 3;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;
  if ((smem_u32(smem) & 1023u) != 0) __trap();    // SWIZZLE_128B slabs need 1024-byte alignment
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_w3, 1);
    mbar_init(bar_a, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid < kHid) sB2[tid] = P.b2[tid];
  for (int i = tid; i < 448; i += 256) sB3[i] = i < P.n3p ? P.b3[i] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  const int nchunks = (P.n3p + 127) / 128;
  const uint32_t w3_bytes = (uint32_t)(2 * (P.n3p < 128 ? P.n3p : 128) * 128);

  auto load_tile = [&](long long tile) {          // tid 0: A operand of `tile`
    const int n = (int)(tile / P.tiles_per_img);
    const int p0 = (int)((tile - (long long)n * P.tiles_per_img) * kTileP);
    mbar_expect_tx(bar_a, (uint32_t)(KS1 * kSlab));
    tma_load_3d(sA0, &amap, bar_a, 0, p0, n);
    tma_load_3d(sA0 + kSlab, &amap, bar_a, 64, p0, n);
    if (KS1 == 4) {
      tma_load_3d(sA0 + 2 * kSlab, &bmap, bar_a, 0, p0, n);
      tma_load_3d(sA0 + 3 * kSlab, &bmap, bar_a, 64, p0, n);
    }
  };
  auto load_w3 = [&](int chunk) {                 // tid 0
    mbar_expect_tx(bar_w3, w3_bytes);
    tma_load_2d(sW3, &w3map, bar_w3, 0, chunk * 128);
    tma_load_2d(sW3 + kSlab, &w3map, bar_w3, 64, chunk * 128);
  };

  if (tid == 0) {
    mbar_expect_tx(bar_w, (uint32_t)((KS1 + 2) * kSlab));
    for (int kb = 0; kb < KS1; ++kb) tma_load_2d(sW1 + kb * kSlab, &w1map, bar_w, kb * 64, 0);
    for (int kb = 0; kb < 2; ++kb) tma_load_2d(sW2 + kb * kSlab, &w2map, bar_w, kb * 64, 0);
    load_w3(0);
    if ((long long)blockIdx.x < P.ntiles) load_tile(blockIdx.x);
  }
  uint32_t ph_mma = 0, ph_w3 = 0, ph_a = 0;
  bool first = true;
  long long cur_n = -1;

  for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long n = tile / P.tiles_per_img;
    const long long p = (tile - n * P.tiles_per_img) * kTileP + row;
    const bool valid = p < P.hw;
    if (n != cur_n) {                             // (per-image) first-layer bias
      if (tid < kHid) sB1[tid] = P.b1[n * P.b1_img + tid];
      cur_n = n;
      __syncthreads();
    }

    // ---- layer 1 ----
    if (tid == 0) {
      if (first) mbar_wait(bar_w, 0);
      mbar_wait(bar_a, ph_a);
      tcgen05_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, kHid);
#pragma unroll 1
      for (int k = 0; k < KS1 * 4; ++k) {
        const uint64_t ad = umma_smem_desc_sw128(sA0 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        const uint64_t bd = umma_smem_desc_sw128(sW1 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        umma_bf16(tmem, ad, bd, idesc, k > 0);
      }
      umma_commit(bar_mma);
    }
    first = false;
    ph_a ^= 1;
    mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
    tcgen05_fence_after();
    if (tid == 0 && tile + gridDim.x < P.ntiles) load_tile(tile + gridDim.x);  // A0 is free
    hidden_epilogue_half(tmem_row, sB1, P.act, sA12, row, half);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();

    // ---- layer 2 ----
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, kHid);
#pragma unroll 1
      for (int k = 0; k < kHid / 16; ++k) {
        const uint64_t ad = umma_smem_desc_sw128(sA12 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        const uint64_t bd = umma_smem_desc_sw128(sW2 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
        umma_bf16(tmem, ad, bd, idesc, k > 0);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
    tcgen05_fence_after();
    hidden_epilogue_half(tmem_row, sB2, P.act, sA12, row, half);   // overwrites layer 1's operand
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();

    // ---- layer 3, chunk by chunk ----
    for (int ch = 0; ch < nchunks; ++ch) {
      const int rows = (P.n3p - ch * 128 < 128) ? (P.n3p - ch * 128) : 128;
      if (tid == 0) {
        if (nchunks > 1 || tile == (long long)blockIdx.x) mbar_wait(bar_w3, ph_w3);
        tcgen05_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, rows);
#pragma unroll 1
        for (int k = 0; k < kHid / 16; ++k) {
          const uint64_t ad = umma_smem_desc_sw128(sA12 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
          const uint64_t bd = umma_smem_desc_sw128(sW3 + (k >> 2) * kSlab) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + (uint32_t)(ch * 128), ad, bd, idesc, k > 0);
        }
        umma_commit(bar_mma);
      }
      if (nchunks > 1) ph_w3 ^= 1;
      mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
      tcgen05_fence_after();
      if (tid == 0 && nchunks > 1) {
        // the chunk buffer is free: stream the next chunk (or chunk 0 of the next tile)
        if (ch + 1 < nchunks) load_w3(ch + 1);
        else if (tile + gridDim.x < P.ntiles) load_w3(0);
      }
      if (P.out_nhwc_bf16) {
        // bf16 NHWC: this thread's pixel, its 64 channels = 128 contiguous bytes
        uint4 *dst = reinterpret_cast<uint4 *>(
            reinterpret_cast<__nv_bfloat16 *>(P.y) + n * P.y_img + p * kHid);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int c0 = half * 64 + it * 32;
          float v[32];
          tmem_ld_32x32b_x32(tmem_row + c0, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c0 + 8 * j;
              const float4 ba = *reinterpret_cast<const float4 *>(sB3 + c);
              const float4 bb = *reinterpret_cast<const float4 *>(sB3 + c + 4);
              uint4 q;
              q.x = pack_bf16(v[8 * j + 0] + ba.x, v[8 * j + 1] + ba.y);
              q.y = pack_bf16(v[8 * j + 2] + ba.z, v[8 * j + 3] + ba.w);
              q.z = pack_bf16(v[8 * j + 4] + bb.x, v[8 * j + 5] + bb.y);
              q.w = pack_bf16(v[8 * j + 6] + bb.z, v[8 * j + 7] + bb.w);
              dst[c >> 3] = q;
            }
          }
        }
      } else {
        // fp32 NCHW: a warp stores 32 consecutive pixels of one channel
        float *yp = reinterpret_cast<float *>(P.y) + n * P.y_img + p;
        const int base = ch * 128 + half * 64;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int col = base + it * 32;
          if (col < ch * 128 + rows) {             // warp-uniform
            float v[32];
            tmem_ld_32x32b_x32(tmem_row + col, v);
            if (valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col + i < P.cout) yp[(long long)(col + i) * P.hw] = v[i] + sB3[col + i];
            }
          }
        }
      }
    }
    tcgen05_fence_before();   // the next tile's MMAs overwrite the accumulator
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int KS1>
static int run_chain_v2(const ChainArgsV2 &a, const void *xa, long long a_img, const void *xb,
                        long long b_img, long long n_img, const void *w1, const void *w2,
                        const void *w3, cudaStream_t st) {
  CUtensorMap ma, mb, m1, m2, m3;
  if (!encode_tensor_map_bf16_3d_sw128(&ma, xa, kHid, (uint64_t)a.hw, (uint64_t)n_img,
                                       (uint64_t)a_img, kTileP))
    return SBMC_ECUDA;
  mb = ma;
  if (KS1 == 4 && !encode_tensor_map_bf16_3d_sw128(&mb, xb, kHid, (uint64_t)a.hw,
                                                   (uint64_t)n_img, (uint64_t)b_img, kTileP))
    return SBMC_ECUDA;
  if (!encode_tensor_map_bf16_2d_sw128(&m1, w1, KS1 * 64, kHid, 64, kHid) ||
      !encode_tensor_map_bf16_2d_sw128(&m2, w2, kHid, kHid, 64, kHid) ||
      !encode_tensor_map_bf16_2d_sw128(&m3, w3, kHid, (uint64_t)a.n3p, 64,
                                       (uint32_t)(a.n3p < 128 ? a.n3p : 128)))
    return SBMC_ECUDA;
  const size_t smem = (size_t)(KS1 + 2 + KS1 + 2 + 2) * kSlab + (2 * kHid + 448) * 4 + 64;
  auto kern = conv1x1_chain_nhwc_kernel<KS1>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  {
    KernelTimer timer(SBMC_KERNEL_CONV1X1, st);
    kern<<<(unsigned)grid, 256, smem, st>>>(ma, mb, m1, m2, m3, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace sbmc

extern "C" int sbmc_conv1x1_chain_f32(const float *xa, int ca, int64_t a_img_stride,
                                      const float *xb, int cb, int64_t b_img_stride,
                                      int b_broadcast, const void *w1, const float *b1,
                                      const void *w2, const float *b2, const void *w3,
                                      const float *b3, int k1p, int cout, int n3p, int act,
                                      float *y, int64_t y_img_stride, int64_t n_img,
                                      int64_t hw, void *stream) {
  using namespace sbmc;
  if (n_img < 0 || hw < 0 || ca < 1 || cb < 0 || cout < 1) {
    set_error("conv1x1_chain: invalid shape");
    return SBMC_EINVAL;
  }
  if (n_img == 0 || hw == 0) return SBMC_OK;
  if (!xa || (cb > 0 && !xb) || !w1 || !w2 || !w3 || !b1 || !b2 || !b3 || !y) {
    set_error("conv1x1_chain: null pointer argument");
    return SBMC_EINVAL;
  }
  if ((k1p != 128 && k1p != 256) || ca + cb > k1p || n3p % 16 != 0 || n3p < cout ||
      n3p > 512 || n3p < 16) {
    set_error("conv1x1_chain: unsupported sizes cin=%d k1p=%d cout=%d n3p=%d", ca + cb, k1p,
              cout, n3p);
    return SBMC_EUNSUPPORTED;
  }
  ChainArgs a;
  a.xa = xa; a.xb = cb > 0 ? xb : nullptr;
  a.a_img = a_img_stride; a.b_img = b_img_stride;
  a.ca = ca; a.cb = cb; a.b_bcast = b_broadcast ? 1 : 0;
  a.b1 = b1; a.b2 = b2; a.b3 = b3;
  a.y = y; a.y_img = y_img_stride;
  a.cout = cout; a.n3p = n3p; a.act = act ? 1 : 0;
  a.hw = hw;
  a.tiles_per_img = (hw + kTileP - 1) / kTileP;
  a.ntiles = a.tiles_per_img * n_img;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  note_path(1);
  if (k1p == 128) return run_chain<128>(a, w1, w2, w3, st);
  return run_chain<256>(a, w1, w2, w3, st);
}


extern "C" int sbmc_conv1x1_chain_nhwc_bf16(const void *xa, int64_t a_img_stride, const void *xb,
                                            int64_t b_img_stride, const void *w1,
                                            const float *b1, int64_t b1_img_stride,
                                            const void *w2, const float *b2, const void *w3,
                                            const float *b3, int cout, int n3p, int act,
                                            void *y, int64_t y_img_stride, int out_nhwc_bf16,
                                            int64_t n_img, int64_t hw, void *stream) {
  using namespace sbmc;
  if (n_img < 0 || hw < 0 || cout < 1) {
    set_error("conv1x1_chain_nhwc: invalid shape");
    return SBMC_EINVAL;
  }
  if (n_img == 0 || hw == 0) return SBMC_OK;
  if (!xa || !w1 || !w2 || !w3 || !b1 || !b2 || !b3 || !y) {
    set_error("conv1x1_chain_nhwc: null pointer argument");
    return SBMC_EINVAL;
  }
  if (n3p % 16 != 0 || n3p < cout || n3p > 512 || n3p < 16 || (out_nhwc_bf16 && cout != 128) ||
      hw >= (1ll << 31) || n_img >= (1ll << 31)) {
    set_error("conv1x1_chain_nhwc: unsupported sizes cout=%d n3p=%d", cout, n3p);
    return SBMC_EUNSUPPORTED;
  }
  ChainArgsV2 a;
  a.b1 = b1; a.b1_img = b1_img_stride; a.b2 = b2; a.b3 = b3;
  a.y = y; a.y_img = y_img_stride; a.out_nhwc_bf16 = out_nhwc_bf16 ? 1 : 0;
  a.cout = cout; a.n3p = n3p; a.act = act ? 1 : 0;
  a.hw = hw;
  a.tiles_per_img = (hw + kTileP - 1) / kTileP;
  a.ntiles = a.tiles_per_img * n_img;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  note_path(1);
  if (!xb) return run_chain_v2<2>(a, xa, a_img_stride, nullptr, 0, n_img, w1, w2, w3, st);
  return run_chain_v2<4>(a, xa, a_img_stride, xb, b_img_stride, n_img, w1, w2, w3, st);
}
