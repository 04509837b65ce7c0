// chain_v3.cu -- warp-specialised, software-pipelined per-sample 1x1 ConvChain on
// tcgen05 (round 2; supersedes the serial conv1x1_chain_nhwc_kernel on the
// inference path).
//
// Reference computation (sbmc/models.py:143-181,195-199; sbmc/modules.py:34-125):
//   embedding_XX     : f_s' = chain(cat(f_s, prop | gf)),   reduced = mean_s f_s'
//   kernel_regressor : logits_s = chain(cat(f_s, prop))
// a 3-layer per-pixel MLP (hidden width 128, ReLU / LeakyReLU), applied to every
// sample s of every pixel.
//
// One CTA per SM, persistent over 128-pixel tiles.  An ITEM is (tile, sample pair):
// the two samples of a pair are two independent streams that run the chain in
// lock step, so while the epilogue warps of one stream convert an accumulator the
// tensor pipe executes the other stream's layer (ping-pong), and both streams
// share what the pair has in common -- the `prop` operand of the tile and every
// weight chunk:
//
//   warp 0      TMA producer: F_a, F_b (the pair's sample features), P (the tile's
//               propagated features, loaded once per tile for all samples)
//   warp 1      MMA issuer (one thread): layer 1 reads F / P from shared memory
//               (SS), layers 2 and 3 read the hidden activations FROM TENSOR MEMORY
//               (tcgen05.mma with the A operand in TMEM): the activations never
//               touch shared memory, which is what leaves room for two streams
//   warp 2      TMA producer of the last layer's weight chunks (regressor only:
//               441 x 128 bf16 does not fit beside W1 / W2, it streams from L2
//               through a two-deep ring, one chunk serving both streams)
//   warp 3      TMEM allocation
//   warps 4-11  epilogue group of stream a, warps 12-19 of stream b: tcgen05.ld of
//               the fp32 accumulator -> bias, activation -> packed bf16 ->
//               tcgen05.st as the next layer's A operand; last layer: bf16 NHWC
//               store (embedding) or fp32 NCHW store (logits)
//
// TMEM (512 columns): per stream X (128, accumulator) + Y (64, bf16 A operand);
// the remaining 128 columns hold the running SUM OVER SAMPLES of the embedding's
// last layer -- every sample's last-layer MMAs are issued a second time into that
// accumulator, so the reference's `mean over spp` (models.py:181) costs no
// epilogue work and is taken in fp32 before any bf16 rounding -- or, for the
// regressor, a third 64-column chunk accumulator per stream.
#include <cuda_bf16.h>

#include "umma.cuh"

namespace sbmc {

namespace v3 {

constexpr int kHid = 128;
constexpr int kTileP = 128;
constexpr int kSlab = 128 * 128;          // 128 rows x 64 bf16
constexpr int kCtrlWarps = 4;
constexpr int kEgWarps = 8;               // epilogue warps per stream
constexpr int kThreads = 32 * (kCtrlWarps + 2 * kEgWarps);
constexpr int kChunk = 64;                // output channels per last-layer chunk (regressor)
#ifdef SBMC_CHAIN_NOSTORE                 // developer experiment: no output stores
constexpr bool kStoreOn = false;
#else
constexpr bool kStoreOn = true;
#endif

struct Args {
  const float *b1; long long b1_img;      // first-layer bias, optionally per image
  const float *b2, *b3;
  void *out; long long out_img, out_smp;  // elements between images / samples
  void *mean; long long mean_img;         // embedding: mean over the samples (or null)
  int mean_f32; float inv_spp;
  int cout, n3p;
  int s0, ns;                             // samples [s0, s0 + ns) of the feature tensor
  int hw, tiles_per_img;
  long long ntiles;
  long long *trace;                       // developer timeline (SBMC_CHAIN_TRACE builds), or null
};

// Developer timeline: four threads of CTA 0 (producer, MMA issuer, lane 0 of the first
// warp of each epilogue group) append (code, clock) pairs to their own region with
// plain stores (no atomics: a returning atomic would sit on the critical path).
// Compiled out by default.
#ifdef SBMC_CHAIN_TRACE
#define TRACE_DECL(role) int trace_n_ = 0; const int trace_role_ = (role)
#define TRACE(code)                                                                  \
  do {                                                                               \
    if (P.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_n_ < 2000) {  \
      P.trace[(trace_role_ * 2000 + trace_n_) * 2] = (code);                         \
      P.trace[(trace_role_ * 2000 + trace_n_) * 2 + 1] = clock64();                  \
      ++trace_n_;                                                                    \
    }                                                                                \
  } while (0)
#else
#define TRACE_DECL(role) do { } while (0)
#define TRACE(code) do { } while (0)
#endif

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&v);
}
template <int LEAKY>
__device__ __forceinline__ float activate(float v) {
  return LEAKY ? fmaxf(v, 0.01f * v) : fmaxf(v, 0.f);
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void eg_barrier(int id) {      // the 256 threads of one group
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kEgWarps * 32) : "memory");
}
__device__ __forceinline__ void stg128(void *p, const uint4 &v) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg32_stream(float *p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Barrier slots.
enum {
  B_W = 0, B_F_FULL, B_F_EMPTY, B_P_FULL, B_P_EMPTY,
  B_ACC0, B_ACC1, B_AR0, B_AR1, B_M_FULL, B_M_EMPTY,
  B_W3F0, B_W3F1, B_W3E0, B_W3E1,
  B_OF = 15,        // o_full[e][3]
  B_OE = 21,        // o_empty[e][3]
  B_COUNT = 27
};

// Hidden-layer epilogue of one epilogue thread: its row's columns
// [64 half, 64 half + 64) of X -> bias, activation, bf16 -> Y columns [32 half, +32).
template <int LEAKY>
__device__ __forceinline__ void hidden_epilogue(uint32_t x, uint32_t y, const float *sbias,
                                                int half) {
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int c0 = half * 64 + it * 32;
    float v[32];
    tmem_ld_32x32b_x32(x + c0, v);
    uint32_t q[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = *reinterpret_cast<const float4 *>(sbias + c0 + 4 * j);
      q[2 * j] = pack_bf16(activate<LEAKY>(v[4 * j] + b.x), activate<LEAKY>(v[4 * j + 1] + b.y));
      q[2 * j + 1] =
          pack_bf16(activate<LEAKY>(v[4 * j + 2] + b.z), activate<LEAKY>(v[4 * j + 3] + b.w));
    }
    tmem_st_32x32b_x16(y + (c0 >> 1), q);
  }
  tmem_wait_st();
}

// KS1: 64-channel slabs of the first layer's K (2: features only, 4: features + prop).
// REGRESS: 0 embedding (bf16 NHWC out + mean), 1 regressor (fp32 NCHW logits).
template <int KS1, int REGRESS, int LEAKY>
__global__ void __launch_bounds__(kThreads, 1)
chain_v3_kernel(const __grid_constant__ CUtensorMap fmap,      // feats {128, hw, spp, n}
                const __grid_constant__ CUtensorMap pmap,      // prop  {128, hw, n}
                const __grid_constant__ CUtensorMap w1map,
                const __grid_constant__ CUtensorMap w2map,
                const __grid_constant__ CUtensorMap w3map, const Args P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sW1 = smem;
  unsigned char *sW2 = sW1 + KS1 * kSlab;
  unsigned char *sW3 = sW2 + 2 * kSlab;             // embedding: resident; regressor: 2 x 16 KB ring
  unsigned char *sF = sW3 + 2 * kSlab;              // F_a | F_b, 2 slabs each
  unsigned char *sP = sF + 4 * kSlab;               // 2 slabs (KS1 == 4)
  float *sB1 = reinterpret_cast<float *>(sP + (KS1 == 4 ? 2 : 0) * kSlab);   // [2][128]
  float *sB2 = sB1 + 2 * kHid;
  float *sB3 = sB2 + kHid;                          // embedding only
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB3 + (REGRESS ? 0 : kHid));
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
    for (int i = 0; i < B_COUNT; ++i) {
      const bool eg = (i == B_AR0 || i == B_AR1 || i == B_M_EMPTY || i >= B_OE);
      mbar_init(bars + i, eg ? kEgWarps : 1);
    }
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc(tmem_slot, 512);
  if (tid < kHid) {
    sB2[tid] = P.b2[tid];
    if (!REGRESS) sB3[tid] = P.b3[tid];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int npairs = (P.ns + 1) >> 1;
  const int nchunks = (P.n3p + kChunk - 1) / kChunk;
  const bool do_mean = !REGRESS && P.mean != nullptr;

  if (warp == 0) {
    // ===================== TMA producer: weights, F, P =====================
    if (lane == 0) {
      const uint32_t wbytes = (uint32_t)((KS1 + 2 + (REGRESS ? 0 : 2)) * kSlab);
      mbar_expect_tx(bars + B_W, wbytes);
      for (int kb = 0; kb < KS1; ++kb) tma_load_2d(sW1 + kb * kSlab, &w1map, bars + B_W, kb * 64, 0);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(sW2 + kb * kSlab, &w2map, bars + B_W, kb * 64, 0);
      if (!REGRESS)
        for (int kb = 0; kb < 2; ++kb) tma_load_2d(sW3 + kb * kSlab, &w3map, bars + B_W, kb * 64, 0);
      uint32_t ph_fe = 0, ph_pe = 0;
      TRACE_DECL(0);
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int n = (int)(tile / P.tiles_per_img);
        const int p0 = (int)(tile - (long long)n * P.tiles_per_img) * kTileP;
        for (int j = 0; j < npairs; ++j) {
          const bool bvalid = 2 * j + 1 < P.ns;
          mbar_wait(bars + B_F_EMPTY, ph_fe ^ 1); ph_fe ^= 1;
          TRACE(1);                                   // F slot free, issuing the loads
          mbar_expect_tx(bars + B_F_FULL, (uint32_t)((bvalid ? 4 : 2) * kSlab));
          for (int e = 0; e < (bvalid ? 2 : 1); ++e)
            for (int kb = 0; kb < 2; ++kb)
              tma_load_4d(sF + (2 * e + kb) * kSlab, &fmap, bars + B_F_FULL, kb * 64, p0,
                          P.s0 + 2 * j + e, n);
          if (KS1 == 4 && j == 0) {
            mbar_wait(bars + B_P_EMPTY, ph_pe ^ 1); ph_pe ^= 1;
            mbar_expect_tx(bars + B_P_FULL, (uint32_t)(2 * kSlab));
            for (int kb = 0; kb < 2; ++kb)
              tma_load_3d(sP + kb * kSlab, &pmap, bars + B_P_FULL, kb * 64, p0, n);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== TMA producer: last-layer weight chunks =====================
    if (REGRESS && lane == 0) {
      uint32_t ph_e = 0;                         // bit `ring`
      int ring = 0;
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x)
        for (int j = 0; j < npairs; ++j)
          for (int c = 0; c < nchunks; ++c) {
            // (a box that hangs over the last row is zero-filled and still counts in full)
            mbar_wait(bars + B_W3E0 + ring, ((ph_e >> ring) & 1) ^ 1); ph_e ^= 1u << ring;
            mbar_expect_tx(bars + B_W3F0 + ring, (uint32_t)(2 * min(kChunk, P.n3p) * 128));
            for (int kb = 0; kb < 2; ++kb)
              tma_load_2d(sW3 + ring * kSlab + kb * (kSlab / 2), &w3map, bars + B_W3F0 + ring,
                          kb * 64, c * kChunk);
            ring ^= 1;
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (uniform) control flow and waits on the barriers; one
    // elected lane issues the MMAs and the commits (see umma.cuh::elect_one).
    {
      const uint32_t idesc_h = umma_idesc_bf16(128, kHid);
      // descriptors of slab 0 of every operand; slab s is + s * (kSlab >> 4), K step k
      // inside a slab is + 2 k (16-byte units)
      const uint64_t dF = umma_smem_desc_sw128(sF), dP = umma_smem_desc_sw128(sP);
      const uint64_t dW1 = umma_smem_desc_sw128(sW1), dW2 = umma_smem_desc_sw128(sW2);
      const uint64_t dW3 = umma_smem_desc_sw128(sW3);
      constexpr uint64_t kSlabD = kSlab >> 4;
      // phase bits live in one register each (no dynamically indexed local arrays)
      uint32_t ph_f = 0, ph_p = 0, ph_me = 0;
      uint32_t ph_ar = 0;                        // bit e
      uint32_t ph_w3f = 0;                       // bit ring
      uint32_t ph_oe = 0;                        // bit 3 e + ob
      uint32_t has_prev = 0;                     // bit e
      int ring = 0;
      TRACE_DECL(1);
      auto wait_ar = [&](int e) {
        mbar_wait(bars + B_AR0 + e, (ph_ar >> e) & 1);
        ph_ar ^= 1u << e;
      };
      mbar_wait(bars + B_W, 0);
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        for (int j = 0; j < npairs; ++j) {
          const int ne = (2 * j + 1 < P.ns) ? 2 : 1;
          const bool last = j == npairs - 1;
          // ---- layer 1: X_e = [F_e | P] . W1^T  (operands in shared memory) ----
          TRACE(10);
          mbar_wait(bars + B_F_FULL, ph_f); ph_f ^= 1;
          if (KS1 == 4 && j == 0) { mbar_wait(bars + B_P_FULL, ph_p); ph_p ^= 1; }
          TRACE(11);                                  // F (and P) landed
          for (int e = 0; e < ne; ++e) {
            if ((has_prev >> e) & 1) wait_ar(e);      // X_e drained by the previous item
            has_prev |= 1u << e;
            TRACE(12 + e);                            // X_e free, issuing layer 1
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t x = tmem + e * 192;
              const uint64_t dFe = dF + (uint64_t)(2 * e) * kSlabD;
#pragma unroll
              for (int k = 0; k < KS1 * 4; ++k) {
                const uint64_t ad = (k < 8 ? dFe + (uint64_t)(k >> 2) * kSlabD
                                           : dP + (uint64_t)((k - 8) >> 2) * kSlabD) +
                                    (uint64_t)((k & 3) * 2);
                const uint64_t bd = dW1 + (uint64_t)(k >> 2) * kSlabD + (uint64_t)((k & 3) * 2);
                umma_bf16(x, ad, bd, idesc_h, k > 0);
              }
              umma_commit(bars + B_ACC0 + e);
              if (e == ne - 1) {
                umma_commit(bars + B_F_EMPTY);
                if (KS1 == 4 && last) umma_commit(bars + B_P_EMPTY);
              }
            }
            __syncwarp();
          }
          // ---- layer 2: X_e = Y_e . W2^T  (A operand in tensor memory) ----
          for (int e = 0; e < ne; ++e) {
            wait_ar(e);
            TRACE(14 + e);                            // E1_e done, issuing layer 2
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t x = tmem + e * 192, y = x + 128;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16_ts(x, y + k * 8, dW2 + (uint64_t)(k >> 2) * kSlabD + (uint64_t)((k & 3) * 2),
                             idesc_h, k > 0);
              umma_commit(bars + B_ACC0 + e);
            }
            __syncwarp();
          }
          // ---- layer 3 ----
          if (!REGRESS) {
            if (do_mean && j == 0) { mbar_wait(bars + B_M_EMPTY, ph_me ^ 1); ph_me ^= 1; }
            for (int e = 0; e < ne; ++e) {
              wait_ar(e);
              TRACE(16 + e);                          // E2_e done, issuing layer 3
              tcgen05_fence_after();
              if (elect_one()) {
                const uint32_t x = tmem + e * 192, y = x + 128;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const uint64_t bd = dW3 + (uint64_t)(k >> 2) * kSlabD + (uint64_t)((k & 3) * 2);
                  umma_bf16_ts(x, y + k * 8, bd, idesc_h, k > 0);
                  if (do_mean) umma_bf16_ts(tmem + 384, y + k * 8, bd, idesc_h, (j | e | k) > 0);
                }
                umma_commit(bars + B_ACC0 + e);
                if (do_mean && last && e == ne - 1) umma_commit(bars + B_M_FULL);
              }
              __syncwarp();
            }
          } else {
            for (int c = 0; c < nchunks; ++c) {
              const int rows = min(kChunk, P.n3p - c * kChunk);
              const uint32_t idesc_c = umma_idesc_bf16(128, rows);
              const int ob = c % 3;
              mbar_wait(bars + B_W3F0 + ring, (ph_w3f >> ring) & 1); ph_w3f ^= 1u << ring;
              for (int e = 0; e < ne; ++e) {
                if (c == 0) wait_ar(e);
                if (c >= 3) {
                  mbar_wait(bars + B_OE + 3 * e + ob, (ph_oe >> (3 * e + ob)) & 1);
                  ph_oe ^= 1u << (3 * e + ob);
                }
                tcgen05_fence_after();
                if (elect_one()) {
                  const uint32_t y = tmem + e * 192 + 128;
                  const uint32_t o = (ob < 2) ? tmem + e * 192 + ob * 64 : tmem + 384 + e * 64;
                  const uint64_t wb = dW3 + (uint64_t)ring * kSlabD;
#pragma unroll
                  for (int k = 0; k < 8; ++k)
                    umma_bf16_ts(o, y + k * 8,
                                 wb + (uint64_t)(k >> 2) * (kSlabD / 2) + (uint64_t)((k & 3) * 2),
                                 idesc_c, k > 0);
                  umma_commit(bars + B_OF + 3 * e + ob);
                  if (e == ne - 1) umma_commit(bars + B_W3E0 + ring);
                }
                __syncwarp();
              }
              ring ^= 1;
            }
          }
        }
      }
      // drain: every MMA and every commit-arrive has landed before the CTA may exit
      if (elect_one()) umma_commit(bars + B_W);
      __syncwarp();
      mbar_wait(bars + B_W, 1);
    }
  } else if (warp >= kCtrlWarps) {
    // ===================== epilogue groups =====================
    const int e = (warp - kCtrlWarps) / kEgWarps;             // stream
    const int wi = (warp - kCtrlWarps) % kEgWarps;
    const int quad = warp & 3, half = wi >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t x = lane_base + e * 192, y = x + 128;
    float *sb1 = sB1 + e * kHid;
    uint32_t ph_acc = 0, ph_m = 0, ph_of = 0;       // ph_of: bit ob
    TRACE_DECL(2 + e);
    long long cur_n = -1;
    for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const long long n = tile / P.tiles_per_img;
      const long long p = (tile - n * P.tiles_per_img) * kTileP + row;
      const bool valid = kStoreOn && p < P.hw;
      for (int j = 0; j < npairs; ++j) {
        const int sl = 2 * j + e;                    // sample index within this launch
        if (sl >= P.ns) continue;
        if (n != cur_n) {                            // (per-image) first-layer bias
          eg_barrier(1 + e);
          const int t = wi * 32 + lane;
          if (t < kHid) sb1[t] = P.b1[n * P.b1_img + t];
          eg_barrier(1 + e);
          cur_n = n;
        }
        // ---- layers 1 and 2: accumulator -> next layer's A operand in TMEM ----
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          mbar_wait(bars + B_ACC0 + e, ph_acc); ph_acc ^= 1;
          if (wi == 0 && lane == 0) TRACE(20 + 10 * e + 2 * layer);       // accumulator ready
          tcgen05_fence_after();
          hidden_epilogue<LEAKY>(x, y, layer ? sB2 : sb1, half);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + B_AR0 + e);
          if (wi == 0 && lane == 0) TRACE(21 + 10 * e + 2 * layer);       // epilogue done
        }
        if (!REGRESS) {
          // ---- embedding output: bf16, channels innermost ----
          mbar_wait(bars + B_ACC0 + e, ph_acc); ph_acc ^= 1;
          if (wi == 0 && lane == 0) TRACE(24 + 10 * e);
          tcgen05_fence_after();
          // this thread's 64 channels -> 32 packed words -> quad transpose -> 4 x 256-bit
          // stores, each completing 8 lines (see umma.cuh::quad_transpose32)
          {
            uint32_t q[32];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int c0 = half * 64 + it * 32;
              float v[32];
              tmem_ld_32x32b_x32(x + c0, v);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b = *reinterpret_cast<const float4 *>(sB3 + c0 + 4 * g);
                q[16 * it + 2 * g] = pack_bf16(v[4 * g] + b.x, v[4 * g + 1] + b.y);
                q[16 * it + 2 * g + 1] = pack_bf16(v[4 * g + 2] + b.z, v[4 * g + 3] + b.w);
              }
            }
            quad_transpose32(q, lane);
            const long long prow = p - (lane & 3);             // first pixel of the quad
            __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(P.out) + n * P.out_img +
                                 (long long)(P.s0 + sl) * P.out_smp + prow * kHid + half * 64 +
                                 (lane & 3) * 16;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (kStoreOn && prow + k < P.hw) stg256(dst + k * kHid, q + 8 * k);
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + B_AR0 + e);
          if (wi == 0 && lane == 0) TRACE(25 + 10 * e);
          // ---- mean over the samples: read once per tile by the last stream ----
          if (do_mean && sl == P.ns - 1) {
            mbar_wait(bars + B_M_FULL, ph_m); ph_m ^= 1;
            tcgen05_fence_after();
            uint32_t mq[32];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int c0 = half * 64 + it * 32;
              float v[32];
              tmem_ld_32x32b_x32(lane_base + 384 + c0, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], P.inv_spp, sB3[c0 + i]);
              {
                if (P.mean_f32 && valid) {
                  float *m = reinterpret_cast<float *>(P.mean) + n * P.mean_img + p * kHid + c0;
#pragma unroll
                  for (int g = 0; g < 8; ++g) {
                    uint4 q;
                    q.x = __float_as_uint(v[4 * g]); q.y = __float_as_uint(v[4 * g + 1]);
                    q.z = __float_as_uint(v[4 * g + 2]); q.w = __float_as_uint(v[4 * g + 3]);
                    stg128(m + 4 * g, q);
                  }
                }
              }
              if (!P.mean_f32) {
#pragma unroll
                for (int g = 0; g < 16; ++g) mq[16 * it + g] = pack_bf16(v[2 * g], v[2 * g + 1]);
              }
            }
            if (!P.mean_f32) {
              quad_transpose32(mq, lane);
              const long long prow = p - (lane & 3);
              __nv_bfloat16 *m = reinterpret_cast<__nv_bfloat16 *>(P.mean) + n * P.mean_img +
                                 prow * kHid + half * 64 + (lane & 3) * 16;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (kStoreOn && prow + k < P.hw) stg256(m + k * kHid, mq + 8 * k);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + B_M_EMPTY);
          }
        } else {
          // ---- logits: fp32 NCHW, a warp stores 32 consecutive pixels of one channel ----
          float *yp = reinterpret_cast<float *>(P.out) + n * P.out_img +
                      (long long)(P.s0 + sl) * P.out_smp + p;
#pragma unroll 1
          for (int c = 0; c < nchunks; ++c) {
            const int ob = c % 3;
            const int rows = min(kChunk, P.n3p - c * kChunk);
            mbar_wait(bars + B_OF + 3 * e + ob, (ph_of >> ob) & 1); ph_of ^= 1u << ob;
            tcgen05_fence_after();
            const uint32_t o = (ob < 2) ? x + ob * 64 : lane_base + 384 + e * 64;
            const int col = c * kChunk + half * 32;
            if (half * 32 < rows) {                   // warp-uniform
              float v[32];
              tmem_ld_32x32b_x32(o + half * 32, v);
#pragma unroll
              for (int g = 0; g < 8; ++g) {           // + bias (uniform 128-bit loads)
                const float4 b = __ldg(reinterpret_cast<const float4 *>(P.b3 + col + 4 * g));
                v[4 * g + 0] += b.x; v[4 * g + 1] += b.y;
                v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
              }
              if (valid) {
                // one running pointer, one add per store; the channel bound is tested
                // per 32-channel block (warp-uniform), not per store
                float *ptr = yp + (long long)col * P.hw;
                if (col + 32 <= P.cout) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    stg32_stream(ptr, v[i]);
                    ptr += P.hw;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    if (col + i < P.cout) stg32_stream(ptr, v[i]);
                    ptr += P.hw;
                  }
                }
              }
            }
            if (c + 3 < nchunks) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bars + B_OE + 3 * e + ob);
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + B_AR0 + e);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}

template <int KS1, int REGRESS, int LEAKY>
static int launch(const Args &a, const CUtensorMap &fm, const CUtensorMap &pm,
                  const CUtensorMap &m1, const CUtensorMap &m2, const CUtensorMap &m3,
                  cudaStream_t st) {
  const size_t smem = (size_t)(KS1 + 2 + 2 + 4 + (KS1 == 4 ? 2 : 0)) * kSlab +
                      (size_t)(3 * kHid + (REGRESS ? 0 : kHid)) * sizeof(float) +
                      B_COUNT * sizeof(uint64_t) + 16;
  auto kern = chain_v3_kernel<KS1, REGRESS, LEAKY>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  {
    KernelTimer timer(SBMC_KERNEL_CONV1X1, st);
    kern<<<(unsigned)grid, kThreads, smem, st>>>(fm, pm, m1, m2, m3, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace v3
}  // namespace sbmc

#ifdef SBMC_CHAIN_TRACE
extern "C" __attribute__((visibility("default"))) void *sbmc_b200_debug_pointer = nullptr;
#endif

// Public entry point (include/sbmc_b200.h).
extern "C" int sbmc_chain_samples_nhwc_bf16(
    const void *feats, int64_t f_img_stride, int64_t f_smp_stride, int64_t spp_total,
    const void *prop, int64_t p_img_stride, const void *w1, const float *b1, int64_t b1_img_stride,
    const void *w2, const float *b2, const void *w3, const float *b3, int cout, int n3p, int act,
    int regress, void *out, int64_t out_img_stride, int64_t out_smp_stride, void *mean,
    int64_t mean_img_stride, int mean_f32, int64_t n_img, int64_t sample0, int64_t nsamples,
    int64_t hw, void *stream) {
  using namespace sbmc;
  if (n_img < 0 || hw < 0 || cout < 1 || nsamples < 0 || sample0 < 0 ||
      sample0 + nsamples > spp_total) {
    set_error("chain_samples: invalid shape");
    return SBMC_EINVAL;
  }
  if (n_img == 0 || hw == 0 || nsamples == 0) return SBMC_OK;
  if (!feats || !w1 || !w2 || !w3 || !b1 || !b2 || !b3 || !out) {
    set_error("chain_samples: null pointer argument");
    return SBMC_EINVAL;
  }
  if (n3p % 16 != 0 || n3p < cout || n3p > 512 || n3p < 16 || (!regress && (cout != 128 || n3p != 128)) ||
      (regress && mean) || hw >= (1ll << 31) || n_img >= (1ll << 31) || spp_total >= (1ll << 31)) {
    set_error("chain_samples: unsupported sizes cout=%d n3p=%d regress=%d", cout, n3p, regress);
    return SBMC_EUNSUPPORTED;
  }
  v3::Args a;
  a.b1 = b1; a.b1_img = b1_img_stride; a.b2 = b2; a.b3 = b3;
  a.out = out; a.out_img = out_img_stride; a.out_smp = out_smp_stride;
  a.mean = mean; a.mean_img = mean_img_stride; a.mean_f32 = mean_f32 ? 1 : 0;
  a.inv_spp = 1.0f / (float)nsamples;
  a.cout = cout; a.n3p = n3p;
  a.s0 = (int)sample0; a.ns = (int)nsamples;
  a.hw = (int)hw;
  a.tiles_per_img = (int)((hw + v3::kTileP - 1) / v3::kTileP);
  a.ntiles = (long long)a.tiles_per_img * n_img;
  a.trace = nullptr;
#ifdef SBMC_CHAIN_TRACE
  a.trace = static_cast<long long *>(sbmc_b200_debug_pointer);
#endif
  CUtensorMap fm, pm, m1, m2, m3;
  {
    const uint64_t dims[4] = {128, (uint64_t)hw, (uint64_t)spp_total, (uint64_t)n_img};
    const uint64_t str[3] = {256, (uint64_t)f_smp_stride * 2, (uint64_t)f_img_stride * 2};
    const uint32_t box[4] = {64, v3::kTileP, 1, 1};
    if (!encode_tensor_map_bf16_sw128(&fm, feats, 4, dims, str, box)) return SBMC_ECUDA;
  }
  pm = fm;
  if (prop && !encode_tensor_map_bf16_3d_sw128(&pm, prop, 128, (uint64_t)hw, (uint64_t)n_img,
                                               (uint64_t)p_img_stride, v3::kTileP))
    return SBMC_ECUDA;
  const int ks1 = prop ? 4 : 2;
  const uint32_t w3rows = regress ? (uint32_t)(n3p < v3::kChunk ? n3p : v3::kChunk) : 128u;
  if (!encode_tensor_map_bf16_2d_sw128(&m1, w1, ks1 * 64, 128, 64, 128) ||
      !encode_tensor_map_bf16_2d_sw128(&m2, w2, 128, 128, 64, 128) ||
      !encode_tensor_map_bf16_2d_sw128(&m3, w3, 128, (uint64_t)n3p, 64, w3rows))
    return SBMC_ECUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  note_path(1);
  const int key = (ks1 == 4 ? 4 : 0) | (regress ? 2 : 0) | (act ? 1 : 0);
  switch (key) {
    case 0: return v3::launch<2, 0, 0>(a, fm, pm, m1, m2, m3, st);
    case 1: return v3::launch<2, 0, 1>(a, fm, pm, m1, m2, m3, st);
    case 2: return v3::launch<2, 1, 0>(a, fm, pm, m1, m2, m3, st);
    case 3: return v3::launch<2, 1, 1>(a, fm, pm, m1, m2, m3, st);
    case 4: return v3::launch<4, 0, 0>(a, fm, pm, m1, m2, m3, st);
    case 5: return v3::launch<4, 0, 1>(a, fm, pm, m1, m2, m3, st);
    case 6: return v3::launch<4, 1, 0>(a, fm, pm, m1, m2, m3, st);
    default: return v3::launch<4, 1, 1>(a, fm, pm, m1, m2, m3, st);
  }
}
