// linear.cu -- y[p][co] = act(sum_ci x[p][ci] * w[co][ci] + bias[co]) on bf16
// channels-innermost activations: one 1x1-convolution layer as a tcgen05 GEMM.
//
// Used by the opt-in mixed-precision TRAINING path of the per-sample 1x1 ConvChains
// (sbmc/modules.py:34-125; embedding_XX and kernel_regressor of sbmc/models.py:86-102):
// in training every layer's output has to exist in memory for the backward pass, so the
// chain runs layer by layer (forward: three launches; data gradient: three launches on
// the transposed weights; weight gradients: library GEMMs).  Inference uses the fused,
// pipelined chain kernel instead (csrc/chain_v3.cu).
//
// Same machinery as csrc/conv3x3.cu with one tap and no halo: a CTA owns 256
// consecutive pixels (two M = 128 row blocks) x NT in {128, 256} output channels; per
// 64-channel slab one TMA box [256 px x 64 ch] (A) and one [NT x 64] (B) through
// mbarrier rings; fp32 accumulators in TMEM (double-buffered for NT = 128); the MMA warp
// issues through an elected lane; 16 epilogue warps add the bias, apply ReLU /
// LeakyReLU and store bf16 or fp32 rows with 256-bit stores.
#include <cuda_bf16.h>

#include "umma.cuh"

namespace sbmc {
namespace lin {

constexpr int kTilePx = 256;
constexpr int kASlab = kTilePx * 128;              // 32 KB
constexpr int kAStages = 3;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (4 + kEpiWarps);
__host__ __device__ constexpr int b_stages(int nt) { return nt == 128 ? 6 : 3; }

struct Args {
  const float *bias;       // may be null
  void *out;
  int act;                 // 0 none, 1 ReLU, 2 LeakyReLU(0.01)
  int out_f32;
  long long P;             // pixels (rows)
  int Cin, Cout;
  long long tiles_p;
  long long ntiles;
};

enum { B_AF = 0, B_AE = 3, B_BF = 6, B_BE = 12, B_ACCF = 18, B_ACCE = 20, B_COUNT = 22 };

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(kThreads, 1)
linear_kernel(const __grid_constant__ CUtensorMap amap,      // x {Cin, P}
              const __grid_constant__ CUtensorMap wmap,      // w {Cin, Cout}
              const Args P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int kBStages = b_stages(NT);
  constexpr int kBStage = NT * 128;
  unsigned char *sA = smem;
  unsigned char *sB = smem + kAStages * kASlab;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + kBStages * kBStage);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
    if (lane == 0) {            // A producer
      uint32_t ph = 0;
      int ab = 0;
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const long long p0 = (tile % P.tiles_p) * kTilePx;
        for (int s = 0; s < nslabs; ++s) {
          mbar_wait(bars + B_AE + ab, ((ph >> ab) & 1) ^ 1); ph ^= 1u << ab;
          mbar_expect_tx(bars + B_AF + ab, (uint32_t)kASlab);
          tma_load_2d(sA + ab * kASlab, &amap, bars + B_AF + ab, s * 64, (int)p0);
          ab = (ab + 1 == kAStages) ? 0 : ab + 1;
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {            // B producer
      uint32_t ph = 0;
      int st = 0;
      for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int n0 = (int)(tile / P.tiles_p) * NT;
        for (int s = 0; s < nslabs; ++s) {
          mbar_wait(bars + B_BE + st, ((ph >> st) & 1) ^ 1); ph ^= 1u << st;
          mbar_expect_tx(bars + B_BF + st, (uint32_t)kBStage);
          tma_load_2d(sB + st * kBStage, &wmap, bars + B_BF + st, s * 64, n0);
          st = (st + 1 == kBStages) ? 0 : st + 1;
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp, elected lane issues (umma.cuh::elect_one)
    const uint32_t idesc = umma_idesc_bf16(128, NT);
    const uint64_t dA = umma_smem_desc_sw128(sA), dB = umma_smem_desc_sw128(sB);
    uint32_t ph_a = 0, ph_b = 0, ph_acc = 0;
    int ab = 0, st = 0, it = 0;
    for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const int buf = (NT == 128) ? (it & 1) : 0;
      mbar_wait(bars + B_ACCE + buf, ((ph_acc >> buf) & 1) ^ 1); ph_acc ^= 1u << buf;
      for (int s = 0; s < nslabs; ++s) {
        mbar_wait(bars + B_AF + ab, (ph_a >> ab) & 1); ph_a ^= 1u << ab;
        mbar_wait(bars + B_BF + st, (ph_b >> st) & 1); ph_b ^= 1u << st;
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t da = dA + (uint64_t)ab * (kASlab >> 4);
          const uint64_t bd0 = dB + (uint64_t)st * (kBStage >> 4);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint64_t ad0 = da + (uint64_t)(g * 128 * 8);
            const uint32_t d = tmem + ((NT == 128) ? buf * 256 + g * 128 : g * 256);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc, (s | k) > 0);
          }
          umma_commit(bars + B_BE + st);
          umma_commit(bars + B_AE + ab);
          if (s == nslabs - 1) umma_commit(bars + B_ACCF + buf);
        }
        __syncwarp();
        ab = (ab + 1 == kAStages) ? 0 : ab + 1;
        st = (st + 1 == kBStages) ? 0 : st + 1;
      }
    }
  } else if (warp >= 4) {
    // epilogue: warp -> (lane quadrant, row block g, column half)
    const int quad = warp & 3, g = ((warp - 4) >> 2) & 1, part = (warp - 4) >> 3;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const long long p = (tile % P.tiles_p) * kTilePx + g * 128 + quad * 32 + lane;
      const int n0 = (int)(tile / P.tiles_p) * NT;
      const int buf = (NT == 128) ? (it & 1) : 0;
      mbar_wait(bars + B_ACCF + buf, (ph >> buf) & 1); ph ^= 1u << buf;
      tcgen05_fence_after();
      const bool valid = p < P.P;
#pragma unroll 1
      for (int c0 = part * (NT / 2); c0 < (part + 1) * (NT / 2); c0 += 32) {
        const uint32_t col = (NT == 128) ? buf * 256 + g * 128 + c0 : g * 256 + c0;
        float v[32];
        tmem_ld_32x32b_x32(lane_base + col, v);
        if (P.bias) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(P.bias + n0 + c0 + 4 * q4));
            v[4 * q4] += b.x; v[4 * q4 + 1] += b.y; v[4 * q4 + 2] += b.z; v[4 * q4 + 3] += b.w;
          }
        }
        if (P.act == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (P.act == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.01f * v[i]);
        }
        if (valid) {
          if (P.out_f32) {
            float *dst = static_cast<float *>(P.out) + p * P.Cout + n0 + c0;
#pragma unroll
            for (int k = 0; k < 4; ++k) stg256(dst + 8 * k, reinterpret_cast<const uint32_t *>(v) + 8 * k);
          } else {
            uint32_t q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
            __nv_bfloat16 *dst = static_cast<__nv_bfloat16 *>(P.out) + p * P.Cout + n0 + c0;
            stg256(dst, q);
            stg256(dst + 16, q + 8);
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

template <int NT>
static int launch(const Args &a, const CUtensorMap &am, const CUtensorMap &wm, cudaStream_t st) {
  const size_t smem = (size_t)kAStages * kASlab + (size_t)b_stages(NT) * NT * 128 +
                      B_COUNT * sizeof(uint64_t) + 16;
  auto kern = linear_kernel<NT>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  {
    KernelTimer timer(SBMC_KERNEL_CONV1X1, st);
    kern<<<(unsigned)grid, kThreads, smem, st>>>(am, wm, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace lin
}  // namespace sbmc

extern "C" int sbmc_linear_nhwc_bf16(const void *x, const void *w, const float *bias, void *y,
                                     int64_t pixels, int cin, int cout, int act, int out_f32,
                                     void *stream) {
  using namespace sbmc;
  if (pixels < 0 || cin < 1 || cout < 1 || act < 0 || act > 2) {
    set_error("linear: invalid shape");
    return SBMC_EINVAL;
  }
  if (pixels == 0) return SBMC_OK;
  if (!x || !w || !y) {
    set_error("linear: null pointer argument");
    return SBMC_EINVAL;
  }
  if (cin % 64 != 0 || cout % 128 != 0 || pixels >= (1ll << 31)) {
    set_error("linear: needs cin %% 64 == 0 and cout %% 128 == 0 (got %d, %d)", cin, cout);
    return SBMC_EUNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
       reinterpret_cast<uintptr_t>(w)) & 31) {
    set_error("linear: pointers must be 32-byte aligned");
    return SBMC_EALIGN;
  }
  const int nt = (cout % 256 == 0) ? 256 : 128;
  lin::Args a;
  a.bias = bias; a.out = y; a.act = act; a.out_f32 = out_f32 ? 1 : 0;
  a.P = pixels; a.Cin = cin; a.Cout = cout;
  a.tiles_p = (pixels + lin::kTilePx - 1) / lin::kTilePx;
  a.ntiles = a.tiles_p * (cout / nt);
  CUtensorMap am, wm;
  if (!encode_tensor_map_bf16_2d_sw128(&am, x, (uint64_t)cin, (uint64_t)pixels, 64, lin::kTilePx) ||
      !encode_tensor_map_bf16_2d_sw128(&wm, w, (uint64_t)cin, (uint64_t)cout, 64, (uint32_t)nt))
    return SBMC_ECUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  note_path(1);
  return nt == 256 ? lin::launch<256>(a, am, wm, st) : lin::launch<128>(a, am, wm, st);
}
