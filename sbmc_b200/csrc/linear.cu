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
  const __nv_bfloat16 *mask;   // may be null: [rows][Cout], out *= act'(mask) (mask_act)
  int act;                 // 0 none, 1 ReLU, 2 LeakyReLU(0.01)
  int mask_act;            // derivative selected by the sign of mask: 1 ReLU, 2 LeakyReLU
  int out_mode;            // 0 bf16 rows, 1 fp32 rows, 2 fp32 channel planes
  long long P;             // rows
  int Cin, Cout;           // Cin = CinA + CinB
  int slabs_a;             // 64-channel slabs that come from the first source
  long long hw, spp;       // row r = (image b, sample s, pixel p), r = (b spp + s) hw + p
  long long out_img_stride, out_smp_stride;   // plane mode: elements between images / samples
  int cout_valid;          // plane mode: channels >= cout_valid are not stored
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
linear_kernel(const __grid_constant__ CUtensorMap amap,      // x {CinA, P}
              const __grid_constant__ CUtensorMap amap2,     // xb {CinB, pixels} (or = amap)
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
        // second source: one row per PIXEL, shared by the samples of the pixel
        // (hw % 256 == 0 is checked by the host, so a tile never straddles two samples)
        const long long q0 = (P.slabs_a < nslabs) ? (p0 / (P.spp * P.hw)) * P.hw + p0 % P.hw : 0;
        for (int s = 0; s < nslabs; ++s) {
          mbar_wait(bars + B_AE + ab, ((ph >> ab) & 1) ^ 1); ph ^= 1u << ab;
          mbar_expect_tx(bars + B_AF + ab, (uint32_t)kASlab);
          if (s < P.slabs_a)
            tma_load_2d(sA + ab * kASlab, &amap, bars + B_AF + ab, s * 64, (int)p0);
          else
            tma_load_2d(sA + ab * kASlab, &amap2, bars + B_AF + ab, (s - P.slabs_a) * 64, (int)q0);
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
      float *plane0 = nullptr;
      if (P.out_mode == 2 && valid) {
        const long long img = p / P.hw;
        plane0 = static_cast<float *>(P.out) + (img / P.spp) * P.out_img_stride +
                 (img % P.spp) * P.out_smp_stride + (p - img * P.hw);
      }
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
        if (P.mask && valid)
          apply_act_mask32(v, P.mask + p * P.Cout + n0 + c0, (P.mask_act == 2) ? 0.01f : 0.f);
        if (valid) {
          if (P.out_mode == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + c0 + i < P.cout_valid) plane0[(long long)(n0 + c0 + i) * P.hw] = v[i];
          } else if (P.out_mode == 1) {
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
static int launch(const Args &a, const CUtensorMap &am, const CUtensorMap &am2,
                  const CUtensorMap &wm, cudaStream_t st) {
  const size_t smem = (size_t)kAStages * kASlab + (size_t)b_stages(NT) * NT * 128 +
                      B_COUNT * sizeof(uint64_t) + 16;
  auto kern = linear_kernel<NT>;
  static bool configured = false;     // once: the call is not a stream operation
  if (!configured) {
    SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
  {
    KernelTimer timer(SBMC_KERNEL_CONV1X1, st);
    kern<<<(unsigned)grid, kThreads, smem, st>>>(am, am2, wm, a);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace lin
}  // namespace sbmc

extern "C" int sbmc_linear2_nhwc_bf16(const void *x, int cin_a, const void *xb, int cin_b,
                                      int64_t hw, int64_t spp, const void *w,
                                      const float *bias, const void *mask, int mask_act,
                                      void *y, int out_mode, int64_t out_img_stride,
                                      int64_t out_smp_stride, int cout_valid, int64_t rows,
                                      int cout, int act, void *stream) {
  using namespace sbmc;
  if (rows < 0 || cin_a < 1 || cin_b < 0 || cout < 1 || act < 0 || act > 2 || out_mode < 0 ||
      out_mode > 2 || (mask && (mask_act < 1 || mask_act > 2))) {
    set_error("linear: invalid argument");
    return SBMC_EINVAL;
  }
  if (rows == 0) return SBMC_OK;
  if (!x || !w || !y || (cin_b > 0 && !xb)) {
    set_error("linear: null pointer argument");
    return SBMC_EINVAL;
  }
  if (cin_a % 64 != 0 || cin_b % 64 != 0 || cout % 128 != 0 || rows >= (1ll << 31)) {
    set_error("linear: needs cin %% 64 == 0 and cout %% 128 == 0 (got %d + %d, %d)", cin_a,
              cin_b, cout);
    return SBMC_EUNSUPPORTED;
  }
  if ((cin_b > 0 || out_mode == 2) && (hw < 1 || spp < 1 || rows % hw != 0)) {
    set_error("linear: rows must be images x samples x hw pixels");
    return SBMC_EINVAL;
  }
  if (cin_b > 0 && hw % lin::kTilePx != 0) {
    set_error("linear: the two-source form needs hw %% %d == 0 (got %lld)", lin::kTilePx,
              (long long)hw);
    return SBMC_EUNSUPPORTED;
  }
  if (mask && out_mode == 2) {
    set_error("linear: no mask in plane mode");
    return SBMC_EUNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
       reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(xb) |
       reinterpret_cast<uintptr_t>(mask)) & 31) {
    set_error("linear: pointers must be 32-byte aligned");
    return SBMC_EALIGN;
  }
  const int nt = (cout % 256 == 0) ? 256 : 128;
  lin::Args a;
  a.bias = bias; a.out = y; a.act = act; a.out_mode = out_mode;
  a.mask = static_cast<const __nv_bfloat16 *>(mask); a.mask_act = mask_act;
  a.P = rows; a.Cin = cin_a + cin_b; a.Cout = cout; a.slabs_a = cin_a / 64;
  a.hw = hw > 0 ? hw : 1; a.spp = spp > 0 ? spp : 1;
  a.out_img_stride = out_img_stride; a.out_smp_stride = out_smp_stride;
  a.cout_valid = cout_valid > 0 ? cout_valid : cout;
  a.tiles_p = (rows + lin::kTilePx - 1) / lin::kTilePx;
  a.ntiles = a.tiles_p * (cout / nt);
  CUtensorMap am, am2, wm;
  if (!encode_tensor_map_bf16_2d_sw128(&am, x, (uint64_t)cin_a, (uint64_t)rows, 64, lin::kTilePx) ||
      !encode_tensor_map_bf16_2d_sw128(&wm, w, (uint64_t)a.Cin, (uint64_t)cout, 64, (uint32_t)nt))
    return SBMC_ECUDA;
  am2 = am;
  if (cin_b > 0 && !encode_tensor_map_bf16_2d_sw128(&am2, xb, (uint64_t)cin_b,
                                                    (uint64_t)(rows / spp), 64, lin::kTilePx))
    return SBMC_ECUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  note_path(1);
  return nt == 256 ? lin::launch<256>(a, am, am2, wm, st) : lin::launch<128>(a, am, am2, wm, st);
}

extern "C" int sbmc_linear_nhwc_bf16(const void *x, const void *w, const float *bias, void *y,
                                     int64_t pixels, int cin, int cout, int act, int out_f32,
                                     void *stream) {
  return sbmc_linear2_nhwc_bf16(x, cin, nullptr, 0, 0, 0, w, bias, nullptr, 0, y,
                                out_f32 ? 1 : 0, 0, 0, 0, pixels, cout, act, stream);
}
