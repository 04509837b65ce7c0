// weight_bank.cu -- weight normalization of ALL convolutions of the model in one launch,
// forward and backward (training pipeline, sbmc_b200/weight_bank.py).
//
// Every convolution of the reference model is wrapped in `nn.utils.weight_norm`
// (sbmc/modules.py:84-87,176-179): w[co] = g[co] * v[co] / ||v[co]||.  In eager PyTorch a
// training step spends ~10 small kernels per convolution on it (the normalization, its
// backward, layout changes and casts for the tensor-core operands; 57 convolutions).  Here
//
//   wn_prepare_kernel   reads v, g of every convolution and writes the two bf16 operand
//                       layouts the tcgen05 kernels consume: F = [T][cout_pad][cin_pad]
//                       (forward: conv3x3.cu / linear.cu B operand) and D = the operand of
//                       the data-gradient call ([T][cin][cout] with the taps flipped for
//                       3x3, the transpose [cin_pad][cout_pad] for 1x1), plus 1 / ||v||;
//   wn_backward_kernel  reads the weight gradients dW = [T][cout][cin] (fp32, exactly what
//                       csrc/wgrad.cu writes) and produces dv, dg:
//                         dg[co] = <dW[co], v[co]> / ||v[co]||
//                         dv[co] = g[co] / ||v[co]|| * (dW[co] - v[co] <dW[co], v[co]> / ||v[co]||^2)
//
// A CTA owns 8 output channels of one convolution.  Forward: [8][64 ci][T] tiles go through
// shared memory so that both the [co][ci][t] side (v) and the [t][co][ci] / transposed
// sides (F, D) move in contiguous 16-byte pieces.  Backward: one warp per channel.  Tables: entries int64 [ne][16], blocks int64 [nb][2]
// = (entry, first channel).
#include <cuda_bf16.h>

#include "common.cuh"

namespace sbmc {
namespace wb {

constexpr int kCo = 8, kCi = 64, kMaxT = 9;

struct Entry {
  const float *v, *g;
  __nv_bfloat16 *F, *D;
  float *rn;
  const float *dW;
  float *dv, *dg;
  int cout, cin, T, coutp, cinp;
};

__device__ __forceinline__ Entry load_entry(const long long *tab, long long e) {
  const long long *r = tab + 16 * e;
  Entry x;
  x.v = reinterpret_cast<const float *>(static_cast<uintptr_t>(r[0]));
  x.g = reinterpret_cast<const float *>(static_cast<uintptr_t>(r[1]));
  x.F = reinterpret_cast<__nv_bfloat16 *>(static_cast<uintptr_t>(r[2]));
  x.D = reinterpret_cast<__nv_bfloat16 *>(static_cast<uintptr_t>(r[3]));
  x.rn = reinterpret_cast<float *>(static_cast<uintptr_t>(r[4]));
  x.dW = reinterpret_cast<const float *>(static_cast<uintptr_t>(r[5]));
  x.dv = reinterpret_cast<float *>(static_cast<uintptr_t>(r[6]));
  x.dg = reinterpret_cast<float *>(static_cast<uintptr_t>(r[7]));
  x.cout = (int)r[8]; x.cin = (int)r[9]; x.T = (int)r[10]; x.coutp = (int)r[11]; x.cinp = (int)r[12];
  return x;
}

__device__ __forceinline__ float warp_sum(float a) {
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&v);
}

typedef float Tile[kCo][kCi * kMaxT + 1];

// T is a compile-time constant (1 or 9) and full tiles (8 channels x 64 inputs) take
// shift-only index arithmetic and 16-byte stores; ragged tiles take the generic loops.
template <int T>
__device__ __forceinline__ void prepare_body(const Entry &E, int co0, Tile &tile, const float *scale) {
  const int row = E.cin * T;
  const int nco = min(kCo, E.cout - co0);
  const int dco = (T == 1) ? E.coutp : E.cout;
  for (int ci0 = 0; ci0 < E.cin; ci0 += kCi) {
    const int nci = min(kCi, E.cin - ci0);
    const int run = nci * T;
    const bool full = nco == kCo && nci == kCi;
    if (full) {
      for (int i = threadIdx.x; i < kCo * kCi * T; i += 256) {
        const int c = i / (kCi * T), r = i - c * (kCi * T);
        tile[c][r] = E.v[(long long)(co0 + c) * row + ci0 * T + r] * scale[c];
      }
    } else {
      for (int i = threadIdx.x; i < nco * run; i += 256) {
        const int c = i / run, r = i - c * run;
        tile[c][r] = E.v[(long long)(co0 + c) * row + ci0 * T + r] * scale[c];
      }
    }
    __syncthreads();
    if (full) {
      // F[t][co][ci0 + 8 q .. +7]: one 16-byte store per (t, co, q)
      for (int i = threadIdx.x; i < T * kCo * 8; i += 256) {
        const int q = i & 7, c = (i >> 3) & 7, t = i >> 6;
        const float *p = &tile[c][q * 8 * T + t];
        uint4 w;
        w.x = pack2(p[0], p[T]); w.y = pack2(p[2 * T], p[3 * T]);
        w.z = pack2(p[4 * T], p[5 * T]); w.w = pack2(p[6 * T], p[7 * T]);
        *reinterpret_cast<uint4 *>(E.F + ((long long)t * E.coutp + co0 + c) * E.cinp + ci0 + q * 8) = w;
      }
      // D[..][ci][co0 .. co0 + 7]: one 16-byte store per (t, ci)
      for (int i = threadIdx.x; i < T * kCi; i += 256) {
        const int ci = i & 63, t = i >> 6;
        const int r = ci * T + t;
        uint4 w;
        w.x = pack2(tile[0][r], tile[1][r]); w.y = pack2(tile[2][r], tile[3][r]);
        w.z = pack2(tile[4][r], tile[5][r]); w.w = pack2(tile[6][r], tile[7][r]);
        const long long plane = (T == 1) ? 0 : (long long)(T - 1 - t) * E.cin * E.cout;
        *reinterpret_cast<uint4 *>(E.D + plane + (long long)(ci0 + ci) * dco + co0) = w;
      }
    } else {
      for (int i = threadIdx.x; i < T * nco * nci; i += 256) {
        const int ci = i % nci, c = (i / nci) % nco, t = i / (nci * nco);
        E.F[((long long)t * E.coutp + co0 + c) * E.cinp + ci0 + ci] =
            __float2bfloat16_rn(tile[c][ci * T + t]);
      }
      for (int i = threadIdx.x; i < T * nci * nco; i += 256) {
        const int c = i % nco, ci = (i / nco) % nci, t = i / (nco * nci);
        const long long plane = (T == 1) ? 0 : (long long)(T - 1 - t) * E.cin * E.cout;
        E.D[plane + (long long)(ci0 + ci) * dco + co0 + c] = __float2bfloat16_rn(tile[c][ci * T + t]);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
wn_prepare_kernel(const long long *__restrict__ entries, const long long *__restrict__ blocks) {
  __shared__ Tile tile;
  __shared__ float scale[kCo];
  const Entry E = load_entry(entries, blocks[2 * blockIdx.x]);
  const int co0 = (int)blocks[2 * blockIdx.x + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = E.cin * E.T;                        // floats of one output channel of v
  {
    const int co = co0 + warp;
    float ss = 0.f;
    if (co < E.cout) {
      const float *p = E.v + (long long)co * row;
      for (int i = lane; i < row; i += 32) ss = fmaf(p[i], p[i], ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) {
      const float rn = (co < E.cout) ? 1.f / sqrtf(ss) : 0.f;
      if (co < E.cout) E.rn[co] = rn;
      scale[warp] = (co < E.cout) ? E.g[co] * rn : 0.f;
    }
  }
  __syncthreads();
  if (E.T == 9) prepare_body<9>(E, co0, tile, scale);
  else prepare_body<1>(E, co0, tile, scale);
}

// Backward: one WARP per output channel, no shared memory.  dW[t][co][:] rows are read
// coalesced; the channel's row of v (cin * T floats, at most 27 KB) is re-read from L1 for
// every tap; dv[co][:] is written as one contiguous run.
template <int T>
__device__ __forceinline__ void backward_body(const Entry &E, int co) {
  const int lane = threadIdx.x & 31;
  const int row = E.cin * T;
  const float *v = E.v + (long long)co * row;
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float *g = E.dW + ((long long)t * E.cout + co) * E.cin;
    for (int ci = lane; ci < E.cin; ci += 32) acc = fmaf(g[ci], v[ci * T + t], acc);
  }
  const float dot = warp_sum(acc);
  const float rn = E.rn[co];
  const float coef = E.g[co] * rn, k2 = dot * rn * rn;
  if (lane == 0) E.dg[co] = dot * rn;
  float *dv = E.dv + (long long)co * row;
  const float *g0 = E.dW + (long long)co * E.cin;
  const long long plane = (long long)E.cout * E.cin;
  for (int r = lane; r < row; r += 32) {
    const int ci = r / T, t = r - ci * T;
    dv[r] = coef * (g0[t * plane + ci] - v[r] * k2);
  }
}

__global__ void __launch_bounds__(256)
wn_backward_kernel(const long long *__restrict__ entries, const long long *__restrict__ blocks) {
  const Entry E = load_entry(entries, blocks[2 * blockIdx.x]);
  const int co = (int)blocks[2 * blockIdx.x + 1] + (threadIdx.x >> 5);
  if (co >= E.cout) return;
  if (E.T == 9) backward_body<9>(E, co);
  else backward_body<1>(E, co);
}

}  // namespace wb
}  // namespace sbmc

extern "C" int sbmc_weight_bank_run(const int64_t *entries, const int64_t *blocks, int64_t nblocks,
                                    int backward, void *stream) {
  using namespace sbmc;
  if (nblocks < 0 || nblocks > 0x7FFFFFFF) {
    set_error("weight_bank: invalid block count");
    return SBMC_EINVAL;
  }
  if (nblocks == 0) return SBMC_OK;
  if (!entries || !blocks) {
    set_error("weight_bank: null pointer argument");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelTimer timer(SBMC_KERNEL_OTHER, st);
    if (backward)
      wb::wn_backward_kernel<<<(unsigned)nblocks, 256, 0, st>>>(
          reinterpret_cast<const long long *>(entries), reinterpret_cast<const long long *>(blocks));
    else
      wb::wn_prepare_kernel<<<(unsigned)nblocks, 256, 0, st>>>(
          reinterpret_cast<const long long *>(entries), reinterpret_cast<const long long *>(blocks));
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  note_path(1);
  return SBMC_OK;
}
