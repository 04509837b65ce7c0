// umma_probe2.cu -- developer probe (round 2) for the tcgen05 features the pipelined
// 1x1 chain and the 3x3 implicit-GEMM kernels build on.  Each test is a separate
// invocation so that a fault / hang in one does not take the others down:
//
//   umma_probe2 ts        A operand in TMEM (written with tcgen05.st as packed bf16),
//                         B in shared memory: D = A . B^T vs a CPU GEMM
//   umma_probe2 shift     A operand = rows [s, s+128) of a taller SWIZZLE_128B slab
//                         (descriptor start address moved by s rows, with and without
//                         the base-offset field): the "halo tile, nine shifted views"
//                         trick of the 3x3 convolution
//   umma_probe2 pair      cta_group::2: M=256 over a CTA pair, B split along N
//   umma_probe2 rate      cycles per MMA for SS / TS, N = 64 / 128 / 256, and for the
//                         CTA pair; TMEM read rate of the epilogue (4 / 8 warps)
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../sbmc_b200/csrc/umma.cuh"
using namespace sbmc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, bool acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, bool acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// ---------------------------------------------------------------------------------
// ts: D[128 x N] = A[128 x K] (TMEM, bf16 packed two per column) . B[N x K]^T (smem)
// ---------------------------------------------------------------------------------
template <int N, int K>
__global__ void __launch_bounds__(128) k_ts(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sB = smem;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  const uint32_t row_addr = tm + ((uint32_t)(warp * 32) << 16);
  for (int row = tid; row < N; row += 128)
    for (int kb = 0; kb < K / 64; ++kb)
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4 *>(sB + kb * N * 128 + sw128_offset(row, j)) =
            *reinterpret_cast<const uint4 *>(B + (size_t)row * K + kb * 64 + j * 8);
  // A: thread = row; element k lives in column 256 + k/2 (low half = even k)
  for (int c = 0; c < K / 2; c += 8) {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) r[j] = *reinterpret_cast<const uint32_t *>(A + (size_t)tid * K + 2 * (c + j));
    tmem_st_32x32b_x8(row_addr + 256 + c, r);
  }
  tmem_wait_st();
  fence_proxy_async();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t bd = umma_smem_desc_sw128(sB + (k >> 2) * N * 128) + (uint64_t)((k & 3) * 2);
      umma_bf16_ts(tm, tm + 256 + k * 8, bd, idesc, k > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32b_x32(row_addr + c0, v);
    for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------
// shift: A slab has ROWS rows (multiple of 8) x 64 bf16, SWIZZLE_128B by absolute
// address; D = A[s : s+128] . B^T.  mode 0: start address + s*128, base offset 0;
// mode 1: same with base_offset = (start >> 7) & 7.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_shift(const __nv_bfloat16 *A, const __nv_bfloat16 *B,
                                               float *D, int rows, int shift, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                      // rows x 128 B
  unsigned char *sB = smem + 160 * 128 * 2;      // K = 128: two slabs of 128 rows... (K = 64 here: one)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  for (int row = tid; row < rows; row += 128)
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4 *>(sA + sw128_offset(row, j)) =
          *reinterpret_cast<const uint4 *>(A + (size_t)row * 64 + j * 8);
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4 *>(sB + sw128_offset(tid, j)) =
        *reinterpret_cast<const uint4 *>(B + (size_t)tid * 64 + j * 8);
  fence_proxy_async();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = umma_smem_desc_sw128(sA + shift * 128) + (uint64_t)(k * 2);
      if (mode == 1) ad |= (uint64_t)(((smem_u32(sA) + shift * 128) >> 7) & 7) << 49;
      const uint64_t bd = umma_smem_desc_sw128(sB) + (uint64_t)(k * 2);
      umma_bf16(tm, ad, bd, idesc, k > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) D[(size_t)tid * 128 + c0 + i] = v[i];
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------
// pair: cta_group::2.  CTA r holds A rows [128 r, 128 r + 128) and B rows
// [N/2 r, N/2 r + N/2); D[256 x N]: CTA r's TMEM holds rows [128 r, +128).
// ---------------------------------------------------------------------------------
template <int N, int K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
k_pair(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int reps, long long *cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                          // K/64 slabs of 128 rows
  unsigned char *sB = smem + (K / 64) * 128 * 128;   // K/64 slabs of N/2 rows
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int kb = 0; kb < K / 64; ++kb)
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4 *>(sA + kb * 128 * 128 + sw128_offset(tid, j)) =
          *reinterpret_cast<const uint4 *>(A + (size_t)(rank * 128 + tid) * K + kb * 64 + j * 8);
  for (int row = tid; row < N / 2; row += 128)
    for (int kb = 0; kb < K / 64; ++kb)
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4 *>(sB + kb * (N / 2) * 128 + sw128_offset(row, j)) =
            *reinterpret_cast<const uint4 *>(B + (size_t)(rank * (N / 2) + row) * K + kb * 64 + j * 8);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, N);
    t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < K / 16; ++k) {
        const uint64_t ad = umma_smem_desc_sw128(sA + (k >> 2) * 128 * 128) + (uint64_t)((k & 3) * 2);
        const uint64_t bd = umma_smem_desc_sw128(sB + (k >> 2) * (N / 2) * 128) + (uint64_t)((k & 3) * 2);
        umma_bf16_2sm(tm, ad, bd, idesc, (r | k) > 0);
      }
    umma_commit_2sm(&bar, 3);
  }
  mbar_wait(&bar, 0);
  if (rank == 0 && tid == 0) { t1 = clock64(); if (cycles) cycles[blockIdx.x / 2] = t1 - t0; }
  tcgen05_fence_after();
  if (D)
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int i = 0; i < 32; ++i) D[(size_t)(rank * 128 + tid) * N + c0 + i] = v[i];
    }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------
// rate: one CTA per SM, `reps` x (K/16) MMAs back to back on resident operands.
// ---------------------------------------------------------------------------------
template <int N, int TS>
__global__ void __launch_bounds__(256) k_rate(int reps, long long *cycles, long long *ld_cycles, int ld_warps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                     // 2 slabs (K = 128)
  unsigned char *sB = smem + 2 * 128 * 128;     // 2 slabs of N rows
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  for (int i = tid; i < (2 * 128 * 128 + 2 * N * 128) / 4; i += 256) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t bd = umma_smem_desc_sw128(sB + (k >> 2) * N * 128) + (uint64_t)((k & 3) * 2);
        if (TS) {
          umma_bf16_ts(tm, tm + 256 + k * 8, bd, idesc, (r | k) > 0);
        } else {
          const uint64_t ad = umma_smem_desc_sw128(sA + (k >> 2) * 128 * 128) + (uint64_t)((k & 3) * 2);
          umma_bf16(tm, ad, bd, idesc, (r | k) > 0);
        }
      }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  tcgen05_fence_after();
  // epilogue read rate: ld_warps warps read 128 columns of their lane quadrant, 64 times
  if (warp < ld_warps) {
    const int colbase = (ld_warps == 8) ? (warp >> 2) * 64 : 0;
    const int ncols = (ld_warps == 8) ? 64 : 128;
    float acc = 0.f;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < 64; ++it)
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        float v[32];
        tmem_ld_32x32b_x32(tm + ((uint32_t)((warp & 3) * 32) << 16) + colbase + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
      }
    const long long t1 = clock64();
    if (acc == 12345.f) cycles[0] = 0;
    if ((tid & 31) == 0 && warp == 0) ld_cycles[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static void fill(std::vector<__nv_bfloat16> &h, std::vector<float> &f) {
  for (size_t i = 0; i < h.size(); ++i) {
    h[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.f);
    f[i] = __bfloat162float(h[i]);
  }
}

template <int N, int K>
static int test_ts() {
  std::vector<__nv_bfloat16> hA(128 * K), hB(N * K);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(3); fill(hA, fA); fill(hB, fB);
  __nv_bfloat16 *dA, *dB; float *dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(K / 64) * N * 128;
  CK(cudaFuncSetAttribute(k_ts<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_ts<N, K><<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ts N=%d K=%d FAULT %s\n", N, K, cudaGetErrorString(e)); return 2; }
  std::vector<float> hD(128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)fA[m * K + k] * fB[n * K + k];
      maxerr = fmax(maxerr, fabs(acc - hD[m * N + n]));
    }
  printf("ts N=%d K=%d max abs err %.3e %s\n", N, K, maxerr, maxerr < 1e-2 ? "ok" : "WRONG");
  return maxerr < 1e-2 ? 0 : 3;
}

static int test_shift() {
  const int rows = 152;
  std::vector<__nv_bfloat16> hA(rows * 64), hB(128 * 64);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(5); fill(hA, fA); fill(hB, fB);
  __nv_bfloat16 *dA, *dB; float *dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 160 * 128 * 2 + 128 * 128;
  CK(cudaFuncSetAttribute(k_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int bad = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 1, 2, 3, 5, 7, 8, 9, 17, 24}) {
      k_shift<<<1, 128, smem>>>(dA, dB, dD, rows, shift, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift mode=%d s=%d FAULT %s\n", mode, shift, cudaGetErrorString(e)); return 2; }
      std::vector<float> hD(128 * 128);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
          double acc = 0;
          for (int k = 0; k < 64; ++k) acc += (double)fA[(m + shift) * 64 + k] * fB[n * 64 + k];
          maxerr = fmax(maxerr, fabs(acc - hD[m * 128 + n]));
        }
      printf("shift mode=%d (base_offset %s) s=%2d max abs err %.3e %s\n", mode, mode ? "set" : "0", shift,
             maxerr, maxerr < 1e-2 ? "ok" : "WRONG");
      if (mode == 0 && maxerr >= 1e-2) bad = 1;
    }
  return bad;
}

template <int N, int K>
static int test_pair(int reps, bool check) {
  std::vector<__nv_bfloat16> hA(256 * K), hB(N * K);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(7); fill(hA, fA); fill(hB, fB);
  __nv_bfloat16 *dA, *dB; float *dD; long long *dC;
  const int pairs = check ? 1 : 74;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 256 * N * 4));
  CK(cudaMalloc(&dC, pairs * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(K / 64) * (128 + N / 2) * 128;
  CK(cudaFuncSetAttribute(k_pair<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_pair<N, K><<<2 * pairs, 128, smem>>>(dA, dB, check ? dD : nullptr, reps, dC);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("pair N=%d K=%d FAULT %s\n", N, K, cudaGetErrorString(e)); return 2; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> hc(pairs);
  CK(cudaMemcpy(hc.data(), dC, pairs * 8, cudaMemcpyDeviceToHost));
  if (!check) {
    const double nm = (double)reps * (K / 16);
    printf("pair rate M=256 N=%d: %.1f cycles / MMA (pair 0), kernel %.3f ms, %.1f TFLOP/s over %d pairs\n",
           N, hc[0] / nm, ms, 2.0 * 256 * N * 16 * nm * pairs / (ms * 1e-3) / 1e12, pairs);
    return 0;
  }
  std::vector<float> hD(256 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)fA[m * K + k] * fB[n * K + k];
      maxerr = fmax(maxerr, fabs(acc - hD[m * N + n]));
    }
  printf("pair M=256 N=%d K=%d max abs err %.3e %s\n", N, K, maxerr, maxerr < 1e-2 ? "ok" : "WRONG");
  return maxerr < 1e-2 ? 0 : 3;
}

template <int N, int TS>
static void test_rate(int reps, int ld_warps) {
  long long *dC, *dL;
  CK(cudaMalloc(&dC, 148 * 8)); CK(cudaMalloc(&dL, 148 * 8));
  const size_t smem = 2 * 128 * 128 + 2 * N * 128;
  CK(cudaFuncSetAttribute(k_rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_rate<N, TS><<<148, 256, smem>>>(64, dC, dL, ld_warps);
  cudaEventRecord(e0);
  k_rate<N, TS><<<148, 256, smem>>>(reps, dC, dL, ld_warps);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rate N=%d TS=%d FAULT %s\n", N, TS, cudaGetErrorString(e)); exit(2); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc[148], hl[148];
  CK(cudaMemcpy(hc, dC, sizeof(hc), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hl, dL, sizeof(hl), cudaMemcpyDeviceToHost));
  const double nm = (double)reps * 8;
  printf("rate %s M=128 N=%3d: %.1f cycles / MMA (CTA 0), kernel %.3f ms, %.1f TFLOP/s chip; "
         "TMEM read with %d warps: %.0f cycles per 128x128 fp32 accumulator\n",
         TS ? "TS" : "SS", N, hc[0] / nm, ms, 2.0 * 128 * N * 16 * nm * 148 / (ms * 1e-3) / 1e12, ld_warps,
         hl[0] / 64.0);
}

int main(int argc, char **argv) {
  const char *t = argc > 1 ? argv[1] : "";
  if (!strcmp(t, "ts")) return test_ts<128, 128>() | test_ts<64, 128>() | test_ts<224, 128>();
  if (!strcmp(t, "shift")) return test_shift();
  if (!strcmp(t, "pair")) return test_pair<256, 128>(1, true) | test_pair<128, 128>(1, true);
  if (!strcmp(t, "pair_rate")) { test_pair<256, 128>(2048, false); test_pair<128, 128>(2048, false); return 0; }
  if (!strcmp(t, "rate")) {
    test_rate<64, 0>(2048, 4); test_rate<128, 0>(2048, 4); test_rate<256, 0>(2048, 8);
    test_rate<128, 1>(2048, 8); test_rate<256, 1>(2048, 4);
    return 0;
  }
  printf("usage: umma_probe2 ts|shift|pair|pair_rate|rate\n");
  return 9;
}
