// umma_probe.cu -- developer probe for the tcgen05 building blocks used by
// conv1x1.cu: D[128 x N] (TMEM, fp32) = A[128 x K] * B[N x K]^T with bf16
// K-major operands in 128B-swizzled shared memory.  Variant 0: both operands
// written by threads; variant 1: B loaded by TMA with a SWIZZLE_128B tensor map.
// usage: umma_probe <variant> <N> <K>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sbmc_b200/csrc/umma.cuh"
using namespace sbmc;

template <int N, int K>
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16 *A, const __nv_bfloat16 *B,
                                             const __grid_constant__ CUtensorMap bmap,
                                             float *D, int variant) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = smem;                       // K/64 slabs of 128 rows x 128 B
  unsigned char *sB = smem + (K / 64) * 128 * 128;  // K/64 slabs of N rows x 128 B
  __shared__ __align__(8) uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // A: thread = row; 16-byte chunks of 8 bf16 along K
  for (int kb = 0; kb < K / 64; ++kb)
    for (int j = 0; j < 8; ++j) {
      const uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)tid * K + kb * 64 + j * 8);
      *reinterpret_cast<uint4 *>(sA + kb * 128 * 128 + sw128_offset(tid, j)) = v;
    }
  if (variant == 0) {
    for (int row = tid; row < N; row += 128)
      for (int kb = 0; kb < K / 64; ++kb)
        for (int j = 0; j < 8; ++j) {
          const uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)row * K + kb * 64 + j * 8);
          *reinterpret_cast<uint4 *>(sB + kb * N * 128 + sw128_offset(row, j)) = v;
        }
  } else if (tid == 0) {
    mbar_expect_tx(&bar_tma, (uint32_t)(N * K * 2));
    for (int kb = 0; kb < K / 64; ++kb)
      tma_load_2d(sB + kb * N * 128, &bmap, &bar_tma, kb * 64, 0);
  }
  fence_proxy_async();
  __syncthreads();
  if (variant == 1) mbar_wait(&bar_tma, 0);
  if (tid == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, N);
    for (int k = 0; k < K / 16; ++k) {
      const int kb = k / 4, ki = k % 4;
      const uint64_t ad = umma_smem_desc_sw128(sA + kb * 128 * 128) + (uint64_t)(ki * 2);
      const uint64_t bd = umma_smem_desc_sw128(sB + kb * N * 128) + (uint64_t)(ki * 2);
      umma_bf16(tmem_base, ad, bd, idesc, k > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N, int K>
int run(int variant) {
  std::vector<__nv_bfloat16> hA(128 * K), hB(N * K);
  std::vector<float> fA(128 * K), fB(N * K);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float *dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap bmap;
  if (!encode_tensor_map_bf16_2d_sw128(&bmap, dB, K, N, 64, N)) { printf("encode failed: %s\n", sbmc_b200_last_error()); return 1; }
  const size_t smem = (size_t)(K / 64) * (128 + N) * 128 + 1024;
  cudaFuncSetAttribute(probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<N, K><<<1, 128, smem>>>(dA, dB, bmap, dD, variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d N=%d K=%d FAULT %s\n", variant, N, K, cudaGetErrorString(e)); return 2; }
  std::vector<float> hD(128 * N);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)fA[m * K + k] * fB[n * K + k];
      const double err = fabs(acc - hD[m * N + n]);
      if (err > maxerr) maxerr = err;
    }
  printf("variant %d N=%d K=%d max abs err %.3e %s\n", variant, N, K, maxerr, maxerr < 1e-2 ? "ok" : "WRONG");
  return maxerr < 1e-2 ? 0 : 3;
}

int main(int argc, char **argv) {
  const int variant = atoi(argv[1]), N = atoi(argv[2]), K = atoi(argv[3]);
  if (N == 128 && K == 64) return run<128, 64>(variant);
  if (N == 128 && K == 128) return run<128, 128>(variant);
  if (N == 128 && K == 256) return run<128, 256>(variant);
  if (N == 224 && K == 128) return run<224, 128>(variant);
  if (N == 64 && K == 128) return run<64, 128>(variant);
  printf("unsupported N/K\n");
  return 9;
}
