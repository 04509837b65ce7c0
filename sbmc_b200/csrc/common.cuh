// common.cuh -- shared device helpers (sm_100a PTX wrappers) and host plumbing.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sbmc_b200.h"

namespace sbmc {

typedef long long i64;

// ---- host side -----------------------------------------------------------
void set_error(const char *fmt, ...);
void note_path(int path);
void warn_generic(const char *op, int c, int kh, int kw, long long w);
void count_launch(int n = 1);
bool force_generic();
int num_sms();

// Brackets one kernel launch with CUDA events when timing is enabled
// (sbmc_b200_timing_enable); a no-op otherwise.
class KernelTimer {
 public:
  KernelTimer(int kind, cudaStream_t st);
  ~KernelTimer();
  KernelTimer(const KernelTimer &) = delete;
  KernelTimer &operator=(const KernelTimer &) = delete;

 private:
  int kind_;
  cudaStream_t st_;
  cudaEvent_t a_;
};

// cuTensorMapEncodeTiled, resolved at run time through the CUDA runtime so the
// library has no link-time dependency on libcuda (it must load, and export its
// symbols, on a machine without a driver).
bool encode_tensor_map_f32(CUtensorMap *map, const void *base, int rank,
                           const uint64_t *dims, const uint64_t *strides_bytes,
                           const uint32_t *box);

#define SBMC_CUDA_OK(expr)                                                   \
  do {                                                                       \
    cudaError_t e__ = (expr);                                                \
    if (e__ != cudaSuccess) {                                                \
      ::sbmc::set_error("%s failed: %s (%s:%d)", #expr,                      \
                        cudaGetErrorString(e__), __FILE__, __LINE__);        \
      return SBMC_ECUDA;                                                     \
    }                                                                        \
  } while (0)

static inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

// kernel launchers (one per translation unit); all return SBMC_* codes.
int launch_fwd(const float *data_ext, const float *weights, float *output,
               float *sum_w, i64 n, int c, i64 h, i64 w, int kh, int kw,
               int halo_top, int halo_bot, cudaStream_t st);
// d_weights needs the image (data_ext) around the band, d_data scatters into
// the rows around the band (d_data_ext): the two may use different halos.
int launch_bwd_dweights(const float *data_ext, const float *d_output,
                        const float *d_sum_w, float *d_weights, i64 n, int c,
                        i64 h, i64 w, int kh, int kw, int halo_top, int halo_bot,
                        cudaStream_t st);
int launch_bwd_ddata(const float *weights, const float *d_output,
                     float *d_data_ext, i64 n, int c, i64 h, i64 w, int kh,
                     int kw, int halo_top, int halo_bot, cudaStream_t st);
int launch_s2g(const float *scatter, float *gather, i64 n, int kh, int kw,
               i64 h, i64 w, cudaStream_t st);

int generic_fwd(const float *data_ext, const float *weights, float *output,
                float *sum_w, i64 n, int c, i64 h, i64 w, int kh, int kw,
                int halo_top, int halo_bot, cudaStream_t st);
int generic_bwd_dweights(const float *data_ext, const float *d_output,
                         const float *d_sum_w, float *d_weights, i64 n, int c,
                         i64 h, i64 w, int kh, int kw, int halo_top,
                         int halo_bot, cudaStream_t st);
int generic_bwd_ddata(const float *weights, const float *d_output,
                      float *d_data_ext, i64 n, int c, i64 h, i64 w, int kh,
                      int kw, int halo_top, int halo_bot, cudaStream_t st);
int generic_s2g(const float *scatter, float *gather, i64 n, int kh, int kw,
                i64 h, i64 w, cudaStream_t st);

// ---- device side -----------------------------------------------------------
#ifdef __CUDACC__

// Streaming 128-bit load: read-only path, do not allocate in L1 (each weight
// is touched exactly once).
__device__ __forceinline__ float4 ldg_stream(const float *p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Streaming 128-bit store (write-once outputs): evict-first, no L1 allocate.
__device__ __forceinline__ void stg_stream(float *p, const float4 &v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// Store-policy variants for the tuning sweep: 0 = .cs (evict-first), 1 = default
// write-back, 2 = no L1 allocation.
template <int POLICY>
__device__ __forceinline__ void stg_policy(float *p, const float4 &v) {
  if (POLICY == 1)
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  else if (POLICY == 2)
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  else
    stg_stream(p, v);
}
// L2 eviction-priority policies: the K*K-sized streams are touched once
// (evict_first), the image-sized operands are re-read by the next kernel of the
// forward/backward sequence (evict_last).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_hint(const float *p, uint64_t policy) {
  float4 v;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void stg_hint(float *p, const float4 &v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy)
               : "memory");
}
__device__ __forceinline__ float4 ldg_cached(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA tiled load, 4-D box, completion signalled on an mbarrier.  Out-of-bounds
// elements of the box are zero-filled: this *is* the reference's
// constant_exterior(…, 0) boundary condition.
__device__ __forceinline__ void tma_load_4d(void *smem_dst,
                                            const CUtensorMap *map,
                                            uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::"
      "bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_hint(void *smem_dst, const CUtensorMap *map,
                                                 uint64_t *bar, int c0, int c1, int c2, int c3,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}
// TMA tiled store, 4-D box; out-of-bounds elements are not written.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map,
                                             const void *smem_src, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, "
      "%5}], [%1];" ::"l"(map),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_commit_group() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

#endif  // __CUDACC__

}  // namespace sbmc
