// umma.cuh -- tcgen05 / TMEM / UMMA-descriptor helpers (sm_100a inline PTX).
//
// Bit layouts follow the CUTLASS definitions vendored in this image
// (cute/arch/mma_sm100_desc.hpp: UMMA::SmemDescriptor, UMMA::InstrDescriptor;
// cute/atom/mma_traits_sm100.hpp: canonical K-major SWIZZLE_128B layout
// "Swizzle<3,4,3> o ((8,n),2):((8T,SBO),1)" in 16-byte units).
#pragma once
#include "common.cuh"

namespace sbmc {

bool encode_tensor_map_bf16_2d_sw128(CUtensorMap *map, const void *base, uint64_t inner,
                                     uint64_t rows, uint32_t box_inner, uint32_t box_rows);

bool encode_tensor_map_bf16_3d_sw128(CUtensorMap *map, const void *base, uint64_t inner,
                                     uint64_t rows, uint64_t n, uint64_t img_stride,
                                     uint32_t box_rows);

bool encode_tensor_map_bf16_sw128(CUtensorMap *map, const void *base, int rank,
                                  const uint64_t *dims, const uint64_t *strides_bytes,
                                  const uint32_t *box);

#ifdef __CUDACC__

// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a K-major
// SWIZZLE_128B slab (rows of 128 bytes, 8-row groups of 1024 bytes): the chunk
// index is XOR-ed with the row index modulo 8.  The slab base must be 1024-byte
// aligned.
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row group stride 1024 B.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(const void *smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
  return d;
}

// Instruction descriptor: A, B bf16 K-major, D fp32, dense, M x N tile.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// All previously issued MMAs of this thread arrive on the mbarrier when done
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(smem_u32(bar))
               : "memory");
}

// One lane of a converged warp (elect.sync).  MMA issue belongs inside
// `if (elect_one()) { ... }` of a warp whose control flow is otherwise uniform: under a
// plain `if (lane == 0)` the compiler cannot prove the tcgen05 operands uniform and wraps
// every UTCHMMA in an ELECT / BRA.U.ANY loop (about ten extra instructions per MMA, which
// made the single issuing thread the bottleneck of the N = 128 kernels).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// One full warp allocates `cols` TMEM columns (power of two >= 32); the base
// address lands in *dst (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t *dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::
                   "r"(smem_u32(dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols)
               : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes
// 32 (w % 4) .. +31: taddr = base + (lane_base << 16) + column).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map,
                                            uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map,
                                            uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- epilogue stores: 4 x 4 transpose of 32-byte elements inside a lane quad ----
// A TMEM epilogue thread owns one tensor row (pixel): its 128 output bytes are
// contiguous, but one warp-wide store then touches 32 different 128-byte lines and
// the LSU retires ~1 line per cycle -- measured: the 16-byte-per-lane stores of the
// embedding chain cost as much as the rest of the kernel (profiles/r2e_*).  Two
// butterfly exchanges (shfl.xor 2, then 1) turn "lane q holds the four 32-byte
// elements of row q" into "lane q holds element q of the four rows of its quad",
// so that a 256-bit store instruction writes 8 complete 128-byte lines.
// In: r[8 c + i] = word i of element c of this lane's row.  Out: r[8 k + i] = word i
// of element (lane & 3) of the row of lane (lane & ~3) + k.
__device__ __forceinline__ void quad_transpose32(uint32_t (&r)[32], int lane) {
  const bool b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t lo = r[i], hi = r[i + 16];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, b1 ? lo : hi, 2);
    r[i] = b1 ? recv : lo;
    r[i + 16] = b1 ? hi : recv;
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t x = r[16 * rr + i], y = r[16 * rr + 8 + i];
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, b0 ? x : y, 1);
      r[16 * rr + i] = b0 ? recv : x;
      r[16 * rr + 8 + i] = b0 ? y : recv;
    }
}
// 256-bit global store (sm_100: STG.256), p 32-byte aligned.
__device__ __forceinline__ void stg256(void *p, const uint32_t *v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// Backward of ReLU / LeakyReLU(0.01) fused into an epilogue: v[i] *= act'(pre-activation),
// read off the sign of the saved OUTPUT m[i] of that activation (bf16, 32 consecutive
// channels of this thread's row, 16-byte aligned): m > 0 <=> pre-activation > 0, and the
// derivative at 0 is `slope` (0 for ReLU) like torch's.
__device__ __forceinline__ void apply_act_mask32(float (&v)[32], const void *m, float slope) {
  const uint4 *mp = reinterpret_cast<const uint4 *>(m);
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const uint4 mm = __ldg(mp + q4);
    const uint32_t w4[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = w4[j] & 0xFFFFu, hi = w4[j] >> 16;
      // bf16 > 0: sign bit clear and not zero
      if (!(lo != 0 && lo < 0x8000u)) v[8 * q4 + 2 * j] *= slope;
      if (!(hi != 0 && hi < 0x8000u)) v[8 * q4 + 2 * j + 1] *= slope;
    }
  }
}

#endif  // __CUDACC__
}  // namespace sbmc
