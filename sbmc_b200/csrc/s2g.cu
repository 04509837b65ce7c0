// s2g.cu -- Scatter2Gather: TMA-fed shifted plane copies.
//
// gather[n][dy][dx][y][x] = scatter[n][KH-1-dy][KW-1-dx][y+dy-c0h][x+dx-c0w],
// zero when the source pixel is outside the image
// (reference: src/scatter2gather.cpp:34-47).  That is K*K*N independent shifted
// 2-D plane copies with zero fill and no arithmetic.
//
// A CTA owns a (ROWS x 128)-pixel tile of one image and walks all K*K taps.  A
// producer warp pulls, for every tap, the box of the transposed source plane at
// the shifted coordinates with TMA into a shared-memory ring; elements outside
// the image (including negative coordinates) arrive as 0.0f -- the reference's
// boundary condition for free.  TMA needs the box to start on a 16-byte boundary
// (measured on B200: any other innermost coordinate faults with "illegal
// instruction", profiles/r1c_tma_coordinate_probe.txt), so the x shift
// s = dx - c0w is split into an aligned part, applied to the box coordinate, and
// a residue r = s mod 4 in [0, 4): boxes with r != 0 are 4 columns wider and the
// consumer warps apply the residue when they read shared memory (two aligned
// 128-bit loads + a warp-uniform select).  Each thread then writes its 4 pixels
// with one coalesced 128-bit streaming store.  Bit-exact: bytes are only moved.
#include "kw_kernels.cuh"

namespace sbmc {

template <int ROWS, int STAGES, int TPS>
__global__ void __launch_bounds__((ROWS + 1) * 32)
s2g_kernel(const __grid_constant__ CUtensorMap map128,   // box 128 x ROWS
           const __grid_constant__ CUtensorMap map132,   // box 132 x ROWS
           float *__restrict__ gather, int H, int W, int KH, int KW, int xtiles,
           int ytiles) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem_raw = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  constexpr int kSlot = tap_slot_floats(ROWS);   // floats per tap slot (max box)
  float *ring = reinterpret_cast<float *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)STAGES * TPS * kSlot);
  uint64_t *empty = full + STAGES;

  const TileCoord tc = decode_tile(blockIdx.x, xtiles, ytiles);
  const int X0 = tc.xt * kTileW, Y0 = tc.yt * ROWS;
  const int c0h = (KH - 1) / 2, c0w = (KW - 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = KH * KW;
  const int nstages = (taps + TPS - 1) / TPS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ROWS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == ROWS) {  // ---- producer warp: one lane drives the copy engine ----
    if (lane == 0) {
      for (int it = 0; it < nstages; ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
        const int t0 = it * TPS;
        const int cnt = (taps - t0 < TPS) ? (taps - t0) : TPS;
        uint32_t bytes = 0;
        for (int j = 0; j < cnt; ++j) {
          const int dx = (t0 + j) % KW;
          bytes += (uint32_t)(ROWS * (((dx - c0w) & 3) ? kBoxW : kTileW) * sizeof(float));
        }
        mbar_expect_tx(&full[s], bytes);
        for (int j = 0; j < cnt; ++j) {
          const int tap = t0 + j;
          const int dy = tap / KW, dx = tap - dy * KW;
          const int sh = dx - c0w;
          tma_load_4d(ring + ((size_t)s * TPS + j) * kSlot, (sh & 3) ? &map132 : &map128,
                      &full[s], X0 + (sh & ~3), Y0 + dy - c0h,
                      (KH - 1 - dy) * KW + (KW - 1 - dx), tc.n);
        }
      }
    }
    return;
  }

  // ---- consumer warps: warp r owns row Y0 + r, lane l pixels X0 + 4l .. +3 ----
  const int y = Y0 + warp, x0 = X0 + 4 * lane;
  const bool valid = (y < H) && (x0 < W);
  const long long plane = (long long)H * W;
  float *gp = gather + (long long)tc.n * taps * plane + (long long)y * W + x0;
  int tap = 0;
  for (int it = 0; it < nstages; ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
    const int cnt = (taps - tap < TPS) ? (taps - tap) : TPS;
    for (int j = 0; j < cnt; ++j, ++tap) {
      const int dx = tap % KW;
      const int r = (dx - c0w) & 3;
      const int pitch = r ? kBoxW : kTileW;
      const float4 v = lds_shifted4(ring + ((size_t)s * TPS + j) * kSlot + (size_t)warp * pitch,
                                    lane, r);
      if (valid) stg_stream(gp + (long long)tap * plane, v);
    }
    __syncwarp();
    if (lane == 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s]))
                   : "memory");
  }
}

static bool make_plane_map(CUtensorMap *map, const float *base, i64 n, i64 taps,
                           i64 h, i64 w, int bx, int by) {
  const uint64_t dims[4] = {(uint64_t)w, (uint64_t)h, (uint64_t)taps, (uint64_t)n};
  const uint64_t strides[3] = {(uint64_t)w * 4, (uint64_t)w * h * 4,
                               (uint64_t)w * h * taps * 4};
  const uint32_t box[4] = {(uint32_t)bx, (uint32_t)by, 1u, 1u};
  return encode_tensor_map_f32(map, base, 4, dims, strides, box);
}

template <int ROWS, int STAGES, int TPS>
int run_s2g(const float *scatter, float *gather, i64 n, int kh, int kw, i64 h, i64 w,
            cudaStream_t st) {
  CUtensorMap m128, m132;
  if (!make_plane_map(&m128, scatter, n, (i64)kh * kw, h, w, kTileW, ROWS) ||
      !make_plane_map(&m132, scatter, n, (i64)kh * kw, h, w, kBoxW, ROWS))
    return SBMC_ECUDA;
  const size_t smem = (size_t)STAGES * TPS * tap_slot_floats(ROWS) * 4 + 2 * STAGES * 8 + 128;
  auto kern = s2g_kernel<ROWS, STAGES, TPS>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int xt = (int)ceil_div(w, kTileW), yt = (int)ceil_div(h, ROWS);
  const unsigned grid = (unsigned)((i64)xt * yt * n);
  {
    KernelTimer timer(SBMC_KERNEL_S2G, st);
    kern<<<grid, (ROWS + 1) * 32, smem, st>>>(m128, m132, gather, (int)h, (int)w, kh, kw,
                                              xt, yt);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

// tuned configuration (profiles/r1e_sweep.txt): 8 rows, 3 stages of 7 taps
constexpr int kS2gRows = 8, kS2gStages = 3, kS2gTps = 7;

int launch_s2g(const float *scatter, float *gather, i64 n, int kh, int kw, i64 h,
               i64 w, cudaStream_t st) {
  const bool tma_ok =
      !force_generic() && (w % 4 == 0) && w >= 4 &&
      (reinterpret_cast<uintptr_t>(scatter) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(gather) & 15) == 0 && w < (1ll << 31) - 512 &&
      h < (1ll << 31) - 512 && (i64)kh * kw < (1ll << 20) &&
      ceil_div(w, kTileW) * ceil_div(h, kS2gRows) * n < 0x7fffffffll &&
      (unsigned long long)w * h * kh * kw * 4ull < (1ull << 40);
  if (!tma_ok) {
    note_path(2);
    warn_generic("scatter2gather", 0, kh, kw, w);
    return generic_s2g(scatter, gather, n, kh, kw, h, w, st);
  }
  note_path(1);
  return run_s2g<kS2gRows, kS2gStages, kS2gTps>(scatter, gather, n, kh, kw, h, w, st);
}

}  // namespace sbmc
