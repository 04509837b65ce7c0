// s2g.cu -- Scatter2Gather as a pure TMA copy engine program.
//
// gather[n][dy][dx][y][x] = scatter[n][KH-1-dy][KW-1-dx][y+dy-c0h][x+dx-c0w],
// zero when the source pixel is outside the image
// (reference: src/scatter2gather.cpp:34-47).  That is K*K*N independent shifted
// 2-D plane copies with zero fill, so the kernel has no arithmetic at all: one
// elected thread per CTA walks a list of (box, tap, n) jobs, pulls each source
// box with a TMA tiled load at the SHIFTED coordinates (out-of-bounds elements,
// including negative coordinates, arrive as 0.0f -- the boundary condition for
// free), and pushes the same shared-memory buffer back out with a TMA tiled
// store at the aligned destination coordinates (out-of-bounds part clipped).
// A ring of STAGES buffers keeps several loads and stores in flight per CTA.
// Bit-exact by construction (bytes are only moved).
#include "common.cuh"

namespace sbmc {

struct S2GJobs {
  int kh, kw, bx, by;          // kernel size, box size
  int nxb, nyb;                // boxes per row / column
  long long njobs;             // nxb * nyb * kh * kw * n
};

template <int STAGES>
__global__ void __launch_bounds__(32)
s2g_tma_kernel(const __grid_constant__ CUtensorMap smap,
               const __grid_constant__ CUtensorMap gmap, const S2GJobs J) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[STAGES];
  if (threadIdx.x != 0) return;  // a single thread drives the copy engine

  const uint32_t box_bytes = (uint32_t)(J.bx * J.by * sizeof(float));
  const int c0h = (J.kh - 1) / 2, c0w = (J.kw - 1) / 2;
  const long long taps = (long long)J.kh * J.kw;
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < J.njobs ? (J.njobs - first + step - 1) / step : 0;

  for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
  fence_mbar_init();
  prefetch_tensormap(&smap);
  prefetch_tensormap(&gmap);

  auto coords = [&](long long k, int &x, int &y, int &tap, int &n) {
    long long job = first + k * step;
    x = (int)(job % J.nxb) * J.bx;
    job /= J.nxb;
    y = (int)(job % J.nyb) * J.by;
    job /= J.nyb;
    tap = (int)(job % taps);
    n = (int)(job / taps);
  };
  auto issue = [&](long long k) {
    int x, y, tap, n;
    coords(k, x, y, tap, n);
    const int dy = tap / J.kw, dx = tap % J.kw;
    const int s = (int)(k % STAGES);
    mbar_expect_tx(&bars[s], box_bytes);
    tma_load_4d(smem_raw + (size_t)s * box_bytes, &smap, &bars[s], x + dx - c0w,
                y + dy - c0h, (J.kh - 1 - dy) * J.kw + (J.kw - 1 - dx), n);
  };

  constexpr int AHEAD = STAGES - 1;
  for (long long k = 0; k < AHEAD && k < mine; ++k) issue(k);
  for (long long k = 0; k < mine; ++k) {
    const int s = (int)(k % STAGES);
    mbar_wait(&bars[s], (uint32_t)((k / STAGES) & 1));
    fence_proxy_async();
    int x, y, tap, n;
    coords(k, x, y, tap, n);
    tma_store_4d(&gmap, smem_raw + (size_t)s * box_bytes, x, y, tap, n);
    tma_commit_group();
    if (k + AHEAD < mine) {
      // the stage about to be refilled was last read by the store of job k-1:
      // allow only the newest store (job k) to still be reading shared memory.
      tma_wait_group_read<1>();
      issue(k + AHEAD);
    }
  }
  tma_wait_group<0>();
}

static bool make_plane_map(CUtensorMap *map, const float *base, i64 n, i64 taps,
                           i64 h, i64 w, int bx, int by) {
  const uint64_t dims[4] = {(uint64_t)w, (uint64_t)h, (uint64_t)taps, (uint64_t)n};
  const uint64_t strides[3] = {(uint64_t)w * 4, (uint64_t)w * h * 4,
                               (uint64_t)w * h * taps * 4};
  const uint32_t box[4] = {(uint32_t)bx, (uint32_t)by, 1u, 1u};
  return encode_tensor_map_f32(map, base, 4, dims, strides, box);
}

// Tuning knobs (exposed for the sweep tool through launch_s2g_cfg).
int launch_s2g_cfg(const float *scatter, float *gather, i64 n, int kh, int kw,
                   i64 h, i64 w, int bx, int by, int stages, int ctas_per_sm,
                   cudaStream_t st) {
  const i64 taps = (i64)kh * kw;
  CUtensorMap smap, gmap;
  if (!make_plane_map(&smap, scatter, n, taps, h, w, bx, by) ||
      !make_plane_map(&gmap, gather, n, taps, h, w, bx, by))
    return SBMC_ECUDA;
  S2GJobs J;
  J.kh = kh; J.kw = kw; J.bx = bx; J.by = by;
  J.nxb = (int)ceil_div(w, bx);
  J.nyb = (int)ceil_div(h, by);
  J.njobs = (long long)J.nxb * J.nyb * taps * n;
  const size_t smem = (size_t)stages * bx * by * sizeof(float);
  i64 grid = (i64)num_sms() * ctas_per_sm;
  if (grid > J.njobs) grid = J.njobs;
  if (grid < 1) grid = 1;
#define SBMC_S2G(S)                                                            \
  case S: {                                                                    \
    auto kern = s2g_tma_kernel<S>;                                             \
    SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<(unsigned)grid, 32, smem, st>>>(smap, gmap, J);                     \
  } break;
  KernelTimer timer(SBMC_KERNEL_S2G, st);
  switch (stages) {
    SBMC_S2G(2) SBMC_S2G(3) SBMC_S2G(4) SBMC_S2G(6) SBMC_S2G(8)
    default:
      set_error("s2g: unsupported stage count %d", stages);
      return SBMC_EINVAL;
  }
#undef SBMC_S2G
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

int launch_s2g(const float *scatter, float *gather, i64 n, int kh, int kw, i64 h,
               i64 w, cudaStream_t st) {
  const bool tma_ok =
      !force_generic() && (w % 4 == 0) && w >= 4 &&
      (reinterpret_cast<uintptr_t>(scatter) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(gather) & 15) == 0 && w < (1ll << 31) - 512 &&
      h < (1ll << 31) - 512 && (i64)kh * kw * n < (1ll << 31) &&
      (unsigned long long)w * h * kh * kw * 4ull < (1ull << 40);
  if (!tma_ok) {
    note_path(2);
    return generic_s2g(scatter, gather, n, kh, kw, h, w, st);
  }
  note_path(1);
  // box: up to 256 x 16 floats (16 KB); 4 stages, 3 CTAs per SM
  int bx = (int)(w < 256 ? w : 256);
  int by = (int)(h < 16 ? h : 16);
  return launch_s2g_cfg(scatter, gather, n, kh, kw, h, w, bx, by, 4, 3, st);
}

}  // namespace sbmc
