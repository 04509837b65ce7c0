// kw_launch.cuh -- host launch templates for the tuned KernelWeighting kernels
// (shared by the library dispatch in kw_launch.cu and the tuning sweep tool).
#pragma once
#include "kw_kernels.cuh"

namespace sbmc {

// Tensor map over data_ext[n][c][hext][w] with box (TWS, rows, C, 1).
template <int KW>
static bool make_image_map(CUtensorMap *map, const float *data_ext, i64 n, int c,
                           i64 hext, i64 w, int rows) {
  using G = TileGeom<KW>;
  const uint64_t dims[4] = {(uint64_t)w, (uint64_t)hext, (uint64_t)c, (uint64_t)n};
  const uint64_t strides[3] = {(uint64_t)w * 4, (uint64_t)w * hext * 4,
                               (uint64_t)w * hext * c * 4};
  const uint32_t box[4] = {(uint32_t)G::TWS, (uint32_t)rows, (uint32_t)c, 1u};
  return encode_tensor_map_f32(map, data_ext, 4, dims, strides, box);
}

template <int C, int KW, int ROWS>
static inline size_t tile_smem_bytes(int kh) {
  using G = TileGeom<KW>;
  const size_t tile = (size_t)C * (ROWS + kh - 1) * G::TWS * sizeof(float);
  return ((tile + 15) & ~(size_t)15) + 16;
}

// Whether the tuned tile kernels can take this shape at all.
template <int C, int KW, int ROWS>
static inline bool tile_shape_ok(i64 n, i64 h, i64 w, int kh, i64 hext) {
  if (w % 4 != 0 || w > (1ll << 31) - 256 || h > (1ll << 31) - 256) return false;
  if (ROWS + kh - 1 > 256) return false;
  if (tile_smem_bytes<C, KW, ROWS>(kh) > 200 * 1024) return false;
  const i64 tiles = ceil_div(w, kTileW) * ceil_div(h, ROWS) * n;
  if (tiles > 0x7fffffffll) return false;
  if ((unsigned long long)w * hext * C * 4ull >= (1ull << 40)) return false;
  return true;
}

template <int C, int KW, int ROWS, int MINB, int CH>
int run_fwd(const float *data_ext, const float *weights, float *output,
            float *sum_w, i64 n, i64 h, i64 w, int kh, int halo_top,
            int halo_bot, cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  CUtensorMap dmap;
  if (!make_image_map<KW>(&dmap, data_ext, n, C, hext, w, ROWS + kh - 1))
    return SBMC_ECUDA;
  const size_t smem = tile_smem_bytes<C, KW, ROWS>(kh);
  auto kern = kw_fwd_kernel<C, KW, ROWS, MINB, CH>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int xt = (int)ceil_div(w, kTileW), yt = (int)ceil_div(h, ROWS);
  const unsigned grid = (unsigned)((i64)xt * yt * n);
  {
    KernelTimer timer(SBMC_KERNEL_KW_FWD, st);
    kern<<<grid, ROWS * 32, smem, st>>>(dmap, weights, output, sum_w, (int)h, (int)w,
                                        kh, halo_top, xt, yt);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

template <int C, int KW, int ROWS, int MINB, int CH, int STORE = 0>
int run_bwd_dweights(const float *data_ext, const float *d_output,
                     const float *d_sum_w, float *d_weights, i64 n, i64 h, i64 w,
                     int kh, int halo_top, int halo_bot, cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  CUtensorMap dmap;
  if (!make_image_map<KW>(&dmap, data_ext, n, C, hext, w, ROWS + kh - 1))
    return SBMC_ECUDA;
  const size_t smem = tile_smem_bytes<C, KW, ROWS>(kh);
  auto kern = kw_bwd_dweights_kernel<C, KW, ROWS, MINB, CH, STORE>;
  SBMC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int xt = (int)ceil_div(w, kTileW), yt = (int)ceil_div(h, ROWS);
  const unsigned grid = (unsigned)((i64)xt * yt * n);
  {
    KernelTimer timer(SBMC_KERNEL_KW_DWEIGHTS, st);
    kern<<<grid, ROWS * 32, smem, st>>>(dmap, d_output, d_sum_w, d_weights, (int)h,
                                        (int)w, kh, halo_top, xt, yt);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

template <int C, int KW, int NSEG, int MINB, int CH>
int run_bwd_ddata(const float *weights, const float *d_output, float *d_data_ext,
                  i64 n, i64 h, i64 w, int kh, int halo_top, int halo_bot,
                  cudaStream_t st) {
  const i64 hext = h + halo_top + halo_bot;
  const int xt = (int)ceil_div(w, (i64)NSEG * kTileW);
  const i64 ctas = (i64)xt * hext * n;
  if (ctas > 0x7fffffffll) {
    set_error("d_data grid too large");
    return SBMC_EINVAL;
  }
  {
    KernelTimer timer(SBMC_KERNEL_KW_DDATA, st);
    if (xt > 1)  // tile seams accumulate with atomics onto zero
      SBMC_CUDA_OK(cudaMemsetAsync(d_data_ext, 0, sizeof(float) * (size_t)(n * C * hext * w), st));
    kw_bwd_ddata_kernel<C, KW, NSEG, MINB, CH><<<(unsigned)ctas, NSEG * 32, 0, st>>>(
        weights, d_output, d_data_ext, (int)h, (int)w, kh, halo_top, (int)hext, xt);
  }
  count_launch();
  SBMC_CUDA_OK(cudaGetLastError());
  return SBMC_OK;
}

}  // namespace sbmc
