// capi.cu -- the extern "C" device-pointer entry points (include/sbmc_b200.h).
#include "common.cuh"

using sbmc::i64;

static int check_common(i64 n, int c, i64 h, i64 w, int kh, int kw) {
  if (n < 0 || h < 0 || w < 0 || c < 1 || kh < 1 || kw < 1) {
    sbmc::set_error("invalid shape n=%lld c=%d h=%lld w=%lld kh=%d kw=%d",
                    (long long)n, c, (long long)h, (long long)w, kh, kw);
    return SBMC_EINVAL;
  }
  return SBMC_OK;
}

static int check_ptrs(const void *const *ptrs, int count) {
  for (int i = 0; i < count; ++i) {
    if (!ptrs[i]) {
      sbmc::set_error("null pointer argument (#%d)", i);
      return SBMC_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 3) {
      sbmc::set_error("pointer argument #%d is not 4-byte aligned", i);
      return SBMC_EALIGN;
    }
  }
  return SBMC_OK;
}

extern "C" {

int sbmc_scatter2gather_f32(const float *scatter, float *gather, int64_t n,
                            int kh, int kw, int64_t h, int64_t w, void *stream) {
  int rc = check_common(n, 1, h, w, kh, kw);
  if (rc) return rc;
  if (n == 0 || h == 0 || w == 0) return SBMC_OK;
  const void *ptrs[] = {scatter, gather};
  if ((rc = check_ptrs(ptrs, 2))) return rc;
  if (scatter == gather) {
    sbmc::set_error("scatter2gather cannot run in place");
    return SBMC_EINVAL;
  }
  return sbmc::launch_s2g(scatter, gather, n, kh, kw, h, w,
                          static_cast<cudaStream_t>(stream));
}

int sbmc_kernel_weighting_fwd_band_f32(const float *data_ext,
                                       const float *weights, float *output,
                                       float *sum_w, int64_t n, int c, int64_t h,
                                       int64_t w, int kh, int kw, int halo_top,
                                       int halo_bot, void *stream) {
  int rc = check_common(n, c, h, w, kh, kw);
  if (rc) return rc;
  if (halo_top < 0 || halo_bot < 0) {
    sbmc::set_error("negative halo");
    return SBMC_EINVAL;
  }
  if (n == 0 || h == 0 || w == 0) return SBMC_OK;
  const void *ptrs[] = {data_ext, weights, output, sum_w};
  if ((rc = check_ptrs(ptrs, 4))) return rc;
  return sbmc::launch_fwd(data_ext, weights, output, sum_w, n, c, h, w, kh, kw,
                          halo_top, halo_bot, static_cast<cudaStream_t>(stream));
}

int sbmc_kernel_weighting_fwd_f32(const float *data, const float *weights,
                                  float *output, float *sum_w, int64_t n, int c,
                                  int64_t h, int64_t w, int kh, int kw,
                                  void *stream) {
  return sbmc_kernel_weighting_fwd_band_f32(data, weights, output, sum_w, n, c, h,
                                            w, kh, kw, 0, 0, stream);
}

int sbmc_kernel_weighting_bwd_band_f32(const float *data_ext,
                                       const float *weights,
                                       const float *d_output,
                                       const float *d_sum_w, float *d_data_ext,
                                       float *d_weights, int64_t n, int c,
                                       int64_t h, int64_t w, int kh, int kw,
                                       int halo_top, int halo_bot, void *stream) {
  int rc = check_common(n, c, h, w, kh, kw);
  if (rc) return rc;
  if (halo_top < 0 || halo_bot < 0) {
    sbmc::set_error("negative halo");
    return SBMC_EINVAL;
  }
  if (n == 0 || w == 0) return SBMC_OK;
  const void *ptrs[] = {data_ext, weights, d_output, d_sum_w, d_data_ext, d_weights};
  if (h == 0) {  // nothing scatters into the halo rows: they are zero
    if (halo_top + halo_bot == 0) return SBMC_OK;
    if (!d_data_ext) return SBMC_EINVAL;
    SBMC_CUDA_OK(cudaMemsetAsync(
        d_data_ext, 0, sizeof(float) * (size_t)(n * c * (halo_top + halo_bot) * w),
        static_cast<cudaStream_t>(stream)));
    return SBMC_OK;
  }
  if ((rc = check_ptrs(ptrs, 6))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = sbmc::launch_bwd_dweights(data_ext, d_output, d_sum_w, d_weights, n, c, h,
                                 w, kh, kw, halo_top, halo_bot, st);
  if (rc) return rc;
  return sbmc::launch_bwd_ddata(weights, d_output, d_data_ext, n, c, h, w, kh, kw,
                                halo_top, halo_bot, st);
}

int sbmc_kernel_weighting_bwd_f32(const float *data, const float *weights,
                                  const float *sum_w, const float *d_output,
                                  const float *d_sum_w, float *d_data,
                                  float *d_weights, int64_t n, int c, int64_t h,
                                  int64_t w, int kh, int kw, void *stream) {
  (void)sum_w;  // unused by the reference pipeline as well
  return sbmc_kernel_weighting_bwd_band_f32(data, weights, d_output, d_sum_w,
                                            d_data, d_weights, n, c, h, w, kh, kw,
                                            0, 0, stream);
}

}  // extern "C"
