/* cabi_client.c -- a plain C99 client of include/sbmc_b200.h (built and run by
 * tests/test_capi.py): the header must compile as C, every entry point must link,
 * and argument validation must work without touching a device. */
#include <stdio.h>
#include <string.h>

#include "sbmc_b200.h"

int main(void) {
  int failures = 0;
  float dummy[4] = {0.f, 0.f, 0.f, 0.f};
  /* take the address of every entry point so that the linker must resolve it */
  typedef void (*entry_fn)(void);
  const entry_fn entry_points[] = {
      (entry_fn)sbmc_b200_version,
      (entry_fn)sbmc_b200_last_error,
      (entry_fn)sbmc_b200_force_generic,
      (entry_fn)sbmc_b200_last_path,
      (entry_fn)sbmc_b200_launch_count,
      (entry_fn)sbmc_b200_timing_enable,
      (entry_fn)sbmc_b200_timing_collect,
      (entry_fn)sbmc_scatter2gather_f32,
      (entry_fn)sbmc_kernel_weighting_fwd_f32,
      (entry_fn)sbmc_kernel_weighting_bwd_f32,
      (entry_fn)sbmc_progressive_splat_fwd_f32,
      (entry_fn)sbmc_progressive_splat_bwd_f32,
      (entry_fn)sbmc_conv1x1_chain_f32,
      (entry_fn)sbmc_conv1x1_chain_nhwc_bf16,
      (entry_fn)sbmc_chain_samples_nhwc_bf16,
      (entry_fn)sbmc_conv3x3_nhwc_bf16,
      (entry_fn)sbmc_b200_conv3x3_pair,
      (entry_fn)sbmc_b200_conv3x3_linear,
      (entry_fn)sbmc_maxpool2x2_nhwc_bf16,
      (entry_fn)sbmc_linear_nhwc_bf16,
      (entry_fn)sbmc_linear2_nhwc_bf16,
      (entry_fn)sbmc_wgrad_nhwc_bf16,
      (entry_fn)sbmc_wgrad3x3_nhwc_bf16,
      (entry_fn)sbmc_weight_bank_run,
      (entry_fn)sbmc_conv3x3_masked_nhwc_bf16,
      (entry_fn)sbmc_spp_reduce_nhwc_bf16,
      (entry_fn)sbmc_bcast_add_nhwc_bf16,
      (entry_fn)sbmc_maxpool2x2_bwd_nhwc_bf16,
      (entry_fn)sbmc_upsample_bwd_nhwc_bf16,
      (entry_fn)sbmc_dact_bf16,
      (entry_fn)sbmc_colsum_bf16,
      (entry_fn)sbmc_upsample_concat_nhwc_bf16,
      (entry_fn)sbmc_bias_act_nhwc_bf16,
      (entry_fn)sbmc_nchw_to_nhwc_bf16,
      (entry_fn)sbmc_multi_tensor_grad_norm_f32,
      (entry_fn)sbmc_multi_tensor_adam_f32,
      (entry_fn)sbmc_multi_tensor_adam_devstep_f32,
      (entry_fn)sbmc_lz4_frames_inflate,
      (entry_fn)sbmc_tile_assemble_f32,
      (entry_fn)sbmc_kernel_weighting_fwd_band_f32,
      (entry_fn)sbmc_kernel_weighting_bwd_band_f32,
      (entry_fn)sbmc_scatter2gather_host_f32,
      (entry_fn)sbmc_kernel_weighting_fwd_host_f32,
      (entry_fn)sbmc_kernel_weighting_bwd_host_f32,
      (entry_fn)sbmc_b200_host_release,
  };
  size_t i;
  for (i = 0; i < sizeof(entry_points) / sizeof(entry_points[0]); ++i)
    if (!entry_points[i]) ++failures;

  if (sbmc_b200_version() < 100) ++failures;
  /* invalid shape: rejected before any CUDA call, with a message */
  if (sbmc_kernel_weighting_fwd_f32(dummy, dummy, dummy, dummy, 1, 0, 4, 4, 3, 3, NULL) !=
      SBMC_EINVAL)
    ++failures;
  if (strstr(sbmc_b200_last_error(), "invalid shape") == NULL) ++failures;
  /* empty problem: success, nothing touched */
  if (sbmc_scatter2gather_f32(NULL, NULL, 0, 3, 3, 8, 8, NULL) != SBMC_OK) ++failures;
  /* null pointers on a non-empty problem */
  if (sbmc_kernel_weighting_bwd_f32(NULL, NULL, NULL, NULL, NULL, NULL, NULL, 1, 3, 4, 4, 3, 3,
                                    NULL) != SBMC_EINVAL)
    ++failures;
  if (sbmc_progressive_splat_fwd_f32(NULL, NULL, NULL, NULL, NULL, 1, 3, 4, 4, 3, 3, 1, 1,
                                     NULL) != SBMC_EINVAL)
    ++failures;
  printf("entry points: %d, failures: %d\n", (int)(sizeof(entry_points) / sizeof(entry_points[0])),
         failures);
  return failures;
}
