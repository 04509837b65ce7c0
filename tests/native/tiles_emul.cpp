// tiles_emul.cpp -- runs the DEVICE code of the tile reader on the CPU.
//
// sbmc_b200/csrc/lz4_warp.cuh and tiles_body.cuh are written so that the same
// source compiles for the host: the 32 lanes of the warp inflater run one after
// the other, and the assembly kernel's per-thread body is called for every
// thread index of its grid.  This checks the cursor / index arithmetic of the
// kernels against the oracle and the reference reader's fixtures on machines
// without a GPU (tests/test_tiles.py, `-m "not gpu"`).  It is test code: the
// product never loads it.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../sbmc_b200/csrc/lz4_warp.cuh"
#include "../../sbmc_b200/csrc/tiles_body.cuh"

namespace sbmc {
static char g_err[512];
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace sbmc

extern "C" {

const char *emul_last_error() { return sbmc::g_err; }

// lz4_frames_kernel, one "warp" per frame.
int emul_lz4_frames_inflate(const uint8_t *src, const int64_t *table, int64_t nframes, uint8_t *dst,
                            int32_t *status) {
  for (int64_t f = 0; f < nframes; ++f) {
    int64_t out_len = 0;
    int rc = sbmc::lz4::decode_frames(src + table[4 * f], table[4 * f + 1], dst + table[4 * f + 2],
                                      table[4 * f + 3], &out_len);
    if (rc == sbmc::lz4::kOk && out_len != table[4 * f + 3]) rc = sbmc::lz4::kSizeMismatch;
    status[f] = rc;
  }
  return 0;
}

// tile_assemble_kernel<VEC>: same argument check, same grid decomposition.
// Returns the VEC used (4 or 1) or a negative SBMC_E* code.
int emul_tile_assemble_f32(const void *raw, const int64_t *tile_table, int64_t ntiles,
                           int64_t sample_stride_bytes, int ts, int spp, int sample_features,
                           int pixel_features, int path_depth, int flags, float *features,
                           float *radiance, float *low_spp, float *image_data,
                           float *image_data_var, float *target_image, int64_t h, int64_t w,
                           int64_t row0) {
  sbmc::TileAssembleParams p;
  int rc = sbmc::tile_assemble_params(&p, raw, tile_table, ntiles, sample_stride_bytes, ts, spp,
                                      sample_features, pixel_features, path_depth, flags, features,
                                      radiance, low_spp, image_data, image_data_var, target_image,
                                      h, w, row0);
  if (rc == 1) return 0;
  if (rc < 0) return rc;
  const bool vec4 = (ts % 4 == 0) && (w % 4 == 0) && p.aligned16;
  const long long per_row = vec4 ? ts / 4 : ts;
  const long long per_tile = per_row * ts;
  for (long long gid = 0; gid < per_tile * ntiles; ++gid) {
    const long long tile = gid / per_tile;
    const long long rem = gid - tile * per_tile;
    const int y = (int)(rem / per_row);
    const int x = (int)(rem - (long long)y * per_row) * (vec4 ? 4 : 1);
    if (vec4)
      sbmc::tile_assemble_body<4>(p, tile, y, x);
    else
      sbmc::tile_assemble_body<1>(p, tile, y, x);
  }
  return vec4 ? 4 : 1;
}

}  // extern "C"
