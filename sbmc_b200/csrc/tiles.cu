// tiles.cu -- the sample-buffer reader of the denoiser's callers, on the GPU:
// compressed tile bytes cross PCIe, are inflated in HBM (one warp per LZ4
// frame, lz4_warp.cuh) and are assembled into the tensors the model consumes.
//
// Reference: sbmc/datasets.py:570-739 (`TilesDataset._read_compressed`,
// `_read_data`), :744-778 (`_preprocess_standard`) and :920-957
// (`FullImagesDataset.__getitem__`, tiles pasted at (block_y, block_x)).
//
// Layout of one inflated tile (what the renderer wrote, planar [chan][y][x]):
//   image frame   : pixel_features planes of ts*ts fp32 (means, then variances)
//   sample frame s: sample_features (27) fp32 planes, 4*depth probability
//                   planes, 2*depth light-direction planes, depth int16 planes
//                   of bounce-type bit flags
// Outputs are planar fp32 images [..][H][W]; a tile lands at (block_y, block_x).
// Pure HBM streaming: one thread owns VEC consecutive pixels of a tile row and
// walks samples and channels, every access a coalesced 4*VEC-byte plane read /
// write.  Sums over the samples run in sample order (numpy's reduction order).
#include "common.cuh"
#include "lz4_warp.cuh"
#include "tiles_body.cuh"

namespace sbmc {

__global__ void __launch_bounds__(32) lz4_frames_kernel(const uint8_t *__restrict__ src,
                                                        const i64 *__restrict__ table, i64 nframes,
                                                        uint8_t *dst, int *__restrict__ status) {
  const i64 f = blockIdx.x;
  if (f >= nframes) return;
  const i64 src_off = table[4 * f + 0], src_len = table[4 * f + 1];
  const i64 dst_off = table[4 * f + 2], dst_len = table[4 * f + 3];
  int64_t out_len = 0;
  int rc = lz4::decode_frames(src + src_off, src_len, dst + dst_off, dst_len, &out_len);
  if (rc == lz4::kOk && out_len != dst_len) rc = lz4::kSizeMismatch;
  if ((threadIdx.x & 31) == 0) status[f] = rc;
}

template <int VEC>
__global__ void __launch_bounds__(256) tile_assemble_kernel(const TileAssembleParams p) {
  const i64 per_row = p.ts / VEC;
  const i64 per_tile = per_row * p.ts;
  const i64 gid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= per_tile * p.ntiles) return;
  const i64 tile = gid / per_tile;
  const i64 rem = gid - tile * per_tile;
  const int y = (int)(rem / per_row);
  const int x = (int)(rem - (i64)y * per_row) * VEC;
  tile_assemble_body<VEC>(p, tile, y, x);
}

}  // namespace sbmc

using sbmc::i64;

extern "C" {

int sbmc_lz4_frames_inflate(const void *src, const int64_t *frame_table, int64_t nframes,
                            void *dst, int32_t *status, void *stream) {
  if (nframes < 0) {
    sbmc::set_error("negative frame count");
    return SBMC_EINVAL;
  }
  if (nframes == 0) return SBMC_OK;
  if (!src || !frame_table || !dst || !status) {
    sbmc::set_error("null pointer argument");
    return SBMC_EINVAL;
  }
  if (nframes > 0x7FFFFFFF) {
    sbmc::set_error("too many frames for one launch (%lld)", (long long)nframes);
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    sbmc::KernelTimer timer(SBMC_KERNEL_TILES, st);
    sbmc::lz4_frames_kernel<<<(unsigned)nframes, 32, 0, st>>>(
        static_cast<const uint8_t *>(src), reinterpret_cast<const i64 *>(frame_table), nframes,
        static_cast<uint8_t *>(dst), status);
  }
  SBMC_CUDA_OK(cudaGetLastError());
  sbmc::count_launch();
  sbmc::note_path(1);
  return SBMC_OK;
}

int sbmc_tile_assemble_f32(const void *raw, const int64_t *tile_table, int64_t ntiles,
                           int64_t sample_stride_bytes, int ts, int spp, int sample_features,
                           int pixel_features, int path_depth, int flags, float *features,
                           float *radiance, float *low_spp, float *image_data,
                           float *image_data_var, float *target_image, int64_t h, int64_t w,
                           int64_t row0, void *stream) {
  sbmc::TileAssembleParams p;
  int rc = sbmc::tile_assemble_params(&p, raw, tile_table, ntiles, sample_stride_bytes, ts, spp,
                                      sample_features, pixel_features, path_depth, flags, features,
                                      radiance, low_spp, image_data, image_data_var, target_image,
                                      h, w, row0);
  if (rc == 1) return SBMC_OK;  // nothing to do
  if (rc < 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec4 = (ts % 4 == 0) && (w % 4 == 0) && p.aligned16;
  const i64 threads = (i64)ntiles * ts * (vec4 ? ts / 4 : ts);
  const i64 blocks = sbmc::ceil_div(threads, 256);
  if (blocks > 0x7FFFFFFF) {
    sbmc::set_error("tile assembly: too many tiles for one launch");
    return SBMC_EINVAL;
  }
  {
    sbmc::KernelTimer timer(SBMC_KERNEL_TILES, st);
    if (vec4)
      sbmc::tile_assemble_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(p);
    else
      sbmc::tile_assemble_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(p);
  }
  SBMC_CUDA_OK(cudaGetLastError());
  sbmc::count_launch();
  sbmc::note_path(vec4 ? 1 : 2);
  return SBMC_OK;
}

}  // extern "C"
