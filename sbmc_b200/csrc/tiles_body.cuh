// tiles_body.cuh -- per-thread body of the tile assembly kernel (tiles.cu) and
// the host-side argument check that builds its parameter block.  Both are plain
// functions so that tests/tiles_emul.cpp can run the exact device arithmetic on
// the CPU, thread by thread, against the reference reader's output.
//
// Channel order of `features` (sbmc/datasets.py:317-354 labels, :700-711
// keep_idx): [coords 5] radiance 6 (diffuse rgb, specular rgb) [g-buffer 16]
// [p 4*depth] [ld 2*depth] [bt 5 flags x depth, flag-major (:676-680)].
#pragma once
#include <stdint.h>
#include <vector_types.h>

#include "../../include/sbmc_b200.h"

#if defined(__CUDACC__)
#define SBMC_HD __host__ __device__ __forceinline__
#else
#define SBMC_HD static inline
#include <math.h>
#endif

namespace sbmc {

void set_error(const char *fmt, ...);

#define SBMC_TILE_MAX_FEATURES 160
#define SBMC_TILE_BT_BASE 4096  // chan_src >= this: bounce-type flag plane

struct TileAssembleParams {
  const uint8_t *raw;
  const long long *tiles;  // [ntiles][4]: image frame offset, first sample frame offset, block_x, block_y
  long long ntiles, sample_stride, h, w, row0;
  int ts, spp, nf, nchans, pixel_features, depth;
  int float_planes;         // fp32 planes in a sample frame (sample_features + 6*depth)
  int i_diffuse;            // output channel of diffuse_r (specular_r = +3)
  int preprocess;           // 1: log-compress the radiance channels (sbmc mode)
  int aligned16;            // every base pointer / offset allows 16-byte accesses
  float *features, *radiance, *low_spp, *image_data, *image_data_var, *target_image;
  short chan_src[SBMC_TILE_MAX_FEATURES];   // per output channel: source plane / BT_BASE + k
  // the same selection regrouped for the kernel's two loops
  short fl_out[SBMC_TILE_MAX_FEATURES], fl_src[SBMC_TILE_MAX_FEATURES];  // plain fp32 planes
  int n_fl;                 // ... how many (radiance excluded: it has its own code)
  int i_bt;                 // output channel of the first bounce-type plane, -1: none
};

// Fills *p.  Returns 0 = launch, 1 = nothing to do, negative = SBMC_E*.
static inline int tile_assemble_params(TileAssembleParams *p, const void *raw,
                                       const int64_t *tile_table, int64_t ntiles,
                                       int64_t sample_stride_bytes, int ts, int spp,
                                       int sample_features, int pixel_features, int path_depth,
                                       int flags, float *features, float *radiance, float *low_spp,
                                       float *image_data, float *image_data_var,
                                       float *target_image, int64_t h, int64_t w,
                                       int64_t row0) {
  if (ntiles < 0 || ts < 1 || spp < 0 || h < 1 || w < ts || path_depth < 0 || pixel_features < 0) {
    set_error("tile assembly: invalid shape ntiles=%lld ts=%d spp=%d h=%lld w=%lld depth=%d",
              (long long)ntiles, ts, spp, (long long)h, (long long)w, path_depth);
    return SBMC_EINVAL;
  }
  if (sample_features != 27) {  // the reference's keep_idx ranges hard-code 27 (datasets.py:700-709)
    set_error("tile assembly: sample_features must be 27 (got %d)", sample_features);
    return SBMC_EINVAL;
  }
  if (ntiles == 0) return 1;
  if (!raw || !tile_table) {
    set_error("tile assembly: null input pointer");
    return SBMC_EINVAL;
  }
  const bool coords = flags & SBMC_TILE_COORDS, gbuf = flags & SBMC_TILE_GBUFFER,
             lp = flags & SBMC_TILE_P, ld = flags & SBMC_TILE_LD, bt = flags & SBMC_TILE_BT;
  int nf = 0;
  short *cs = p->chan_src;
  const int max_nf = 27 + 6 * path_depth + 5 * path_depth;
  if (max_nf > SBMC_TILE_MAX_FEATURES) {
    set_error("tile assembly: path depth %d too large", path_depth);
    return SBMC_EINVAL;
  }
  if (coords)
    for (int i = 0; i < 5; ++i) cs[nf++] = (short)i;
  p->i_diffuse = nf;
  for (int i = 5; i < 11; ++i) cs[nf++] = (short)i;
  if (gbuf)
    for (int i = 11; i < 27; ++i) cs[nf++] = (short)i;
  if (lp)
    for (int i = 0; i < 4 * path_depth; ++i) cs[nf++] = (short)(27 + i);
  if (ld)
    for (int i = 0; i < 2 * path_depth; ++i) cs[nf++] = (short)(27 + 4 * path_depth + i);
  if (bt)
    for (int i = 0; i < 5 * path_depth; ++i) cs[nf++] = (short)(SBMC_TILE_BT_BASE + i);
  const bool want_samples = spp > 0;
  const bool want_image = pixel_features > 0;
  if (want_samples && (!features || !radiance || !low_spp)) {
    set_error("tile assembly: null sample output pointer");
    return SBMC_EINVAL;
  }
  if (want_image && (!image_data || !image_data_var || !target_image)) {
    set_error("tile assembly: null image output pointer");
    return SBMC_EINVAL;
  }
  if (want_image && pixel_features / 2 < 6) {
    set_error("tile assembly: pixel_features %d leaves no diffuse / specular planes",
              pixel_features);
    return SBMC_EINVAL;
  }
  if (!want_samples && !want_image) return 1;
  const long long need = ((long long)(27 + 6 * path_depth) * 4 + (long long)path_depth * 2) * ts * ts;
  if (want_samples && sample_stride_bytes < need) {
    set_error("tile assembly: sample stride %lld smaller than a sample frame (%lld)",
              (long long)sample_stride_bytes, need);
    return SBMC_EINVAL;
  }
  p->n_fl = 0;
  p->i_bt = -1;
  for (int f = 0; f < nf; ++f) {
    if (f >= p->i_diffuse && f < p->i_diffuse + 6) continue;
    if (cs[f] >= SBMC_TILE_BT_BASE) {
      if (p->i_bt < 0) p->i_bt = f;
      continue;
    }
    p->fl_out[p->n_fl] = (short)f;
    p->fl_src[p->n_fl] = cs[f];
    ++p->n_fl;
  }
  p->raw = static_cast<const uint8_t *>(raw);
  p->tiles = reinterpret_cast<const long long *>(tile_table);
  p->ntiles = ntiles;
  p->sample_stride = sample_stride_bytes;
  p->h = h;
  p->w = w;
  p->row0 = row0;
  p->ts = ts;
  p->spp = spp;
  p->nf = nf;
  p->pixel_features = pixel_features;
  p->nchans = pixel_features / 2;
  p->depth = path_depth;
  p->float_planes = 27 + 6 * path_depth;
  p->preprocess = (flags & SBMC_TILE_LOG_RADIANCE) ? 1 : 0;
  p->features = features;
  p->radiance = radiance;
  p->low_spp = low_spp;
  p->image_data = image_data;
  p->image_data_var = image_data_var;
  p->target_image = target_image;
  uintptr_t bits = reinterpret_cast<uintptr_t>(raw) | (uintptr_t)sample_stride_bytes;
  const float *outs[] = {features, radiance, low_spp, image_data, image_data_var, target_image};
  for (const float *o : outs) bits |= reinterpret_cast<uintptr_t>(o);
  // Frame offsets and block_x come from the table; the caller promises 16-byte
  // frame offsets and block_x % 4 == 0 through SBMC_TILE_ALIGNED.
  p->aligned16 = ((bits & 15) == 0) && (flags & SBMC_TILE_ALIGNED);
  return 0;
}

template <int VEC>
struct Px {
  float v[VEC];
};

template <int VEC>
SBMC_HD Px<VEC> px_load(const float *q) {
  Px<VEC> r;
  if (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4 *>(q);
    r.v[0] = t.x;
    r.v[1 % VEC] = t.y;
    r.v[2 % VEC] = t.z;
    r.v[3 % VEC] = t.w;
  } else {
    for (int i = 0; i < VEC; ++i) r.v[i] = q[i];
  }
  return r;
}

template <int VEC>
SBMC_HD void px_store(float *q, const Px<VEC> &r) {
  if (VEC == 4) {
    float4 t;
    t.x = r.v[0];
    t.y = r.v[1 % VEC];
    t.z = r.v[2 % VEC];
    t.w = r.v[3 % VEC];
    *reinterpret_cast<float4 *>(q) = t;
  } else {
    for (int i = 0; i < VEC; ++i) q[i] = r.v[i];
  }
}

// np.maximum(x, 0) (datasets.py:760-761): NaN and -0.0 pass through.
SBMC_HD float max0(float x) { return x < 0.0f ? 0.0f : x; }

SBMC_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

// One thread: pixels [x, x+VEC) of row y of tile `tile`.
template <int VEC>
SBMC_HD void tile_assemble_body(const TileAssembleParams &p, long long tile, int y, int x) {
  const long long img_off = p.tiles[4 * tile + 0], smp_off = p.tiles[4 * tile + 1];
  const long long bx = p.tiles[4 * tile + 2], by = p.tiles[4 * tile + 3];
  const long long plane_in = (long long)p.ts * p.ts;     // elements
  const long long plane_out = p.h * p.w;                 // elements
  const long long in_px = (long long)y * p.ts + x;
  const long long oy = by + y - p.row0;                  // row in the (band of the) output
  if (oy < 0 || oy >= p.h) return;
  const long long out_px = oy * p.w + bx + x;

  // ---- pixel statistics (datasets.py:592-606) --------------------------------
  if (p.pixel_features > 0) {
    const float *img = reinterpret_cast<const float *>(p.raw + img_off);
    for (int c = 0; c < p.nchans; ++c) {
      px_store<VEC>(p.image_data + c * plane_out + out_px,
                    px_load<VEC>(img + c * plane_in + in_px));
      px_store<VEC>(p.image_data_var + c * plane_out + out_px,
                    px_load<VEC>(img + (long long)(p.nchans + c) * plane_in + in_px));
    }
    for (int c = 0; c < 3; ++c) {  // regression target = diffuse + specular means
      const Px<VEC> d = px_load<VEC>(img + c * plane_in + in_px);
      const Px<VEC> sp = px_load<VEC>(img + (3 + c) * plane_in + in_px);
      Px<VEC> t;
      for (int i = 0; i < VEC; ++i) t.v[i] = d.v[i] + sp.v[i];
      px_store<VEC>(p.target_image + c * plane_out + out_px, t);
    }
  }
  if (p.spp <= 0) return;

  // ---- samples (datasets.py:625-729, 744-778) ----------------------------------
  Px<VEC> acc[3];
  for (int s = 0; s < p.spp; ++s) {
    const uint8_t *frame = p.raw + smp_off + (long long)s * p.sample_stride;
    const float *fl = reinterpret_cast<const float *>(frame);
    const short *bt = reinterpret_cast<const short *>(frame + (long long)p.float_planes * plane_in * 4);
    float *fout = p.features + (long long)s * p.nf * plane_out + out_px;
    // Radiance planes 5..10 of the frame: diffuse rgb then specular rgb.  Emits
    // radiance = diffuse + specular (raw), its running sum over the samples, and
    // the two feature triplets (log-compressed in sbmc mode).
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 3; ++c) {
      const Px<VEC> d = px_load<VEC>(fl + (long long)(5 + c) * plane_in + in_px);
      const Px<VEC> sp = px_load<VEC>(fl + (long long)(8 + c) * plane_in + in_px);
      Px<VEC> rad, od, os;
      for (int i = 0; i < VEC; ++i) {
        rad.v[i] = d.v[i] + sp.v[i];
        if (p.preprocess) {
          const float dd = max0(d.v[i]), ss = max0(sp.v[i]);
          const float total = dd + ss;
          od.v[i] = div_rn(logf(1.0f + total), 10.0f);
          os.v[i] = div_rn(logf(1.0f + ss), 10.0f);
        } else {
          od.v[i] = d.v[i];
          os.v[i] = sp.v[i];
        }
        acc[c].v[i] = (s == 0) ? rad.v[i] : acc[c].v[i] + rad.v[i];
      }
      px_store<VEC>(p.radiance + ((long long)s * 3 + c) * plane_out + out_px, rad);
      px_store<VEC>(fout + (long long)(p.i_diffuse + c) * plane_out, od);
      px_store<VEC>(fout + (long long)(p.i_diffuse + 3 + c) * plane_out, os);
    }
    // Plain planes (coordinates, g-buffer, probabilities, light directions):
    // four loads in flight, then four stores.
    for (int k0 = 0; k0 < p.n_fl; k0 += 4) {
      Px<VEC> v[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 4; ++j)
        if (k0 + j < p.n_fl)
          v[j] = px_load<VEC>(fl + (long long)p.fl_src[k0 + j] * plane_in + in_px);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 4; ++j)
        if (k0 + j < p.n_fl) px_store<VEC>(fout + (long long)p.fl_out[k0 + j] * plane_out, v[j]);
    }
    // Bounce types: one int16 plane per path vertex -> five 0 / 1 planes
    // (reflection, transmission, diffuse, glossy, specular), flag-major.
    if (p.i_bt >= 0) {
      for (int vertex = 0; vertex < p.depth; ++vertex) {
        const short *q = bt + (long long)vertex * plane_in + in_px;
        int bits[VEC];
        if (VEC == 4) {
          const short4 t = *reinterpret_cast<const short4 *>(q);
          bits[0] = t.x;
          bits[1 % VEC] = t.y;
          bits[2 % VEC] = t.z;
          bits[3 % VEC] = t.w;
        } else {
          for (int i = 0; i < VEC; ++i) bits[i] = q[i];
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int flag = 0; flag < 5; ++flag) {
          Px<VEC> o;
          for (int i = 0; i < VEC; ++i) o.v[i] = (float)((bits[i] >> flag) & 1);
          px_store<VEC>(fout + (long long)(p.i_bt + flag * p.depth + vertex) * plane_out, o);
        }
      }
    }
  }
  const float count = (float)p.spp;
  for (int c = 0; c < 3; ++c) {
    Px<VEC> m;
    for (int i = 0; i < VEC; ++i) m.v[i] = div_rn(acc[c].v[i], count);
    px_store<VEC>(p.low_spp + c * plane_out + out_px, m);
  }
}

}  // namespace sbmc
