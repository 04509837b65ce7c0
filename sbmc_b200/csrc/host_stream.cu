// host_stream.cu -- host-buffer entry points (replace the *_cpu_float32 ops).
//
// The K*K weight volume is ~150x larger than the image it filters, so with host
// buffers the job is a PCIe stream: the weight volume is cut into row bands and
// a three-stream pipeline keeps the H2D copy of band b+1, the band kernels of
// band b and the D2H copy of band b-1 in flight at the same time (full-duplex
// PCIe).  The image-sized `data` tensor is uploaded once per image and stays
// whole on the device; a band addresses it through the row-band launchers
// (halo_top = y0, halo_bot = H - y1), the same launchers that serve multi-GPU
// H-sharding.  There is no CPU arithmetic in this file.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace sbmc {

static constexpr int kSlots = 3;
static constexpr int kBufs = 24;

struct HostPipe {
  int device = -1;
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  cudaEvent_t up[kSlots] = {}, done[kSlots] = {}, down[kSlots] = {};
  cudaEvent_t img_done = nullptr;  // last kernel of the previous image
  void *buf[kBufs] = {};
  size_t cap[kBufs] = {};
};

static HostPipe g_pipe;
static std::mutex g_pipe_mu;

static int pipe_release_locked() {
  HostPipe &P = g_pipe;
  if (P.device < 0) return SBMC_OK;
  cudaSetDevice(P.device);
  cudaDeviceSynchronize();
  for (int i = 0; i < kBufs; ++i) {
    if (P.buf[i]) cudaFree(P.buf[i]);
    P.buf[i] = nullptr;
    P.cap[i] = 0;
  }
  for (int i = 0; i < kSlots; ++i) {
    if (P.up[i]) cudaEventDestroy(P.up[i]);
    if (P.done[i]) cudaEventDestroy(P.done[i]);
    if (P.down[i]) cudaEventDestroy(P.down[i]);
    P.up[i] = P.done[i] = P.down[i] = nullptr;
  }
  if (P.img_done) cudaEventDestroy(P.img_done);
  P.img_done = nullptr;
  if (P.s_h2d) cudaStreamDestroy(P.s_h2d);
  if (P.s_comp) cudaStreamDestroy(P.s_comp);
  if (P.s_d2h) cudaStreamDestroy(P.s_d2h);
  P.s_h2d = P.s_comp = P.s_d2h = nullptr;
  P.device = -1;
  return SBMC_OK;
}

static int pipe_init_locked(int device) {
  HostPipe &P = g_pipe;
  if (P.device == device) {
    SBMC_CUDA_OK(cudaSetDevice(device));
    return SBMC_OK;
  }
  pipe_release_locked();
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    set_error("no CUDA device available");
    return SBMC_ENODEV;
  }
  if (device < 0 || device >= count) {
    set_error("device %d out of range (have %d)", device, count);
    return SBMC_EINVAL;
  }
  SBMC_CUDA_OK(cudaSetDevice(device));
  SBMC_CUDA_OK(cudaStreamCreateWithFlags(&P.s_h2d, cudaStreamNonBlocking));
  SBMC_CUDA_OK(cudaStreamCreateWithFlags(&P.s_comp, cudaStreamNonBlocking));
  SBMC_CUDA_OK(cudaStreamCreateWithFlags(&P.s_d2h, cudaStreamNonBlocking));
  for (int i = 0; i < kSlots; ++i) {
    SBMC_CUDA_OK(cudaEventCreateWithFlags(&P.up[i], cudaEventDisableTiming));
    SBMC_CUDA_OK(cudaEventCreateWithFlags(&P.done[i], cudaEventDisableTiming));
    SBMC_CUDA_OK(cudaEventCreateWithFlags(&P.down[i], cudaEventDisableTiming));
  }
  SBMC_CUDA_OK(cudaEventCreateWithFlags(&P.img_done, cudaEventDisableTiming));
  P.device = device;
  return SBMC_OK;
}

static int ensure(int idx, size_t bytes, float **out) {
  HostPipe &P = g_pipe;
  if (bytes < 256) bytes = 256;
  if (P.cap[idx] < bytes) {
    if (P.buf[idx]) SBMC_CUDA_OK(cudaFree(P.buf[idx]));
    P.buf[idx] = nullptr;
    P.cap[idx] = 0;
    SBMC_CUDA_OK(cudaMalloc(&P.buf[idx], bytes));
    P.cap[idx] = bytes;
  }
  *out = static_cast<float *>(P.buf[idx]);
  return SBMC_OK;
}

// rows per band: ~256 MB of weights per band (SBMC_HOST_BAND_MB overrides; measured
// 14.1 vs 13.6 Msamples/s end to end against 64 MB, profiles/r1y_e2e_band.txt), a
// multiple of 8 rows
static i64 band_rows(i64 h, i64 w, i64 taps) {
  static const i64 band_mb = [] {
    const char *e = getenv("SBMC_HOST_BAND_MB");
    const long v = e ? atol(e) : 0;
    return (i64)(v > 0 && v <= 4096 ? v : 256);
  }();
  i64 hb = (band_mb << 20) / (taps * w * 4);
  hb = hb / 8 * 8;
  if (hb < 8) hb = 8;
  if (hb > h) hb = h;
  return hb;
}

// dst[c][dst_row0 + r][x] += src[c][r][x]   (src: [C][rows][w], dst: [C][H][w])
__global__ void add_rows_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                i64 total, i64 src_plane, i64 dst_plane, i64 dst_off) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (i64)gridDim.x * blockDim.x) {
    const i64 c = i / src_plane, r = i % src_plane;
    dst[c * dst_plane + dst_off + r] += src[i];
  }
}

// buffer slots
enum { B_DATA = 0, B_DDATA = 1, B_SCRATCH = 2, B_W = 3, B_A = 6, B_B = 9, B_C = 12 };

static int copy_planes_h2d(float *dst, const float *src, i64 rows_w, i64 src_plane,
                           i64 planes, cudaStream_t st) {
  SBMC_CUDA_OK(cudaMemcpy2DAsync(dst, sizeof(float) * rows_w, src,
                                 sizeof(float) * src_plane, sizeof(float) * rows_w,
                                 (size_t)planes, cudaMemcpyHostToDevice, st));
  return SBMC_OK;
}
static int copy_planes_d2h(float *dst, const float *src, i64 rows_w, i64 dst_plane,
                           i64 planes, cudaStream_t st) {
  SBMC_CUDA_OK(cudaMemcpy2DAsync(dst, sizeof(float) * dst_plane, src,
                                 sizeof(float) * rows_w, sizeof(float) * rows_w,
                                 (size_t)planes, cudaMemcpyDeviceToHost, st));
  return SBMC_OK;
}

// The host entry points select `device` for their own work and give the caller
// its current device back on return.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// After a failure half-way through a pipeline: copies already queued still
// read / write the caller's host buffers, so wait for them before returning
// (keeps the error message of the call that failed).
static int drain_on_error(HostPipe &P, int rc) {
  if (rc != SBMC_OK) {
    if (P.s_h2d) cudaStreamSynchronize(P.s_h2d);
    if (P.s_comp) cudaStreamSynchronize(P.s_comp);
    if (P.s_d2h) cudaStreamSynchronize(P.s_d2h);
  }
  return rc;
}

static int sync_all(HostPipe &P) {
  SBMC_CUDA_OK(cudaStreamSynchronize(P.s_h2d));
  SBMC_CUDA_OK(cudaStreamSynchronize(P.s_comp));
  SBMC_CUDA_OK(cudaStreamSynchronize(P.s_d2h));
  return SBMC_OK;
}

static int fwd_host_locked(const float *data, const float *weights, float *output,
                    float *sum_w, i64 n, int c, i64 h, i64 w, int kh, int kw,
                    int device) {
  int rc = SBMC_OK;
  HostPipe &P = g_pipe;
  const i64 taps = (i64)kh * kw, plane = h * w;
  const i64 hb = band_rows(h, w, taps);
  float *g_data, *g_w[kSlots], *g_o[kSlots], *g_s[kSlots];
  if ((rc = ensure(B_DATA, sizeof(float) * c * plane, &g_data))) return rc;
  for (int s = 0; s < kSlots; ++s) {
    if ((rc = ensure(B_W + s, sizeof(float) * taps * hb * w, &g_w[s]))) return rc;
    if ((rc = ensure(B_A + s, sizeof(float) * c * hb * w, &g_o[s]))) return rc;
    if ((rc = ensure(B_B + s, sizeof(float) * hb * w, &g_s[s]))) return rc;
  }
  i64 job = 0;
  for (i64 img = 0; img < n; ++img) {
    if (img > 0)  // kernels of the previous image still read g_data
      SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_h2d, P.img_done, 0));
    SBMC_CUDA_OK(cudaMemcpyAsync(g_data, data + img * c * plane,
                                 sizeof(float) * c * plane, cudaMemcpyHostToDevice,
                                 P.s_h2d));
    for (i64 y0 = 0; y0 < h; y0 += hb, ++job) {
      const i64 rows = (y0 + hb <= h) ? hb : h - y0;
      const int s = (int)(job % kSlots);
      if (job >= kSlots) {
        SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_h2d, P.done[s], 0));   // W slot consumed
        SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_comp, P.down[s], 0));  // outputs drained
      }
      if ((rc = copy_planes_h2d(g_w[s], weights + img * taps * plane + y0 * w,
                                rows * w, plane, taps, P.s_h2d)))
        return rc;
      SBMC_CUDA_OK(cudaEventRecord(P.up[s], P.s_h2d));
      SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_comp, P.up[s], 0));
      rc = launch_fwd(g_data, g_w[s], g_o[s], g_s[s], 1, c, rows, w, kh, kw,
                      (int)y0, (int)(h - y0 - rows), P.s_comp);
      if (rc) return rc;
      SBMC_CUDA_OK(cudaEventRecord(P.done[s], P.s_comp));
      SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_d2h, P.done[s], 0));
      if ((rc = copy_planes_d2h(output + img * c * plane + y0 * w, g_o[s], rows * w,
                                plane, c, P.s_d2h)))
        return rc;
      SBMC_CUDA_OK(cudaMemcpyAsync(sum_w + img * plane + y0 * w, g_s[s],
                                   sizeof(float) * rows * w, cudaMemcpyDeviceToHost,
                                   P.s_d2h));
      SBMC_CUDA_OK(cudaEventRecord(P.down[s], P.s_d2h));
    }
    SBMC_CUDA_OK(cudaEventRecord(P.img_done, P.s_comp));
  }
  return sync_all(P);
}

static int bwd_host_locked(const float *data, const float *weights, const float *d_output,
                    const float *d_sum_w, float *d_data, float *d_weights, i64 n,
                    int c, i64 h, i64 w, int kh, int kw, int device) {
  int rc = SBMC_OK;
  HostPipe &P = g_pipe;
  const i64 taps = (i64)kh * kw, plane = h * w;
  const i64 hb = band_rows(h, w, taps);
  // rows a band's samples reach in d_data: q = p + dy - (kh-1-c0h)
  const i64 reach_t = kh - 1 - (kh - 1) / 2, reach_b = (kh - 1) / 2;
  float *g_data, *g_ddata, *g_scr, *g_w[kSlots], *g_dw[kSlots], *g_do[kSlots],
      *g_ds[kSlots];
  if ((rc = ensure(B_DATA, sizeof(float) * c * plane, &g_data))) return rc;
  if ((rc = ensure(B_DDATA, sizeof(float) * c * plane, &g_ddata))) return rc;
  if ((rc = ensure(B_SCRATCH, sizeof(float) * c * (hb + kh) * w, &g_scr))) return rc;
  for (int s = 0; s < kSlots; ++s) {
    if ((rc = ensure(B_W + s, sizeof(float) * taps * hb * w, &g_w[s]))) return rc;
    if ((rc = ensure(B_C + s, sizeof(float) * taps * hb * w, &g_dw[s]))) return rc;
    if ((rc = ensure(B_A + s, sizeof(float) * c * hb * w, &g_do[s]))) return rc;
    if ((rc = ensure(B_B + s, sizeof(float) * hb * w, &g_ds[s]))) return rc;
  }
  i64 job = 0;
  for (i64 img = 0; img < n; ++img) {
    if (img > 0) SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_h2d, P.img_done, 0));
    SBMC_CUDA_OK(cudaMemcpyAsync(g_data, data + img * c * plane,
                                 sizeof(float) * c * plane, cudaMemcpyHostToDevice,
                                 P.s_h2d));
    // (s_comp runs in order: the previous image's d_data download is queued on
    // s_d2h behind img_done, and the memset below waits for it via `down`.)
    if (img > 0) SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_comp, P.down[(job - 1) % kSlots], 0));
    SBMC_CUDA_OK(cudaMemsetAsync(g_ddata, 0, sizeof(float) * c * plane, P.s_comp));
    for (i64 y0 = 0; y0 < h; y0 += hb, ++job) {
      const i64 rows = (y0 + hb <= h) ? hb : h - y0;
      const int s = (int)(job % kSlots);
      if (job >= kSlots) {
        SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_h2d, P.done[s], 0));
        SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_comp, P.down[s], 0));
      }
      if ((rc = copy_planes_h2d(g_w[s], weights + img * taps * plane + y0 * w,
                                rows * w, plane, taps, P.s_h2d)))
        return rc;
      if ((rc = copy_planes_h2d(g_do[s], d_output + img * c * plane + y0 * w,
                                rows * w, plane, c, P.s_h2d)))
        return rc;
      SBMC_CUDA_OK(cudaMemcpyAsync(g_ds[s], d_sum_w + img * plane + y0 * w,
                                   sizeof(float) * rows * w, cudaMemcpyHostToDevice,
                                   P.s_h2d));
      SBMC_CUDA_OK(cudaEventRecord(P.up[s], P.s_h2d));
      SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_comp, P.up[s], 0));
      rc = launch_bwd_dweights(g_data, g_do[s], g_ds[s], g_dw[s], 1, c, rows, w, kh,
                               kw, (int)y0, (int)(h - y0 - rows), P.s_comp);
      if (rc) return rc;
      const i64 top = y0 < reach_t ? y0 : reach_t;
      const i64 below = h - y0 - rows;
      const i64 bot = below < reach_b ? below : reach_b;
      rc = launch_bwd_ddata(g_w[s], g_do[s], g_scr, 1, c, rows, w, kh, kw, (int)top,
                            (int)bot, P.s_comp);
      if (rc) return rc;
      {
        const i64 src_plane = (top + rows + bot) * w, total = src_plane * c;
        i64 blocks = ceil_div(total, 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        KernelTimer timer(SBMC_KERNEL_OTHER, P.s_comp);
        add_rows_kernel<<<(unsigned)blocks, 256, 0, P.s_comp>>>(
            g_scr, g_ddata, total, src_plane, plane, (y0 - top) * w);
        count_launch();
        SBMC_CUDA_OK(cudaGetLastError());
      }
      SBMC_CUDA_OK(cudaEventRecord(P.done[s], P.s_comp));
      SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_d2h, P.done[s], 0));
      if ((rc = copy_planes_d2h(d_weights + img * taps * plane + y0 * w, g_dw[s],
                                rows * w, plane, taps, P.s_d2h)))
        return rc;
      if (y0 + hb >= h)  // last band of the image: d_data is complete
        SBMC_CUDA_OK(cudaMemcpyAsync(d_data + img * c * plane, g_ddata,
                                     sizeof(float) * c * plane,
                                     cudaMemcpyDeviceToHost, P.s_d2h));
      SBMC_CUDA_OK(cudaEventRecord(P.down[s], P.s_d2h));
    }
    SBMC_CUDA_OK(cudaEventRecord(P.img_done, P.s_comp));
  }
  return sync_all(P);
}

static int s2g_host_locked(const float *scatter, float *gather, i64 n, int kh, int kw,
                    i64 h, i64 w, int device) {
  // The transpose mixes rows of different taps, so bands do not help: stream one
  // image (all taps) at a time through two device buffers.
  int rc = SBMC_OK;
  HostPipe &P = g_pipe;
  const i64 img_elems = (i64)kh * kw * h * w;
  float *g_in[2], *g_out[2];
  for (int s = 0; s < 2; ++s) {
    if ((rc = ensure(B_W + s, sizeof(float) * img_elems, &g_in[s]))) return rc;
    if ((rc = ensure(B_C + s, sizeof(float) * img_elems, &g_out[s]))) return rc;
  }
  for (i64 img = 0; img < n; ++img) {
    const int s = (int)(img & 1);
    if (img >= 2) {
      SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_h2d, P.done[s], 0));
      SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_comp, P.down[s], 0));
    }
    SBMC_CUDA_OK(cudaMemcpyAsync(g_in[s], scatter + img * img_elems,
                                 sizeof(float) * img_elems, cudaMemcpyHostToDevice,
                                 P.s_h2d));
    SBMC_CUDA_OK(cudaEventRecord(P.up[s], P.s_h2d));
    SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_comp, P.up[s], 0));
    if ((rc = launch_s2g(g_in[s], g_out[s], 1, kh, kw, h, w, P.s_comp))) return rc;
    SBMC_CUDA_OK(cudaEventRecord(P.done[s], P.s_comp));
    SBMC_CUDA_OK(cudaStreamWaitEvent(P.s_d2h, P.done[s], 0));
    SBMC_CUDA_OK(cudaMemcpyAsync(gather + img * img_elems, g_out[s],
                                 sizeof(float) * img_elems, cudaMemcpyDeviceToHost,
                                 P.s_d2h));
    SBMC_CUDA_OK(cudaEventRecord(P.down[s], P.s_d2h));
  }
  return sync_all(P);
}

// Public bodies: serialise on the pipe, select the device, drain on failure.
template <typename Fn>
static int with_pipe(int device, Fn body) {
  std::lock_guard<std::mutex> lock(g_pipe_mu);
  DeviceGuard guard;
  int rc = pipe_init_locked(device);
  if (rc) return rc;
  return drain_on_error(g_pipe, body());
}

static int fwd_host(const float *data, const float *weights, float *output, float *sum_w, i64 n,
                    int c, i64 h, i64 w, int kh, int kw, int device) {
  return with_pipe(device, [&] {
    return fwd_host_locked(data, weights, output, sum_w, n, c, h, w, kh, kw, device);
  });
}

static int bwd_host(const float *data, const float *weights, const float *d_output,
                    const float *d_sum_w, float *d_data, float *d_weights, i64 n, int c, i64 h,
                    i64 w, int kh, int kw, int device) {
  return with_pipe(device, [&] {
    return bwd_host_locked(data, weights, d_output, d_sum_w, d_data, d_weights, n, c, h, w, kh,
                           kw, device);
  });
}

static int s2g_host(const float *scatter, float *gather, i64 n, int kh, int kw, i64 h, i64 w,
                    int device) {
  return with_pipe(device, [&] { return s2g_host_locked(scatter, gather, n, kh, kw, h, w, device); });
}

static int check_host_args(i64 n, int c, i64 h, i64 w, int kh, int kw,
                           const void *const *ptrs, int count) {
  if (n < 0 || h < 0 || w < 0 || c < 1 || kh < 1 || kw < 1) {
    set_error("invalid shape n=%lld c=%d h=%lld w=%lld kh=%d kw=%d", (long long)n,
              c, (long long)h, (long long)w, kh, kw);
    return SBMC_EINVAL;
  }
  if (n == 0 || h == 0 || w == 0) return 1;  // nothing to do
  for (int i = 0; i < count; ++i)
    if (!ptrs[i]) {
      set_error("null pointer argument (#%d)", i);
      return SBMC_EINVAL;
    }
  return SBMC_OK;
}

}  // namespace sbmc

extern "C" {

int sbmc_b200_host_release(void) {
  std::lock_guard<std::mutex> lock(sbmc::g_pipe_mu);
  sbmc::DeviceGuard guard;
  return sbmc::pipe_release_locked();
}

int sbmc_scatter2gather_host_f32(const float *scatter, float *gather, int64_t n,
                                 int kh, int kw, int64_t h, int64_t w, int device) {
  const void *ptrs[] = {scatter, gather};
  int rc = sbmc::check_host_args(n, 1, h, w, kh, kw, ptrs, 2);
  if (rc) return rc > 0 ? SBMC_OK : rc;
  return sbmc::s2g_host(scatter, gather, n, kh, kw, h, w, device);
}

int sbmc_kernel_weighting_fwd_host_f32(const float *data, const float *weights,
                                       float *output, float *sum_w, int64_t n,
                                       int c, int64_t h, int64_t w, int kh, int kw,
                                       int device) {
  const void *ptrs[] = {data, weights, output, sum_w};
  int rc = sbmc::check_host_args(n, c, h, w, kh, kw, ptrs, 4);
  if (rc) return rc > 0 ? SBMC_OK : rc;
  return sbmc::fwd_host(data, weights, output, sum_w, n, c, h, w, kh, kw, device);
}

int sbmc_kernel_weighting_bwd_host_f32(const float *data, const float *weights,
                                       const float *sum_w, const float *d_output,
                                       const float *d_sum_w, float *d_data,
                                       float *d_weights, int64_t n, int c,
                                       int64_t h, int64_t w, int kh, int kw,
                                       int device) {
  (void)sum_w;
  const void *ptrs[] = {data, weights, d_output, d_sum_w, d_data, d_weights};
  int rc = sbmc::check_host_args(n, c, h, w, kh, kw, ptrs, 6);
  if (rc) return rc > 0 ? SBMC_OK : rc;
  return sbmc::bwd_host(data, weights, d_output, d_sum_w, d_data, d_weights, n, c, h,
                        w, kh, kw, device);
}

}  // extern "C"
