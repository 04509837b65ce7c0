// sweep.cu -- tuning sweep for the KernelWeighting kernels (developer tool, not
// part of the library).  Times every compiled (ROWS/NSEG, MINB, CH) variant of
// the forward, d_weights and d_data kernels on the BASELINE config-2 call shape
// (N=4, C=3, 720x1280, K=21) with CUDA events and prints achieved GB/s against
// the algorithmic bytes.  Build: see tools/build_sweep.sh.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../sbmc_b200/csrc/kw_launch.cuh"
#include "../sbmc_b200/csrc/s2g.cu"
#include "../sbmc_b200/csrc/splat.cu"
#include "../sbmc_b200/csrc/splat_bwd.cu"

using namespace sbmc;

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e = (x);                                                      \
    if (e != cudaSuccess) {                                                   \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                 \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

__global__ void fill_kernel(float *p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = (float)(h & 0xffff) / 65536.0f - 0.5f;
  }
}

// reference point: a pure read stream (sum of the volume) with the same load flavour
__global__ void __launch_bounds__(256) read_stream_kernel(const float *p, size_t n4, float *out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(p + 4 * i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) *out = acc;   // never true: keeps the loads alive
}

// Synthetic write-pattern probes (no arithmetic): which store pattern over the
// K*K-plane volume reaches the memset rate?  CTA = ROWS warps; warp r owns row
// Y0 + r and, per tap, writes SEGS consecutive 512-byte segments of that row.
template <int ROWS, int SEGS>
__global__ void __launch_bounds__(ROWS * 32)
write_pattern_kernel(float *dW, int H, int W, int taps, int xtiles, int ytiles) {
  const int xt = blockIdx.x % xtiles;
  const unsigned r = blockIdx.x / xtiles;
  const int yt = r % ytiles, n = r / ytiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = yt * ROWS + warp;
  const int x0 = xt * SEGS * 128 + 4 * lane;
  if (y >= H) return;
  const long long plane = (long long)H * W;
  float *wp = dW + (long long)n * taps * plane + (long long)y * W + x0;
  const float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
  for (int t = 0; t < taps; ++t) {
#pragma unroll
    for (int s = 0; s < SEGS; ++s)
      if (x0 + s * 128 < W) stg_stream(wp + (long long)t * plane + s * 128, v);
  }
}

// The same volume written with TMA tensor stores (cp.async.bulk.tensor ... global.shared,
// SASS UTMASTG): per tap plane one [ROWS x 128 px] box from shared memory, DEPTH bulk
// groups in flight.  Answers "would smem-staged TMA stores lift the d_weights write rate?"
template <int ROWS, int DEPTH>
__global__ void __launch_bounds__(128)
tma_store_pattern_kernel(const __grid_constant__ CUtensorMap map, int taps, int xtiles,
                         int ytiles) {
  extern __shared__ __align__(128) float tile[];          // [2][ROWS][128]
  const int xt = blockIdx.x % xtiles;
  const unsigned r = blockIdx.x / xtiles;
  const int yt = r % ytiles, n = r / ytiles;
  for (int i = threadIdx.x; i < 2 * ROWS * 128; i += 128) tile[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 0; t < taps; ++t) {
      tma_store_4d(&map, tile + (t & 1) * ROWS * 128, xt * 128, yt * ROWS, t, n);
      tma_commit_group();
      tma_wait_group_read<DEPTH>();
    }
    tma_wait_group<0>();
  }
}

struct Ctx {
  i64 n = 4, h = 720, w = 1280;
  int k = 21;
  float *data, *wt, *out, *sw, *dout, *dsw, *ddata, *dwt, *planes;
  cudaStream_t st;
  int iters = 5;
};

template <typename F>
static void time_it(Ctx &c, const char *name, double bytes_per_sample, F &&f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 2; ++i) {
    int rc = f();
    if (rc) {
      printf("%-44s FAILED rc=%d (%s)\n", name, rc, sbmc_b200_last_error());
      return;
    }
  }
  CK(cudaStreamSynchronize(c.st));
  CK(cudaEventRecord(a, c.st));
  for (int i = 0; i < c.iters; ++i) f();
  CK(cudaEventRecord(b, c.st));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  ms /= c.iters;
  const double samples = (double)c.n * c.h * c.w;
  printf("%-44s %8.3f ms  %8.1f GB/s  %8.1f Msamples/s\n", name, ms,
         bytes_per_sample * samples / (ms * 1e-3) / 1e9, samples / (ms * 1e-3) / 1e6);
  fflush(stdout);
}

#define FWD(ROWS, MINB, CH)                                                   \
  time_it(c, "fwd  rows=" #ROWS " minb=" #MINB " ch=" #CH, 4.0 * (441 + 7), [&] { \
    return run_fwd<3, 21, ROWS, MINB, CH>(c.data, c.wt, c.out, c.sw, c.n, c.h, c.w, \
                                          c.k, 0, 0, c.st);                  \
  });
#define DWT(ROWS, MINB, CH)                                                   \
  time_it(c, "dwt  rows=" #ROWS " minb=" #MINB " ch=" #CH, 4.0 * (441 + 7), [&] { \
    return run_bwd_dweights<3, 21, ROWS, MINB, CH>(c.data, c.dout, c.dsw, c.dwt, c.n, \
                                                   c.h, c.w, c.k, 0, 0, c.st); \
  });
#define DDA(NSEG, MINB, CH)                                                   \
  time_it(c, "dda  nseg=" #NSEG " minb=" #MINB " ch=" #CH, 4.0 * (441 + 6), [&] { \
    return run_bwd_ddata<3, 21, NSEG, MINB, CH>(c.wt, c.dout, c.ddata, c.n, c.h, c.w, \
                                                c.k, 0, 0, c.st);             \
  });

#define DWS(ROWS, MINB, CH, ST)                                               \
  time_it(c, "dwt  rows=" #ROWS " minb=" #MINB " ch=" #CH " store=" #ST, 4.0 * (441 + 7), [&] { \
    return run_bwd_dweights<3, 21, ROWS, MINB, CH, ST>(c.data, c.dout, c.dsw, c.dwt, c.n, \
                                                       c.h, c.w, c.k, 0, 0, c.st); \
  });
#define S2G(ROWS, STAGES, TPS)                                                \
  time_it(c, "s2g  rows=" #ROWS " stages=" #STAGES " tps=" #TPS, 8.0 * 441, [&] { \
    return run_s2g<ROWS, STAGES, TPS>(c.wt, c.dwt, c.n, c.k, c.k, c.h, c.w, c.st); \
  });
#define SPL(ROWS, STAGES, TPS)                                                \
  time_it(c, "splat rows=" #ROWS " stages=" #STAGES " tps=" #TPS, 4.0 * (441 + 3 + 10), [&] { \
    return run_splat<3, 21, ROWS, STAGES, TPS>(c.wt, c.data, c.out, c.sw, c.dsw, c.n, c.h, \
                                               c.w, c.k, 1, 0, c.st);         \
  });

#define SPB(ROWS, MINB, CH)                                                   \
  time_it(c, "splat_bwd rows=" #ROWS " minb=" #MINB " ch=" #CH, 4.0 * (2 * 441 + 12), [&] { \
    return run_splat_bwd<3, 21, ROWS, MINB, CH>(c.planes, c.wt, c.data, c.dwt, c.ddata, c.n, \
                                                c.h, c.w, c.k, c.st);         \
  });

int main(int argc, char **argv) {
  Ctx c;
  const char *which = argc > 1 ? argv[1] : "all";
  if (argc > 2) c.n = atoi(argv[2]);
  CK(cudaStreamCreate(&c.st));
  const size_t img = (size_t)c.n * 3 * c.h * c.w, pl = (size_t)c.n * c.h * c.w;
  const size_t vol = (size_t)c.n * c.k * c.k * c.h * c.w;
  CK(cudaMalloc(&c.data, img * 4));  CK(cudaMalloc(&c.out, img * 4));
  CK(cudaMalloc(&c.dout, img * 4));  CK(cudaMalloc(&c.ddata, img * 4));
  CK(cudaMalloc(&c.sw, pl * 4));     CK(cudaMalloc(&c.dsw, pl * 4));
  CK(cudaMalloc(&c.wt, vol * 4));    CK(cudaMalloc(&c.dwt, vol * 4));
  CK(cudaMalloc(&c.planes, pl * 6 * 4));
  fill_kernel<<<1184, 256>>>(c.planes, pl * 6, 5);
  fill_kernel<<<1184, 256>>>(c.data, img, 1);
  fill_kernel<<<1184, 256>>>(c.dout, img, 2);
  fill_kernel<<<1184, 256>>>(c.dsw, pl, 3);
  fill_kernel<<<1184, 256>>>(c.wt, vol, 4);
  CK(cudaDeviceSynchronize());
  // reference points: device-to-device copy and memset of the weight volume
  time_it(c, "memcpy d2d (R+W of the volume)", 4.0 * 441 * 2, [&] {
    return (int)cudaMemcpyAsync(c.dwt, c.wt, vol * 4, cudaMemcpyDeviceToDevice, c.st);
  });
  time_it(c, "memset (W of the volume)", 4.0 * 441, [&] {
    return (int)cudaMemsetAsync(c.dwt, 0, vol * 4, c.st);
  });
  time_it(c, "read stream (R of the volume)", 4.0 * 441, [&] {
    read_stream_kernel<<<148 * 16, 256, 0, c.st>>>(c.wt, vol / 4, c.out);
    return 0;
  });
#define WPAT(ROWS, SEGS)                                                      \
  time_it(c, "write pattern rows=" #ROWS " segs=" #SEGS, 4.0 * 441, [&] {      \
    const int xt = (int)ceil_div(c.w, SEGS * 128), yt = (int)ceil_div(c.h, ROWS); \
    write_pattern_kernel<ROWS, SEGS><<<(unsigned)(xt * yt * c.n), ROWS * 32, 0, c.st>>>( \
        c.dwt, (int)c.h, (int)c.w, c.k * c.k, xt, yt);                        \
    return 0;                                                                 \
  });
#define TPAT(ROWS, DEPTH)                                                     \
  time_it(c, "TMA store pattern rows=" #ROWS " depth=" #DEPTH, 4.0 * 441, [&] { \
    CUtensorMap m;                                                            \
    const uint64_t dims[4] = {(uint64_t)c.w, (uint64_t)c.h, (uint64_t)(c.k * c.k), (uint64_t)c.n}; \
    const uint64_t str[3] = {(uint64_t)c.w * 4, (uint64_t)c.w * c.h * 4,       \
                             (uint64_t)c.w * c.h * c.k * c.k * 4};             \
    const uint32_t box[4] = {128, ROWS, 1, 1};                                \
    if (!encode_tensor_map_f32(&m, c.dwt, 4, dims, str, box)) return 1;       \
    const int xt = (int)ceil_div(c.w, 128), yt = (int)ceil_div(c.h, ROWS);    \
    tma_store_pattern_kernel<ROWS, DEPTH><<<(unsigned)(xt * yt * c.n), 128,   \
                                            2 * ROWS * 128 * 4, c.st>>>(m, c.k * c.k, xt, yt); \
    return 0;                                                                 \
  });
  if (!strcmp(which, "tpat")) {
    WPAT(16, 1)
    TPAT(8, 4) TPAT(16, 4) TPAT(16, 8) TPAT(16, 16) TPAT(32, 8) TPAT(32, 16) TPAT(48, 8)
  }
  if (!strcmp(which, "wpat")) {
    WPAT(16, 1) WPAT(8, 1) WPAT(8, 2) WPAT(4, 4) WPAT(8, 5) WPAT(2, 10) WPAT(1, 10) WPAT(4, 10)
    WPAT(16, 2) WPAT(4, 2)
  }
  const bool all = !strcmp(which, "all");
  if (all || !strcmp(which, "fwd")) {
    FWD(8, 2, 7) FWD(8, 2, 11) FWD(8, 2, 21) FWD(8, 3, 3) FWD(8, 3, 7) FWD(8, 1, 21)
    FWD(4, 4, 7) FWD(4, 6, 3) FWD(16, 1, 7) FWD(16, 1, 11) FWD(4, 2, 21)
  }
  if (all || !strcmp(which, "dwt")) {
    DWT(8, 2, 7) DWT(8, 3, 7) DWT(8, 4, 3) DWT(4, 4, 7) DWT(16, 1, 7) DWT(8, 2, 21)
  }
  if (all || !strcmp(which, "dda")) {
    DDA(10, 1, 7) DDA(10, 1, 11) DDA(10, 1, 21) DDA(5, 2, 7) DDA(5, 1, 21) DDA(5, 1, 11)
    DDA(2, 4, 7) DDA(2, 2, 21) DDA(1, 8, 7) DDA(4, 2, 7) DDA(2, 6, 7) DDA(2, 4, 11)
  }
  if (all || !strcmp(which, "dws")) {
    DWS(16, 1, 7, 2) DWS(16, 1, 7, 3) DWS(8, 2, 7, 3) DWS(8, 2, 7, 2) DWS(16, 1, 7, 0)
  }
  if (all || !strcmp(which, "s2g")) {
    S2G(8, 4, 3) S2G(4, 6, 3) S2G(8, 3, 7) S2G(4, 8, 3) S2G(16, 2, 3) S2G(2, 8, 7)
    S2G(8, 6, 3)
  }
  if (all || !strcmp(which, "splat")) {
    SPL(8, 2, 7) SPL(8, 3, 7) SPL(4, 4, 7) SPL(4, 3, 7) SPL(16, 2, 7) SPL(8, 4, 3) SPL(8, 2, 11)
    SPL(4, 6, 7)
  }
  if (all || !strcmp(which, "spb")) {
    SPB(8, 2, 3) SPB(8, 2, 7) SPB(8, 1, 7) SPB(4, 4, 3) SPB(16, 1, 3) SPB(8, 3, 3) SPB(4, 3, 7)
  }
  return 0;
}
