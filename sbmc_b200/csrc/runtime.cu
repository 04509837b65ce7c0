// runtime.cu -- error reporting, counters and the TMA tensor-map encoder.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "umma.cuh"

namespace sbmc {

static thread_local char g_err[512] = "";
static thread_local int g_path = 0;
static std::atomic<int> g_force_generic{0};
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void note_path(int path) { g_path = path; }

// One line on stderr the first time an operator leaves its tuned sm_100a kernel for the
// one-thread-per-output generic kernel (>= 10x slower): the cliff must not be silent.
// SBMC_B200_QUIET=1 suppresses it; forced generic runs (tests) do not warn.
void warn_generic(const char *op, int c, int kh, int kw, long long w) {
  static std::atomic<int> warned{0};
  if (force_generic() || warned.exchange(1)) return;
  const char *q = getenv("SBMC_B200_QUIET");
  if (q && q[0] == '1') return;
  fprintf(stderr,
          "[sbmc_b200] %s: no tuned kernel for C=%d K=%dx%d W=%lld (tuned: C=3 with K in "
          "{3,5,7,9,...,21}, C=5 with K in {3,5}; W %% 4 == 0, 16-byte aligned pointers): "
          "using the generic kernel, expect >= 10x lower throughput\n",
          op, c, kh, kw, w);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool force_generic() { return g_force_generic.load() != 0; }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

// ---- optional per-kernel device timing (bench.py's roofline leg) -----------
// When enabled, every launcher brackets its kernel with two CUDA events on the
// launching stream; sbmc_b200_timing_collect() sums the elapsed times per kind.
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
struct TimedSpan {
  int kind;
  cudaEvent_t a, b;
};
static std::vector<TimedSpan> g_spans;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t pool_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

KernelTimer::KernelTimer(int kind, cudaStream_t st) : kind_(kind), st_(st), a_(nullptr) {
  if (!g_timing.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_timing_mu);
  a_ = pool_event();
  if (a_) cudaEventRecord(a_, st_);
}
KernelTimer::~KernelTimer() {
  if (!a_) return;
  std::lock_guard<std::mutex> lock(g_timing_mu);
  cudaEvent_t b = pool_event();
  if (b) cudaEventRecord(b, st_);
  g_spans.push_back({kind_, a_, b});
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t,
                                  void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// dims / box are innermost-first; strides_bytes has rank-1 entries (dims 1..).
bool encode_tensor_map_f32(CUtensorMap *map, const void *base, int rank,
                           const uint64_t *dims, const uint64_t *strides_bytes,
                           const uint32_t *box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (dims[i] == 0 || dims[i] > (1ull << 32) || box[i] == 0 || box[i] > 256) {
      set_error("tensor map: dim %d out of range (%llu, box %u)", i,
                (unsigned long long)dims[i], box[i]);
      return false;
    }
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if ((strides_bytes[i] & 15) || strides_bytes[i] >= (1ull << 40)) {
      set_error("tensor map: stride %d = %llu not encodable", i,
                (unsigned long long)strides_bytes[i]);
      return false;
    }
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                  const_cast<void *>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);  // OOB reads give 0.0f
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return false;
  }
  return true;
}

// bf16 [rows][inner] row-major, box {box_inner, box_rows}, 128-byte swizzle (the
// canonical K-major UMMA operand layout: box_inner * 2 bytes must be 128).
bool encode_tensor_map_bf16_2d_sw128(CUtensorMap *map, const void *base, uint64_t inner,
                                     uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  if (box_inner * 2 != 128 || box_rows == 0 || box_rows > 256 || (inner * 2) % 16 != 0) {
    set_error("tensor map (bf16 sw128): unsupported box %u x %u", box_inner, box_rows);
    return false;
  }
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstr[1] = {inner * 2};
  cuuint32_t bdim[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 sw128) failed with CUresult %d", (int)r);
    return false;
  }
  return true;
}

// bf16 activations [n][rows][inner] (channels innermost, `img_stride` elements
// between images), box {64, box_rows, 1}, 128-byte swizzle: one box is one K-major
// UMMA operand slab of box_rows pixels x 64 channels.
bool encode_tensor_map_bf16_3d_sw128(CUtensorMap *map, const void *base, uint64_t inner,
                                     uint64_t rows, uint64_t n, uint64_t img_stride,
                                     uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  if (inner < 64 || (inner * 2) % 16 != 0 || (img_stride * 2) % 16 != 0 || box_rows == 0 ||
      box_rows > 256 || (reinterpret_cast<uintptr_t>(base) & 15)) {
    set_error("tensor map (bf16 3d sw128): unsupported layout");
    return false;
  }
  cuuint64_t gdim[3] = {inner, rows, n};
  cuuint64_t gstr[2] = {inner * 2, img_stride * 2};
  cuuint32_t bdim[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 3d sw128) failed with CUresult %d", (int)r);
    return false;
  }
  return true;
}

// bf16 tensor of any rank <= 5 with channels innermost, box {64, ...}, 128-byte
// swizzle: every (64 channels x box[1] rows) slice of a box is one K-major UMMA
// operand slab.  dims / box are innermost first; strides_bytes[i] is the stride of
// dimension i + 1.
bool encode_tensor_map_bf16_sw128(CUtensorMap *map, const void *base, int rank,
                                  const uint64_t *dims, const uint64_t *strides_bytes,
                                  const uint32_t *box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  if (rank < 2 || rank > 5 || box[0] != 64 || (reinterpret_cast<uintptr_t>(base) & 15)) {
    set_error("tensor map (bf16 sw128): unsupported rank %d / box %u / base alignment", rank,
              box[0]);
    return false;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (dims[i] == 0 || dims[i] > (1ull << 32) || box[i] == 0 || box[i] > 256) {
      set_error("tensor map (bf16 sw128): dim %d out of range (%llu, box %u)", i,
                (unsigned long long)dims[i], box[i]);
      return false;
    }
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if ((strides_bytes[i] & 15) || strides_bytes[i] >= (1ull << 40)) {
      set_error("tensor map (bf16 sw128): stride %d = %llu not encodable", i,
                (unsigned long long)strides_bytes[i]);
      return false;
    }
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                  const_cast<void *>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 sw128, rank %d) failed with CUresult %d", rank,
              (int)r);
    return false;
  }
  return true;
}

}  // namespace sbmc

extern "C" {

int sbmc_b200_version(void) { return 100; /* 0.1.0 */ }
const char *sbmc_b200_last_error(void) { return sbmc::g_err; }
int sbmc_b200_force_generic(int flag) {
  return sbmc::g_force_generic.exchange(flag ? 1 : 0);
}
int sbmc_b200_last_path(void) { return sbmc::g_path; }
int64_t sbmc_b200_launch_count(void) { return sbmc::g_launches.load(); }

int sbmc_b200_timing_enable(int flag) {
  return sbmc::g_timing.exchange(flag ? 1 : 0);
}

int sbmc_b200_timing_collect(double *ms_by_kind, int64_t *launches_by_kind) {
  using namespace sbmc;
  std::lock_guard<std::mutex> lock(g_timing_mu);
  for (int k = 0; k < SBMC_NUM_KERNEL_KINDS; ++k) {
    if (ms_by_kind) ms_by_kind[k] = 0.0;
    if (launches_by_kind) launches_by_kind[k] = 0;
  }
  int rc = SBMC_OK;
  for (const TimedSpan &s : g_spans) {
    float ms = 0.f;
    if (s.b && cudaEventSynchronize(s.b) == cudaSuccess &&
        cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      if (s.kind >= 0 && s.kind < SBMC_NUM_KERNEL_KINDS) {
        if (ms_by_kind) ms_by_kind[s.kind] += ms;
        if (launches_by_kind) launches_by_kind[s.kind] += 1;
      }
    } else {
      set_error("timing: event query failed");
      rc = SBMC_ECUDA;
    }
    if (s.a) g_event_pool.push_back(s.a);
    if (s.b) g_event_pool.push_back(s.b);
  }
  g_spans.clear();
  return rc;
}

}  // extern "C"
