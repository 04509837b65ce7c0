// optim_emul.cpp -- the fused optimizer's device arithmetic on the CPU
// (sbmc_b200/csrc/optim_body.cuh: same table decoding, same chunking, same
// per-element update as optim.cu), for tests/test_optim.py without a GPU.
#include <cmath>

#include "../../sbmc_b200/csrc/optim_body.cuh"

extern "C" {

// multi_sqnorm_kernel + finalize_norm_kernel: partial sums per chunk, then the
// norm and the clip coefficient.
int emul_grad_norm(const long long *tensors, const long long *chunks, long long nchunks,
                   float *partial, float max_norm, float *norm_and_coef) {
  for (long long c = 0; c < nchunks; ++c) {
    const sbmc::MtTensor t = sbmc::mt_tensor(tensors, chunks[2 * c]);
    const long long start = chunks[2 * c + 1];
    const long long stop = (start + SBMC_MT_CHUNK_ELEMS < t.n) ? start + SBMC_MT_CHUNK_ELEMS : t.n;
    float lanes[256] = {0.f};
    for (int tid = 0; tid < 256; ++tid)
      for (long long i = start + tid; i < stop; i += 256) lanes[tid] = fmaf(t.g[i], t.g[i], lanes[tid]);
    float s = 0.f;
    for (int tid = 0; tid < 256; ++tid) s += lanes[tid];
    partial[c] = s;
  }
  double acc = 0.0;
  for (long long c = 0; c < nchunks; ++c) acc += (double)partial[c];
  const float norm = (float)std::sqrt(acc);
  const float coef = max_norm / (norm + 1e-6f);
  norm_and_coef[0] = norm;
  norm_and_coef[1] = coef < 1.0f ? coef : 1.0f;
  return 0;
}

int emul_adam(const long long *tensors, const long long *chunks, long long nchunks,
              const float *clip_coef, double lr, double beta1, double beta2, double eps,
              double bias_correction1, double bias_correction2_sqrt) {
  const sbmc::AdamScalars s =
      sbmc::adam_scalars(lr, beta1, beta2, eps, bias_correction1, bias_correction2_sqrt);
  const float coef = clip_coef ? *clip_coef : 1.0f;
  for (long long c = 0; c < nchunks; ++c) {
    const sbmc::MtTensor t = sbmc::mt_tensor(tensors, chunks[2 * c]);
    const long long start = chunks[2 * c + 1];
    const long long stop = (start + SBMC_MT_CHUNK_ELEMS < t.n) ? start + SBMC_MT_CHUNK_ELEMS : t.n;
    for (long long i = start; i < stop; ++i)
      sbmc::adam_element(t.p + i, t.g + i, t.m + i, t.v + i, coef, s);
  }
  return 0;
}

}  // extern "C"
