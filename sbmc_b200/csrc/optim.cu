// optim.cu -- fused gradient clipping + Adam over all parameters of the model:
// the optimizer half of the reference's training step (sbmc/interfaces.py:78-106).
//
//   1. multi_sqnorm_kernel : one CTA per 64 K-element chunk of one tensor writes the
//      sum of squares of its gradient chunk (no atomics: deterministic);
//   2. finalize_norm_kernel: one CTA adds the partial sums (double), writes the
//      total norm and the clip coefficient min(1, max_norm / (norm + 1e-6)) --
//      the host never waits for it;
//   3. multi_adam_kernel   : scales the gradients by the coefficient and applies
//      Adam to every parameter; pure HBM stream, 4 reads + 3 (4) writes per element.
// Tensors are addressed through device tables built by the caller
// (sbmc_b200/optim.py): tensors int64 [nt][5] = {param, grad, exp_avg, exp_avg_sq,
// numel}, chunks int64 [nc][2] = {tensor index, first element}.
#include "common.cuh"
#include "optim_body.cuh"

namespace sbmc {

__global__ void __launch_bounds__(256) multi_sqnorm_kernel(const long long *__restrict__ tensors,
                                                           const long long *__restrict__ chunks,
                                                           float *__restrict__ partial) {
  const long long c = blockIdx.x;
  const MtTensor t = mt_tensor(tensors, chunks[2 * c]);
  const long long start = chunks[2 * c + 1];
  const long long stop = (start + SBMC_MT_CHUNK_ELEMS < t.n) ? start + SBMC_MT_CHUNK_ELEMS : t.n;
  float acc = 0.f;
  for (long long i = start + threadIdx.x; i < stop; i += 256) {
    const float g = t.g[i];
    acc = fmaf(g, g, acc);
  }
  __shared__ float warp_sum[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += warp_sum[w];
    partial[c] = s;
  }
}

__global__ void __launch_bounds__(256) finalize_norm_kernel(const float *__restrict__ partial,
                                                            long long nchunks, float max_norm,
                                                            float *__restrict__ norm_and_coef) {
  double acc = 0.0;
  for (long long i = threadIdx.x; i < nchunks; i += 256) acc += (double)partial[i];
  __shared__ double warp_sum[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += warp_sum[w];
    const float norm = (float)sqrt(s);
    const float coef = max_norm / (norm + 1e-6f);      // torch.nn.utils.clip_grad_norm_
    norm_and_coef[0] = norm;
    norm_and_coef[1] = coef < 1.0f ? coef : 1.0f;
  }
}

__global__ void __launch_bounds__(256) multi_adam_kernel(const long long *__restrict__ tensors,
                                                         const long long *__restrict__ chunks,
                                                         const float *__restrict__ coef_ptr,
                                                         const AdamScalars s) {
  const long long c = blockIdx.x;
  const MtTensor t = mt_tensor(tensors, chunks[2 * c]);
  const long long start = chunks[2 * c + 1];
  const long long stop = (start + SBMC_MT_CHUNK_ELEMS < t.n) ? start + SBMC_MT_CHUNK_ELEMS : t.n;
  const float coef = coef_ptr ? *coef_ptr : 1.0f;
  for (long long i = start + threadIdx.x; i < stop; i += 256)
    adam_element(t.p + i, t.g + i, t.m + i, t.v + i, coef, s);
}

// Capturable variant (CUDA graphs): the step count lives on the device, so that a replayed
// graph advances the bias corrections; they are derived in double per thread, as the host
// does for the plain entry point.  *step = steps taken so far.
__global__ void __launch_bounds__(256) multi_adam_devstep_kernel(
    const long long *__restrict__ tensors, const long long *__restrict__ chunks,
    const float *__restrict__ coef_ptr, double lr, double beta1, double beta2, double eps,
    const float *__restrict__ step) {
  const double t = (double)(*step) + 1.0;
  const AdamScalars s = adam_scalars(lr, beta1, beta2, eps, 1.0 - pow(beta1, t),
                                     sqrt(1.0 - pow(beta2, t)));
  const long long c = blockIdx.x;
  const MtTensor tn = mt_tensor(tensors, chunks[2 * c]);
  const long long start = chunks[2 * c + 1];
  const long long stop = (start + SBMC_MT_CHUNK_ELEMS < tn.n) ? start + SBMC_MT_CHUNK_ELEMS : tn.n;
  const float coef = coef_ptr ? *coef_ptr : 1.0f;
  for (long long i = start + threadIdx.x; i < stop; i += 256)
    adam_element(tn.p + i, tn.g + i, tn.m + i, tn.v + i, coef, s);
}
__global__ void step_increment_kernel(float *step) { *step += 1.0f; }

}  // namespace sbmc

extern "C" {

int sbmc_multi_tensor_adam_devstep_f32(const int64_t *tensors, const int64_t *chunks,
                                       int64_t nchunks, const float *clip_coef, double lr,
                                       double beta1, double beta2, double eps, float *step,
                                       void *stream) {
  if (nchunks < 0 || nchunks > 0x7FFFFFFF) {
    sbmc::set_error("adam: invalid chunk count %lld", (long long)nchunks);
    return SBMC_EINVAL;
  }
  if (!step || (nchunks > 0 && (!tensors || !chunks))) {
    sbmc::set_error("adam: null pointer argument");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    sbmc::KernelTimer timer(SBMC_KERNEL_OPTIM, st);
    if (nchunks > 0)
      sbmc::multi_adam_devstep_kernel<<<(unsigned)nchunks, 256, 0, st>>>(
          reinterpret_cast<const long long *>(tensors),
          reinterpret_cast<const long long *>(chunks), clip_coef, lr, beta1, beta2, eps, step);
    sbmc::step_increment_kernel<<<1, 1, 0, st>>>(step);
  }
  SBMC_CUDA_OK(cudaGetLastError());
  sbmc::count_launch(2);
  sbmc::note_path(1);
  return SBMC_OK;
}

int sbmc_multi_tensor_grad_norm_f32(const int64_t *tensors, const int64_t *chunks,
                                    int64_t nchunks, float *partial, float max_norm,
                                    float *norm_and_coef, void *stream) {
  if (nchunks < 0 || nchunks > 0x7FFFFFFF) {
    sbmc::set_error("grad_norm: invalid chunk count %lld", (long long)nchunks);
    return SBMC_EINVAL;
  }
  if (!norm_and_coef || (nchunks > 0 && (!tensors || !chunks || !partial))) {
    sbmc::set_error("grad_norm: null pointer argument");
    return SBMC_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sbmc::KernelTimer timer(SBMC_KERNEL_OPTIM, st);
  if (nchunks > 0) {
    sbmc::multi_sqnorm_kernel<<<(unsigned)nchunks, 256, 0, st>>>(
        reinterpret_cast<const long long *>(tensors), reinterpret_cast<const long long *>(chunks),
        partial);
    SBMC_CUDA_OK(cudaGetLastError());
    sbmc::count_launch();
  }
  sbmc::finalize_norm_kernel<<<1, 256, 0, st>>>(partial, nchunks, max_norm, norm_and_coef);
  SBMC_CUDA_OK(cudaGetLastError());
  sbmc::count_launch();
  sbmc::note_path(1);
  return SBMC_OK;
}

int sbmc_multi_tensor_adam_f32(const int64_t *tensors, const int64_t *chunks, int64_t nchunks,
                               const float *clip_coef, double lr, double beta1, double beta2,
                               double eps, double bias_correction1,
                               double bias_correction2_sqrt, void *stream) {
  if (nchunks < 0 || nchunks > 0x7FFFFFFF) {
    sbmc::set_error("adam: invalid chunk count %lld", (long long)nchunks);
    return SBMC_EINVAL;
  }
  if (nchunks == 0) return SBMC_OK;
  if (!tensors || !chunks) {
    sbmc::set_error("adam: null pointer argument");
    return SBMC_EINVAL;
  }
  if (!(bias_correction1 > 0.0) || !(bias_correction2_sqrt > 0.0)) {
    sbmc::set_error("adam: bias corrections must be positive (step >= 1)");
    return SBMC_EINVAL;
  }
  const sbmc::AdamScalars s =
      sbmc::adam_scalars(lr, beta1, beta2, eps, bias_correction1, bias_correction2_sqrt);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    sbmc::KernelTimer timer(SBMC_KERNEL_OPTIM, st);
    sbmc::multi_adam_kernel<<<(unsigned)nchunks, 256, 0, st>>>(
        reinterpret_cast<const long long *>(tensors), reinterpret_cast<const long long *>(chunks),
        clip_coef, s);
  }
  SBMC_CUDA_OK(cudaGetLastError());
  sbmc::count_launch();
  sbmc::note_path(1);
  return SBMC_OK;
}

}  // extern "C"
