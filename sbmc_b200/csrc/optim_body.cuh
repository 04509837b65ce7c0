// optim_body.cuh -- per-element bodies of the fused optimizer step (optim.cu),
// plain functions so that tests/native/optim_emul.cpp runs the same arithmetic
// on the CPU against torch.optim.Adam + clip_grad_norm_.
//
// Reference: the training step of sbmc/interfaces.py:78-106 -- gradient-norm
// clipping at 1000 (`th.nn.utils.clip_grad_norm_`) followed by `Adam.step()` on
// every parameter of the model (~150 tensors, 35 M elements for Multisteps):
// hundreds of small launches in eager PyTorch, three here.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SBMC_OPT_HD __host__ __device__ __forceinline__
#else
#define SBMC_OPT_HD static inline
#include <math.h>
#endif

namespace sbmc {

#include "../../include/sbmc_b200.h"  // SBMC_MT_CHUNK_ELEMS: elements of one tensor per CTA

// Row of the tensor table (int64 [ntensors][5]): param, grad, exp_avg,
// exp_avg_sq pointers and the element count.
struct MtTensor {
  float *p, *g, *m, *v;
  long long n;
};

SBMC_OPT_HD MtTensor mt_tensor(const long long *tensors, long long t) {
  MtTensor r;
  r.p = reinterpret_cast<float *>(static_cast<uintptr_t>(tensors[5 * t + 0]));
  r.g = reinterpret_cast<float *>(static_cast<uintptr_t>(tensors[5 * t + 1]));
  r.m = reinterpret_cast<float *>(static_cast<uintptr_t>(tensors[5 * t + 2]));
  r.v = reinterpret_cast<float *>(static_cast<uintptr_t>(tensors[5 * t + 3]));
  r.n = tensors[5 * t + 4];
  return r;
}

// Derived in double on the host, like torch derives them in Python floats
// (1 - 0.999 evaluated in fp32 is off by 1e-5 relative).
struct AdamScalars {
  float lr_over_bc1;          // lr / (1 - beta1^t)
  float beta2, eps;
  float one_minus_beta1, one_minus_beta2;
  float bc2_sqrt;             // sqrt(1 - beta2^t)
};

SBMC_OPT_HD AdamScalars adam_scalars(double lr, double beta1, double beta2, double eps,
                                       double bias_correction1, double bias_correction2_sqrt) {
  AdamScalars s;
  s.lr_over_bc1 = (float)(lr / bias_correction1);
  s.beta2 = (float)beta2;
  s.eps = (float)eps;
  s.one_minus_beta1 = (float)(1.0 - beta1);
  s.one_minus_beta2 = (float)(1.0 - beta2);
  s.bc2_sqrt = (float)bias_correction2_sqrt;
  return s;
}

// One element of torch.optim.Adam's update (no weight decay, no amsgrad) on a
// gradient scaled by the clipping coefficient `coef` (1 = no clipping).  The
// scaled gradient is written back when it changed, as clip_grad_norm_ does.
SBMC_OPT_HD void adam_element(float *p, float *g, float *m, float *v, float coef,
                              const AdamScalars &s) {
  float grad = *g;
  if (coef != 1.0f) {
    grad *= coef;
    *g = grad;
  }
  const float m_new = *m + (grad - *m) * s.one_minus_beta1;            // lerp_
  const float v_new = *v * s.beta2 + (s.one_minus_beta2 * grad) * grad;   // mul_, addcmul_
  *m = m_new;
  *v = v_new;
  const float denom = sqrtf(v_new) / s.bc2_sqrt + s.eps;
  *p = *p - s.lr_over_bc1 * (m_new / denom);                          // addcdiv_
}

}  // namespace sbmc
