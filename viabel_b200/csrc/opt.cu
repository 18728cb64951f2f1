// Optimiser steps fused with `var_param -= lr * direction`
// (reference optimization.py:188-197 RMSProp, :308-326 Adam, objectives.py:57-59 update).
#include "common.cuh"

namespace vb {

__global__ void rmsprop_step_kernel(double* __restrict__ vp, const double* __restrict__ grad,
                                    double* __restrict__ nu, double* __restrict__ dir, int64_t P, double lr,
                                    double beta, double jitter, int first) {
  PDL_SYNC();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double g = grad[i], g2 = g * g;
    double v = first ? g2 : nu[i];               // state starts at grad**2 (optimization.py:189-190)
    v = v * beta;
    v += (1.0 - beta) * g2;
    nu[i] = v;
    const double dd = g / sqrt(jitter + v);
    if (dir) dir[i] = dd;
    vp[i] -= lr * dd;
  }
}

__global__ void adam_step_kernel(double* __restrict__ vp, const double* __restrict__ grad,
                                 double* __restrict__ m, double* __restrict__ nu, double* __restrict__ dir,
                                 int64_t P, double lr, double beta1, double beta2, double jitter, int first) {
  PDL_SYNC();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double g = grad[i];
    double mi, vi;
    if (first) {
      // optimization.py:314-320: `momentum = grad` aliases the gradient array, so the in-place
      // `momentum *= beta1; momentum += (1-beta1)*grad` sees the already-scaled gradient, and
      // `grad**2` in the nu update is momentum**2.
      double gs = g * beta1;
      mi = gs + (1.0 - beta1) * gs;
      vi = (g * g) * beta2;
      vi += (1.0 - beta2) * (mi * mi);
    } else {
      mi = m[i] * beta1;
      mi += (1.0 - beta1) * g;
      vi = nu[i] * beta2;
      vi += (1.0 - beta2) * (g * g);
    }
    m[i] = mi;
    nu[i] = vi;
    const double dd = mi / sqrt(jitter + vi);
    if (dir) dir[i] = dd;
    vp[i] -= lr * dd;
  }
}

}  // namespace vb
using namespace vb;

extern "C" int vb_rmsprop_step_f64(double* var_param, const double* grad, double* nu, double* direction,
                                   int64_t P, double lr, double beta, double jitter, int first,
                                   cudaStream_t stream) {
  if (P <= 0 || !var_param || !grad || !nu) return set_error(VB_ERR_INVALID_ARG, "rmsprop_step: bad arguments");
  int blocks = (int)((P + 255) / 256);
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  VB_CUDA(launch_pdl(rmsprop_step_kernel, dim3(blocks), dim3(256), stream, var_param, grad, nu, direction, P, lr, beta, jitter, first));
  return VB_OK;
}

extern "C" int vb_adam_step_f64(double* var_param, const double* grad, double* m, double* nu, double* direction,
                                int64_t P, double lr, double beta1, double beta2, double jitter, int first,
                                cudaStream_t stream) {
  if (P <= 0 || !var_param || !grad || !m || !nu) return set_error(VB_ERR_INVALID_ARG, "adam_step: bad arguments");
  int blocks = (int)((P + 255) / 256);
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  VB_CUDA(launch_pdl(adam_step_kernel, dim3(blocks), dim3(256), stream, var_param, grad, m, nu, direction, P, lr, beta1, beta2, jitter, first));
  return VB_OK;
}
