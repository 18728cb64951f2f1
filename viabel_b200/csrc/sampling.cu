// Counter-based base draws (Philox4x32-10): normal, chi-square, Student-t.
// Replaces numpy RandomState.randn / chisquare / standard_t (reference approximations.py:216,
// :274, :345-347).  The numpy MT19937 streams are NOT reproduced; parity is by draw injection.
#include "philox_draws.cuh"

namespace vb {

template <typename T>
__global__ void philox_normal_kernel(T* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset,
                                     int quantize) {
  PDL_SYNC();
  const Philox ph(seed);
  const int64_t pairs = (n + 1) / 2;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pairs;
       p += (int64_t)gridDim.x * blockDim.x) {
    // element i uses counter (offset + i) / 2: a run that starts at an even offset is a
    // contiguous slice of the same infinite stream
    const uint64_t e0 = offset + 2 * (uint64_t)p;
    double z0, z1;
    normal_pair(ph, e0 >> 1, 0, z0, z1);
    if (e0 & 1) {  // odd offset: element e0 is the second of its pair, e0+1 the first of the next
      double a0, a1;
      normal_pair(ph, (e0 >> 1) + 1, 0, a0, a1);
      z0 = z1;
      z1 = a0;
    }
    if (quantize) {
      z0 = quantize_draw(z0, quantize);
      z1 = quantize_draw(z1, quantize);
    }
    out[2 * p] = (T)z0;
    if (2 * p + 1 < n) out[2 * p + 1] = (T)z1;
  }
}

__global__ void philox_chisquare_kernel(double* __restrict__ out, int64_t n, double df, uint64_t seed,
                                        uint64_t offset) {
  const Philox ph(seed);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = 2.0 * gamma_draw(ph, offset + (uint64_t)i, 1, 0.5 * df);
}

__global__ void philox_student_t_kernel(double* __restrict__ out, int64_t n, double df, uint64_t seed,
                                        uint64_t offset, int quantize) {
  const Philox ph(seed);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t e = offset + (uint64_t)i;
    double z, unused;
    normal_pair(ph, e, 2, z, unused);
    const double chi2 = 2.0 * gamma_draw(ph, e, 3, 0.5 * df);
    double tv = z / sqrt(chi2 / df);
    out[i] = quantize ? quantize_draw(tv, quantize) : tv;
  }
}

static inline int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vb
using namespace vb;

extern "C" int vb_philox_normal_f64(double* out, int64_t n, uint64_t seed, uint64_t offset, int quantize,
                                    cudaStream_t stream) {
  if (n < 0 || (n > 0 && !out)) return set_error(VB_ERR_INVALID_ARG, "philox_normal: bad arguments");
  if (n == 0) return VB_OK;
  VB_CUDA(launch_pdl(philox_normal_kernel<double>, dim3(grid_for((n + 1) / 2, 256)), dim3(256), stream, out, n, seed, offset, quantize));
  return VB_OK;
}

extern "C" int vb_philox_normal_f32(float* out, int64_t n, uint64_t seed, uint64_t offset, int quantize,
                                    cudaStream_t stream) {
  if (n < 0 || (n > 0 && !out)) return set_error(VB_ERR_INVALID_ARG, "philox_normal: bad arguments");
  if (n == 0) return VB_OK;
  VB_CUDA(launch_pdl(philox_normal_kernel<float>, dim3(grid_for((n + 1) / 2, 256)), dim3(256), stream, out, n, seed, offset, quantize));
  return VB_OK;
}

extern "C" int vb_philox_chisquare_f64(double* out, int64_t n, double df, uint64_t seed, uint64_t offset,
                                       cudaStream_t stream) {
  if (n < 0 || (n > 0 && !out) || !(df > 0)) return set_error(VB_ERR_INVALID_ARG, "philox_chisquare: bad arguments");
  if (n == 0) return VB_OK;
  philox_chisquare_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, n, df, seed, offset);
  VB_CHECK_LAUNCH();
  return VB_OK;
}

extern "C" int vb_philox_student_t_f64(double* out, int64_t n, double df, uint64_t seed, uint64_t offset,
                                       int quantize, cudaStream_t stream) {
  if (n < 0 || (n > 0 && !out) || !(df > 0)) return set_error(VB_ERR_INVALID_ARG, "philox_student_t: bad arguments");
  if (n == 0) return VB_OK;
  philox_student_t_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, n, df, seed, offset, quantize);
  VB_CHECK_LAUNCH();
  return VB_OK;
}
