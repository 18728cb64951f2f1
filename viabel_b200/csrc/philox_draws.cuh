// Device-side base-draw primitives shared by sampling.cu (stand-alone draw kernels) and engine.cu (draws fused
// into the step's first kernel).  Element e of a stream is a pure function of (seed, e), so any kernel may
// regenerate any sub-range in any thread order and every rank draws identical values.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace vb {

// quantize: 1 = bfloat16 (8-bit mantissa), 2 = float16 (11-bit mantissa); both are exact operands
// of the fp16 tensor-core path (a bf16 value in the normal fp16 range is an fp16 value)
__device__ __forceinline__ double quantize_draw(double x, int mode) {
  if (mode == 2) return (double)__half2float(__float2half_rn((float)x));
  return (double)__bfloat162float(__float2bfloat16_rn((float)x));
}

// two independent N(0,1) from one Philox block (Box-Muller, 53-bit uniforms)
__device__ __forceinline__ void normal_pair(const Philox& ph, uint64_t ctr, uint64_t stream_id, double& z0,
                                            double& z1) {
  uint32_t r[4];
  ph(ctr, stream_id, r);
  const double u1 = u01_53(r[0], r[1]), u2 = u01_53(r[2], r[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

// Marsaglia-Tsang gamma(shape a >= 1/3 boosted), counter = element, attempts on the high word
__device__ inline double gamma_draw(const Philox& ph, uint64_t elem, uint64_t stream_id, double a) {
  double boost = 1.0;
  if (a < 1.0) {
    uint32_t r[4];
    ph(elem, stream_id | (1ull << 62), r);
    boost = pow(u01_53(r[0], r[1]), 1.0 / a);
    a += 1.0;
  }
  const double dd = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
  for (uint64_t attempt = 0; attempt < 64; ++attempt) {
    double x, unused;
    normal_pair(ph, elem, stream_id | ((2 * attempt + 1) << 40), x, unused);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    uint32_t r[4];
    ph(elem, stream_id | ((2 * attempt + 2) << 40), r);
    const double u = u01_53(r[0], r[1]);
    if (log(u) < 0.5 * x * x + dd - dd * v + dd * log(v)) return boost * dd * v;
  }
  return boost * dd;  // unreachable in practice (acceptance > 95% per attempt)
}

// element e of the standard-normal stream (vb_philox_normal_*): counter e >> 1, component e & 1
__device__ __forceinline__ double normal_element(const Philox& ph, uint64_t e, int quantize) {
  double z0, z1;
  normal_pair(ph, e >> 1, 0, z0, z1);
  const double z = (e & 1) ? z1 : z0;
  return quantize ? quantize_draw(z, quantize) : z;
}
// element e of the Student-t stream (vb_philox_student_t_f64)
__device__ __forceinline__ double student_element(const Philox& ph, uint64_t e, double df, int quantize) {
  double z, unused;
  normal_pair(ph, e, 2, z, unused);
  const double chi2 = 2.0 * gamma_draw(ph, e, 3, 0.5 * df);
  const double tv = z / sqrt(chi2 / df);
  return quantize ? quantize_draw(tv, quantize) : tv;
}

}  // namespace vb
