// Internal interface of gemm_f64.cu (float64 DMMA GEMM with fused scaling / bias / division epilogues).
#pragma once
#include "common.cuh"

namespace vb {

struct GemmArgs {
  int M, N, K;
  double alpha;
  const double* A;
  int64_t lda;
  const double* B;
  int64_t ldb;
  double* C;
  int64_t ldc;
  const double* kscale;     // [K] or NULL: A'(m,k) is multiplied by kscale[k]
  const double* rowscale;   // [M] or NULL
  const double* bias;       // [N] or NULL
  const double* divm;       // [M], with divn [N], or both NULL: divide by (divm[m] + divn[n])
  const double* divn;
};

int gemm_f64(const GemmArgs& g, bool trans_a, bool trans_b, cudaStream_t stream);

}  // namespace vb
