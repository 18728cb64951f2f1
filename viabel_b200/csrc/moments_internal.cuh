// Internal (non-ABI) interface of moments.cu: the float64 weighted SYRK on the DMMA pipe.
#pragma once
#include "common.cuh"

namespace vb {

struct SyrkPlan {
  int nblk, pairs, chunks;
  int64_t rows_per_chunk;
  size_t bytes;             // workspace: chunks x d x d doubles
};

void syrk_plan(int64_t n, int d, SyrkPlan& p);

// out[d,d] = scale * sum_n w_n (x_n - centre)(x_n - centre)^T + diag_add * I   (w, centre may be NULL)
int syrk_f64(const double* x, int64_t n, int d, int64_t ldx, const double* w, const double* centre, double scale,
             double diag_add, double* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace vb
