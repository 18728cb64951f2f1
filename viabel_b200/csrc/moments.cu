// Sample moments and the weighted symmetric rank-k update, float64.
//
//   vb_sample_moments_f64   per-coordinate mean and central power sums sum_n (x_nj - mean_j)^{2,4} -- the
//                           sample-moment branch of wasserstein_bounds (diagnostics.py:137-141: mean over draws of
//                           sum_j (x_j - xbar_j)^p, per-coordinate powers) -- and the d x d sample covariance that
//                           all_diagnostics takes from np.cov(samples.T) (diagnostics.py:58-59).
//   syrk_f64 (internal)     out = sum_n w_n (x_n - c)(x_n - c)^T on the FP64 tensor pipe (mma.sync m8n8k4, DMMA):
//                           64 x 64 output tiles of the upper triangle, split over row chunks, deterministic
//                           two-stage sum.  Used for the covariance (w = 1, c = mean) and for the GLM Hessian
//                           X^T diag(c) X of the control-variate ExclusiveKL estimators (glm_point.cu).
#include "moments_internal.cuh"

namespace vb {

// ---- column sums over row chunks: block = 32 columns x 8 row groups; grid (column blocks, chunks) --------------
// MODE 0: sum x;  MODE 1: sum (x - c)^2 and sum (x - c)^4
template <int MODE>
__global__ void __launch_bounds__(256) col_sums_kernel(const double* __restrict__ x, int64_t n, int d, int64_t ldx,
                                                       const double* __restrict__ centre, int64_t rows_per_chunk,
                                                       double* __restrict__ part_a, double* __restrict__ part_b) {
  __shared__ double sm[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < n ? r0 + rows_per_chunk : n;
  double a = 0.0, b = 0.0;
  if (j < d) {
    const double c = MODE == 1 ? centre[j] : 0.0;
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const double v = x[r * ldx + j];
      if (MODE == 0) {
        a += v;
      } else {
        const double e = v - c, e2 = e * e;
        a += e2;
        b += e2 * e2;
      }
    }
  }
  sm[0][ty][tx] = a;
  sm[1][ty][tx] = b;
  __syncthreads();
  if (ty == 0 && j < d) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) { ta += sm[0][r][tx]; tb += sm[1][r][tx]; }
    part_a[(size_t)blockIdx.y * d + j] = ta;
    if (MODE == 1) part_b[(size_t)blockIdx.y * d + j] = tb;
  }
}

// out[j] = scale * sum_chunks part[chunk][j]  (fixed order)
__global__ void col_finish_kernel(const double* __restrict__ part, int chunks, int d, double scale, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  double t = 0.0;
  for (int c = 0; c < chunks; ++c) t += part[(size_t)c * d + j];
  out[j] = t * scale;
}

// ---- SYRK on DMMA ---------------------------------------------------------------------------------------------
constexpr int kSyrkTile = 64, kSyrkK = 32, kSyrkPitch = 68;     // pitch = 4 mod 16: conflict-free fragment loads

__device__ __forceinline__ void dmma884_acc(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// grid (upper-triangle block pairs, row chunks); 8 warps as 4 (rows of the tile) x 2 (columns): warp tile 16 x 32
__global__ void __launch_bounds__(256) syrk_f64_kernel(const double* __restrict__ x, int64_t n, int d, int64_t ldx,
                                                       const double* __restrict__ w, const double* __restrict__ centre,
                                                       int64_t rows_per_chunk, int nblk, double* __restrict__ part) {
  __shared__ double sa[kSyrkK][kSyrkPitch], sb[kSyrkK][kSyrkPitch];
  // decode the upper-triangle pair index
  int bi = 0, rem = blockIdx.x;
  while (rem >= nblk - bi) { rem -= nblk - bi; ++bi; }
  const int bj = bi + rem;
  const int i0 = bi * kSyrkTile, j0 = bj * kSyrkTile;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < n ? r0 + rows_per_chunk : n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;                 // 4 x 2 warps
  const int g = lane >> 2, t = lane & 3;
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  const int lc = threadIdx.x & 63, lr = threadIdx.x >> 6;   // loader: 64 columns x 4 rows per pass
  const double ci = (centre && i0 + lc < d) ? centre[i0 + lc] : 0.0;
  const double cj = (centre && j0 + lc < d) ? centre[j0 + lc] : 0.0;
  for (int64_t rb = r0; rb < r1; rb += kSyrkK) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < kSyrkK / 4; ++p) {
      const int k = p * 4 + lr;
      const int64_t r = rb + k;
      double va = 0.0, vb2 = 0.0;
      if (r < r1) {
        const double wr = w ? w[r] : 1.0;
        if (i0 + lc < d) va = x[r * ldx + i0 + lc] - ci;
        if (j0 + lc < d) vb2 = (x[r * ldx + j0 + lc] - cj) * wr;
      }
      sa[k][lc] = va;
      sb[k][lc] = vb2;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < kSyrkK; k4 += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) af[a] = sa[k4 + t][wm * 16 + a * 8 + g];       // A[i = g][k = t]
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = sb[k4 + t][wn * 32 + b * 8 + g];       // B[k = t][j = g]
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884_acc(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  // C fragment: row g, columns 2t, 2t+1
  double* out = part + (size_t)blockIdx.y * d * d;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + wm * 16 + a * 8 + g;
      const int j = j0 + wn * 32 + b * 8 + 2 * t;
      if (i < d && j < d) out[(size_t)i * d + j] = acc[a][b][0];
      if (i < d && j + 1 < d) out[(size_t)i * d + j + 1] = acc[a][b][1];
    }
}

// out[i][j] = out[j][i] = scale * sum_chunks part[chunk][min][max] + (i == j ? diag_add : 0)   (fixed order)
__global__ void syrk_finish_kernel(const double* __restrict__ part, int chunks, int d, double scale, double diag_add,
                                   double* __restrict__ out) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)d * d) return;
  const int i = (int)(idx / d), j = (int)(idx % d);
  const int a = i < j ? i : j, b = i < j ? j : i;
  // entries of one 64 x 64 tile are only written by the pair (a/64, b/64); within a diagonal tile both triangles are
  double t = 0.0;
  for (int c = 0; c < chunks; ++c) t += part[((size_t)c * d + a) * d + b];
  out[idx] = t * scale + (i == j ? diag_add : 0.0);
}

void syrk_plan(int64_t n, int d, SyrkPlan& p) {
  p.nblk = (d + kSyrkTile - 1) / kSyrkTile;
  p.pairs = p.nblk * (p.nblk + 1) / 2;
  int chunks = (2 * sm_count() + p.pairs - 1) / p.pairs;
  const int64_t max_chunks = (n + 4 * kSyrkK - 1) / (4 * kSyrkK);
  if (chunks > max_chunks) chunks = (int)max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rows = (n + chunks - 1) / chunks;
  rows = (rows + kSyrkK - 1) / kSyrkK * kSyrkK;
  if (rows < kSyrkK) rows = kSyrkK;
  p.rows_per_chunk = rows;
  p.chunks = (int)((n + rows - 1) / rows);
  if (p.chunks < 1) p.chunks = 1;
  p.bytes = align_up(sizeof(double) * (size_t)p.chunks * d * d, 256);
}

int syrk_f64(const double* x, int64_t n, int d, int64_t ldx, const double* w, const double* centre, double scale,
             double diag_add, double* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SyrkPlan p;
  syrk_plan(n, d, p);
  if (!workspace || workspace_bytes < p.bytes) return set_error(VB_ERR_WORKSPACE, "syrk: workspace too small");
  double* part = static_cast<double*>(workspace);
  syrk_f64_kernel<<<dim3(p.pairs, p.chunks), 256, 0, stream>>>(x, n, d, ldx, w, centre, p.rows_per_chunk, p.nblk, part);
  VB_CHECK_LAUNCH();
  const int64_t total = (int64_t)d * d;
  syrk_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(part, p.chunks, d, scale, diag_add, out);
  VB_CHECK_LAUNCH();
  return VB_OK;
}

struct MomPlan {
  int chunks;
  int64_t rows_per_chunk;
  size_t off_a, off_b, off_syrk, total;
};

static void mom_plan(int64_t n, int d, int want_cov, MomPlan& p) {
  const int colblk = (d + 31) / 32;
  int chunks = (4 * sm_count() + colblk - 1) / colblk;
  const int64_t max_chunks = (n + 63) / 64;
  if (chunks > max_chunks) chunks = (int)max_chunks;
  if (chunks < 1) chunks = 1;
  p.rows_per_chunk = (n + chunks - 1) / chunks;
  p.chunks = (int)((n + p.rows_per_chunk - 1) / p.rows_per_chunk);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  p.off_a = take(sizeof(double) * (size_t)p.chunks * d);
  p.off_b = take(sizeof(double) * (size_t)p.chunks * d);
  p.off_syrk = off;
  if (want_cov) {
    SyrkPlan s;
    syrk_plan(n, d, s);
    off += s.bytes;
  }
  p.total = off;
}

}  // namespace vb
using namespace vb;

extern "C" size_t vb_sample_moments_workspace_bytes(int64_t n, int d, int want_cov) {
  if (n <= 0 || d <= 0) return 0;
  MomPlan p;
  mom_plan(n, d, want_cov, p);
  return p.total;
}

extern "C" int vb_sample_moments_f64(const double* x, int64_t n, int d, int64_t ldx, double* mean, double* m2, double* m4,
                                     double* cov, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!x || n <= 0 || d <= 0 || ldx < d || !mean) return set_error(VB_ERR_INVALID_ARG, "sample_moments: bad arguments");
  if ((m2 == nullptr) != (m4 == nullptr)) return set_error(VB_ERR_INVALID_ARG, "sample_moments: m2 and m4 come together");
  MomPlan p;
  mom_plan(n, d, cov != nullptr, p);
  if (!workspace || workspace_bytes < p.total) return set_error(VB_ERR_WORKSPACE, "sample_moments: workspace too small");
  char* ws = static_cast<char*>(workspace);
  double* pa = reinterpret_cast<double*>(ws + p.off_a);
  double* pb = reinterpret_cast<double*>(ws + p.off_b);
  const dim3 grid((d + 31) / 32, p.chunks);
  col_sums_kernel<0><<<grid, 256, 0, stream>>>(x, n, d, ldx, nullptr, p.rows_per_chunk, pa, pb);
  VB_CHECK_LAUNCH();
  col_finish_kernel<<<(d + 255) / 256, 256, 0, stream>>>(pa, p.chunks, d, 1.0 / (double)n, mean);
  VB_CHECK_LAUNCH();
  if (m2) {
    col_sums_kernel<1><<<grid, 256, 0, stream>>>(x, n, d, ldx, mean, p.rows_per_chunk, pa, pb);
    VB_CHECK_LAUNCH();
    col_finish_kernel<<<(d + 255) / 256, 256, 0, stream>>>(pa, p.chunks, d, 1.0, m2);
    VB_CHECK_LAUNCH();
    col_finish_kernel<<<(d + 255) / 256, 256, 0, stream>>>(pb, p.chunks, d, 1.0, m4);
    VB_CHECK_LAUNCH();
  }
  if (cov) {
    // np.cov: divides by n - 1 (NaN for a single draw, like numpy)
    const double scale = n > 1 ? 1.0 / (double)(n - 1) : NAN;
    int rc = syrk_f64(x, n, d, ldx, nullptr, mean, scale, 0.0, cov, ws + p.off_syrk, workspace_bytes - p.off_syrk, stream);
    if (rc) return rc;
  }
  return VB_OK;
}
