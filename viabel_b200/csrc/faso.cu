// Device-side convergence statistics of FASO / RAABBVI (reference optimization.py:550-605 and
// _mc_diagnostics.py:7-184), computed on the iterate history ring that the fused step writes
// (vb_step_buffers.param_hist), batched over all P parameters:
//   vb_faso_rhat_f64     split-R-hat of the last W iterates for several window sizes W at once
//                        (compute_R_hat :124-160, R_hat_convergence_check :163-184) -> max over parameters per window
//   vb_ring_mean_f64     iterate average over the last W rows (optimization.py:563, :568)
//   vb_faso_center_f64   window minus its column means, zero padded to the FFT length, plus the ddof=1 variances
//                        (autocov :21-23, MCSE :119); the FFT itself is a library call (cuFFT through torch.fft)
//   vb_faso_ess_f64      Geyer initial positive / monotone sequence estimator per parameter (ess :56-99)
// Rows of the ring are addressed logically: row i of a window that ends at physical row `end` (exclusive) and
// holds W rows is physical row (end - W + i) mod ring.
#include "common.cuh"

namespace vb {

constexpr int kMaxSeg = 32;
constexpr int kSegGroups = 32;       // row groups per block: block = 32 columns x 32 groups

struct Segs {
  int nseg;
  long long start[kMaxSeg];          // physical row of the segment's first row
  long long count[kMaxSeg];
};

__device__ __forceinline__ long long ring_row(long long start, long long i, long long ring) {
  long long r = start + i;
  return r >= ring ? r - ring : r;
}

// MODE 0: mean[seg][p] = mean of the segment's rows;  MODE 1: css[seg][p] = sum (x - mean)^2
template <int MODE>
__global__ void __launch_bounds__(32 * kSegGroups) ring_seg_kernel(const double* __restrict__ hist, long long ring, int P,
                                                                   Segs sg, const double* __restrict__ mean_in,
                                                                   double* __restrict__ out) {
  __shared__ double sm[kSegGroups][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + x, seg = blockIdx.y;
  const long long cnt = sg.count[seg], st = sg.start[seg];
  double acc = 0.0;
  if (p < P) {
    const double mu = MODE == 1 ? mean_in[(size_t)seg * P + p] : 0.0;
#pragma unroll 4
    for (long long i = y; i < cnt; i += kSegGroups) {
      const double v = hist[(size_t)ring_row(st, i, ring) * P + p];
      if (MODE == 0) acc += v;
      else acc += (v - mu) * (v - mu);
    }
  }
  sm[y][x] = acc;
  __syncthreads();
  if (y == 0 && p < P) {
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < kSegGroups; ++r) t += sm[r][x];
    out[(size_t)seg * P + p] = MODE == 0 ? t / (double)cnt : t;
  }
}

// one block per window: R-hat per parameter from the two halves' means and variances, max over parameters
__global__ void __launch_bounds__(256) rhat_finish_kernel(const double* __restrict__ mean, const double* __restrict__ css,
                                                          int P, Segs sg, double jitter, double* __restrict__ rhat_max) {
  __shared__ double red[32];
  const int w = blockIdx.x;
  const double half = (double)sg.count[2 * w];
  double mx = -INFINITY;
  bool any_nan = false;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const double m1 = mean[(size_t)(2 * w) * P + p], m2 = mean[(size_t)(2 * w + 1) * P + p];
    const double v1 = css[(size_t)(2 * w) * P + p] / (half - 1.0), v2 = css[(size_t)(2 * w + 1) * P + p] / (half - 1.0);
    const double grand = (m1 + m2) / 2.0;
    const double B = half * ((m1 - grand) * (m1 - grand) + (m2 - grand) * (m2 - grand)) / (2 - 1);
    // np.nanmean over the two halves' variances (:156)
    double W;
    if (isnan(v1) && isnan(v2)) W = NAN;
    else if (isnan(v1)) W = v2;
    else if (isnan(v2)) W = v1;
    else W = (v1 + v2) / 2.0;
    W += jitter;
    const double r = sqrt((half - 1.0) / half + B / (half * W));
    if (isnan(r)) any_nan = true;            // np.max propagates NaN
    mx = fmax(mx, r);
  }
  mx = block_max(mx, red);
  const int nan_any = __syncthreads_or(any_nan ? 1 : 0);
  if (threadIdx.x == 0) rhat_max[w] = nan_any ? NAN : mx;
}

// out[r][p] = hist[row r of the window][p] - mean[p] (r < W), 0 (W <= r < m)
__global__ void ring_center_kernel(const double* __restrict__ hist, long long ring, int P, long long start, long long W,
                                   long long m, const double* __restrict__ mean, double* __restrict__ out) {
  const long long total = m * (long long)P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / P;
    const int p = (int)(i - r * P);
    out[i] = r < W ? hist[(size_t)ring_row(start, r, ring) * P + p] - mean[p] : 0.0;
  }
}

// Geyer's estimator on acov[t][p] = scale * raw[t * ld + p], one thread per parameter (coalesced across parameters).
// rho is built in place over raw[] (entry t is only read before it is written).
__global__ void geyer_ess_kernel(double* __restrict__ raw, long long ld, double scale, long long n_draw, int P,
                                 double* __restrict__ ess) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double* a = raw + p;
  const double nd = (double)n_draw;
  const double mean_var = a[0] * scale * nd / (nd - 1.0);
  const double var_plus = mean_var * (nd - 1.0) / nd;
  double even = 1.0;
  double odd = 1.0 - (mean_var - a[ld] * scale) / var_plus;
  const bool nan_seen = isnan(odd);     // the reference's isnan(rho).any(): only rho[1] can be NaN (the others are stored
                                        // only when their pair sum compares >= 0)
  a[0] = even;
  a[ld] = odd;
  long long t = 1;
  while (t < n_draw - 3 && (even + odd) > 0.0) {
    even = 1.0 - (mean_var - a[(t + 1) * ld] * scale) / var_plus;
    odd = 1.0 - (mean_var - a[(t + 2) * ld] * scale) / var_plus;
    if ((even + odd) >= 0) {
      a[(t + 1) * ld] = even;
      a[(t + 2) * ld] = odd;
    } else {
      a[(t + 1) * ld] = 0.0;
      a[(t + 2) * ld] = 0.0;
    }
    t += 2;
  }
  const long long max_t = t - 2;
  if (even > 0) a[(max_t + 1) * ld] = even;
  t = 1;
  while (t <= max_t - 2) {
    if (a[(t + 1) * ld] + a[(t + 2) * ld] > a[(t - 1) * ld] + a[t * ld]) {
      a[(t + 1) * ld] = (a[(t - 1) * ld] + a[t * ld]) / 2.0;
      a[(t + 2) * ld] = a[(t + 1) * ld];
    }
    t += 2;
  }
  double s = 0.0;
  for (long long i = 0; i <= max_t; ++i) s += a[i * ld];
  double tau = -1.0 + 2.0 * s + a[(max_t + 1) * ld];
  tau = fmax(tau, 1.0 / log10(nd));
  ess[p] = nan_seen ? NAN : nd / tau;
}

static int fill_window_segs(long long ring, long long end, const int64_t* windows, int nwin, Segs& sg) {
  sg.nseg = 2 * nwin;
  for (int w = 0; w < nwin; ++w) {
    long long W = windows[w];
    if (W < 4 || W > ring) return set_error(VB_ERR_INVALID_ARG, "faso_rhat: window must have 4 <= W <= ring rows");
    long long start = ((end - W) % ring + ring) % ring;       // first row of the window (oldest)
    if (W & 1) W -= 1;                                        // odd: the most recent row is dropped (:141-143)
    const long long half = W / 2;
    sg.start[2 * w] = start;
    sg.count[2 * w] = half;
    sg.start[2 * w + 1] = (start + half) % ring;
    sg.count[2 * w + 1] = half;
  }
  return VB_OK;
}

}  // namespace vb
using namespace vb;

/* scratch: 2 * (2 nwin) * P doubles */
extern "C" size_t vb_faso_rhat_workspace_bytes(int P, int nwin) {
  if (P <= 0 || nwin <= 0 || 2 * nwin > kMaxSeg) return 0;
  return sizeof(double) * 4 * (size_t)nwin * P;
}

extern "C" int vb_faso_rhat_f64(const double* hist, int64_t ring, int P, int64_t end, const int64_t* windows_host, int nwin,
                                double jitter, double* rhat_max, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!hist || ring <= 0 || P <= 0 || !windows_host || nwin <= 0 || 2 * nwin > kMaxSeg || !rhat_max)
    return set_error(VB_ERR_INVALID_ARG, "faso_rhat: bad arguments (at most 16 windows)");
  if (!workspace || workspace_bytes < vb_faso_rhat_workspace_bytes(P, nwin)) return set_error(VB_ERR_WORKSPACE, "faso_rhat: workspace too small");
  Segs sg;
  int rc = fill_window_segs(ring, end, windows_host, nwin, sg);
  if (rc) return rc;
  double* mean = static_cast<double*>(workspace);
  double* css = mean + (size_t)2 * nwin * P;
  const dim3 grid((P + 31) / 32, 2 * nwin);
  ring_seg_kernel<0><<<grid, 32 * kSegGroups, 0, stream>>>(hist, ring, P, sg, nullptr, mean);
  VB_CHECK_LAUNCH();
  ring_seg_kernel<1><<<grid, 32 * kSegGroups, 0, stream>>>(hist, ring, P, sg, mean, css);
  VB_CHECK_LAUNCH();
  rhat_finish_kernel<<<nwin, 256, 0, stream>>>(mean, css, P, sg, jitter, rhat_max);
  VB_CHECK_LAUNCH();
  return VB_OK;
}

/* mean[P] (and css[P] = sum (x - mean)^2 when css != NULL) over the W rows that end at physical row `end` */
extern "C" int vb_ring_mean_f64(const double* hist, int64_t ring, int P, int64_t end, int64_t W, double* mean, double* css,
                                cudaStream_t stream) {
  if (!hist || ring <= 0 || P <= 0 || W <= 0 || W > ring || !mean) return set_error(VB_ERR_INVALID_ARG, "ring_mean: bad arguments");
  Segs sg;
  sg.nseg = 1;
  sg.start[0] = ((end - W) % ring + ring) % ring;
  sg.count[0] = W;
  const dim3 grid((P + 31) / 32, 1);
  ring_seg_kernel<0><<<grid, 32 * kSegGroups, 0, stream>>>(hist, ring, P, sg, nullptr, mean);
  VB_CHECK_LAUNCH();
  if (css) {
    ring_seg_kernel<1><<<grid, 32 * kSegGroups, 0, stream>>>(hist, ring, P, sg, mean, css);
    VB_CHECK_LAUNCH();
  }
  return VB_OK;
}

/* centered[m, P]: the window minus `mean`, zero padded to m rows (the FFT length next_fast_len(2 W)) */
extern "C" int vb_faso_center_f64(const double* hist, int64_t ring, int P, int64_t end, int64_t W, int64_t m, const double* mean,
                                  double* centered, cudaStream_t stream) {
  if (!hist || ring <= 0 || P <= 0 || W <= 0 || W > ring || m < W || !mean || !centered)
    return set_error(VB_ERR_INVALID_ARG, "faso_center: bad arguments");
  const long long start = ((end - W) % ring + ring) % ring;
  const long long total = m * (long long)P;
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  ring_center_kernel<<<(unsigned)blocks, 256, 0, stream>>>(hist, ring, P, start, W, m, mean, centered);
  VB_CHECK_LAUNCH();
  return VB_OK;
}

/* acov_raw[t * ld + p] * scale = autocovariance at lag t of parameter p (t < n_draw); overwritten with the rho sequence */
extern "C" int vb_faso_ess_f64(double* acov_raw, int64_t ld, double scale, int64_t n_draw, int P, double* ess, cudaStream_t stream) {
  if (!acov_raw || ld < P || n_draw < 4 || P <= 0 || !ess) return set_error(VB_ERR_INVALID_ARG, "faso_ess: bad arguments (n_draw >= 4)");
  geyer_ess_kernel<<<(P + 127) / 128, 128, 0, stream>>>(acov_raw, ld, scale, n_draw, P, ess);
  VB_CHECK_LAUNCH();
  return VB_OK;
}
