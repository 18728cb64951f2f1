// Pareto-smoothed importance sampling on the device (reference viabel/_psis.py:113-209,
// gpdfitnew :212-332, gpinv :335-377, sumlogs :380-396) plus the 2-divergence moments that
// diagnostics.divergence_bound (:148-186) needs, without ever sorting the n log-weights.
//
// The reference does a full argsort of n values to find one order statistic.  Here:
//   pass A (1 read of lw)   : global max, compaction of "candidates" >= t0 (a threshold estimated
//                             from a strided sample so that #candidates ~ 2-3x the tail length M),
//                             and sum exp(x - t0) over everything below t0;
//   small kernels           : radix-select the (M+1)-th largest among the candidates (the cutoff is
//                             an order-statistic VALUE, so this is exact), rank the <= M tail
//                             entries by counting, Zhang-Stephens GPD fit, smoothed quantiles, LSE;
//   pass B (1 read, 1 write): out = x - max - lse, tail entries replaced by their smoothed values,
//                             and sum(out), sum exp(2 out) for the CUBO / ELBO bounds.
// Algorithmic traffic: 24 bytes per draw.  If the sample-based threshold fails (candidate buffer
// overflow or fewer than M+1 candidates -- only possible for adversarial input) status=1 is
// reported and the caller re-runs in exact mode, which finds the cutoff with full radix passes.
#include <float.h>

#include "common.cuh"

namespace vb {

constexpr int kSampleMax = 1 << 17;
constexpr int kSelThreads = 1024;
constexpr int kDigitBits = 11;
constexpr int kBins = 1 << kDigitBits;

// result[] slots (doubles, device memory)
enum { R_KHAT = 0, R_SIGMA, R_N2, R_CUTOFF, R_LSE, R_MAX, R_STATUS, R_SUMV, R_SUMEXP2V, R_M, R_NCAND, R_SMOOTHED,
       R_COUNT = 16 };

struct PsisScalars {           // device-resident control block
  unsigned long long maxkey;   // order-preserving key of the global max
  unsigned long long t0key;    // candidate threshold (key)
  unsigned long long cutkey;   // key of the (M+1)-th largest raw value
  unsigned int ncand;          // candidates found by pass A
  unsigned int ntail;          // n2
  unsigned int status;         // 0 ok, 1 fast path failed (rerun exact)
  unsigned int strict;         // exact mode: candidates are strictly above t0
  double t0;                   // threshold as a double
  double maxv, cutoff, expcut; // max, shifted cutoff, exp(shifted cutoff)
  double body_below;           // sum exp(x - t0) over x below the threshold
  double body_cand;            // sum exp(x - max) over candidates that are not tail
  double tail_sum;             // sum exp(v) over (smoothed) tail values
  double k, sigma, lse;
  double sumv, sumexp2v;       // moments of v = out + lse
  int smoothed;
  int M;
  unsigned long long prefix;   // exact-mode radix state
  unsigned long long kth;
};

__device__ __forceinline__ unsigned long long dkey(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// ---------------------------------------------------------------------------------------------
// single-CTA radix select: key of the K-th largest (K >= 1) among keys[0..cnt)
// ---------------------------------------------------------------------------------------------
__device__ unsigned long long cta_select_kth_largest(const unsigned long long* __restrict__ keys, unsigned int cnt,
                                                     unsigned long long K, unsigned int* hist /*[kBins]*/,
                                                     unsigned long long* sh /*[2]*/) {
  unsigned long long prefix = 0, mask = 0;
  int shift = 64;
  while (shift > 0) {
    const int bits = shift >= kDigitBits ? kDigitBits : shift;
    shift -= bits;
    const unsigned int nb = 1u << bits;
    for (unsigned int b = threadIdx.x; b < nb; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (unsigned int i0 = 0; i0 < cnt; i0 += blockDim.x) {
      const unsigned int i = i0 + threadIdx.x;
      const bool ok = i < cnt && ((keys[i] & mask) == prefix);
      const unsigned int dig = ok ? (unsigned int)((keys[i] >> shift) & (nb - 1)) : 0xffffffffu;
      // warp-aggregated histogram update (top digits are heavily concentrated)
      const unsigned int peers = __match_any_sync(0xffffffffu, dig);
      if (ok && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[dig], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 32) {          // warp 0 scans from the top bin down
      unsigned long long cum = 0;
      unsigned long long found = ~0ull, newK = 0;
      for (int base = (int)nb - 32; base >= 0 && found == ~0ull; base -= 32) {
        const int b = base + 31 - (int)threadIdx.x;        // lane 0 = highest bin of this group
        unsigned long long c = hist[b];
        unsigned long long incl = c;                        // inclusive scan over lanes (descending bins)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
          if ((int)threadIdx.x >= o) incl += t;
        }
        const bool hit = (cum + incl >= K) && (cum + incl - c < K);
        const unsigned int ball = __ballot_sync(0xffffffffu, hit);
        if (ball) {
          const int src = __ffs(ball) - 1;
          found = (unsigned long long)__shfl_sync(0xffffffffu, b, src);
          newK = K - (__shfl_sync(0xffffffffu, cum + incl - c, src));
        }
        cum += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (threadIdx.x == 0) {
        sh[0] = found;
        sh[1] = newK;
      }
    }
    __syncthreads();
    prefix |= sh[0] << shift;
    mask |= (unsigned long long)(nb - 1) << shift;
    K = sh[1];
    __syncthreads();
  }
  return prefix;
}

// ---------------------------------------------------------------------------------------------
__global__ void psis_init_kernel(PsisScalars* sc, int M) {
  sc->maxkey = 0; sc->t0key = 0; sc->cutkey = 0; sc->ncand = 0; sc->ntail = 0; sc->status = 0; sc->strict = 0;
  sc->t0 = -INFINITY; sc->maxv = 0; sc->cutoff = 0; sc->expcut = 0; sc->body_below = 0; sc->body_cand = 0;
  sc->tail_sum = 0; sc->k = INFINITY; sc->sigma = 0; sc->lse = 0; sc->sumv = 0; sc->sumexp2v = 0;
  sc->smoothed = 0; sc->M = M; sc->prefix = 0; sc->kth = 0;
}

__global__ void psis_sample_kernel(const double* __restrict__ lw, int64_t n, int64_t stride, int m,
                                   unsigned long long* __restrict__ skeys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) skeys[i] = dkey(lw[(int64_t)i * stride]);
}

// threshold = R-th largest of the sample (R >= m means "everything is a candidate")
__global__ void __launch_bounds__(kSelThreads) psis_sample_select_kernel(const unsigned long long* __restrict__ skeys,
                                                                          int m, unsigned int R, PsisScalars* sc) {
  __shared__ unsigned int hist[kBins];
  __shared__ unsigned long long sh[2];
  if (R >= (unsigned)m) {
    if (threadIdx.x == 0) { sc->t0key = 0; sc->t0 = -INFINITY; }
    return;
  }
  const unsigned long long key = cta_select_kth_largest(skeys, (unsigned)m, R, hist, sh);
  if (threadIdx.x == 0) { sc->t0key = key; sc->t0 = dkey_inv(key); }
}

// pass A: max, candidate compaction, sum exp(x - t0) below the threshold
__global__ void __launch_bounds__(256) psis_pass_a_kernel(const double* __restrict__ lw, int64_t n, PsisScalars* sc,
                                                          double* __restrict__ cand_x, int64_t* __restrict__ cand_i,
                                                          unsigned long long* __restrict__ cand_k, unsigned int cap,
                                                          double* __restrict__ blk_sum) {
  __shared__ double red[32];
  __shared__ unsigned long long redk[32];
  const double t0 = sc->t0;
  const bool strict = sc->strict != 0;
  const bool all = !(t0 > -INFINITY);       // threshold -inf: everything is a candidate
  double mx = -INFINITY, acc = 0.0;
  const int lane = threadIdx.x & 31;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n4 = n / 4;
  auto handle = [&](double x, int64_t i, bool valid) {
    const bool is_cand = valid && (all || (strict ? (x > t0) : (x >= t0)));
    if (valid) mx = fmax(mx, x);
    if (valid && !is_cand) acc += exp(x - t0);
    const unsigned int ball = __ballot_sync(0xffffffffu, is_cand);
    if (ball) {
      unsigned int base = 0;
      const int leader = __ffs(ball) - 1;
      if (lane == leader) base = atomicAdd(&sc->ncand, __popc(ball));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (is_cand) {
        const unsigned int slot = base + __popc(ball & ((1u << lane) - 1));
        if (slot < cap) {
          cand_x[slot] = x;
          cand_i[slot] = i;
          cand_k[slot] = dkey(x);
        }
      }
    }
  };
  const bool aligned = (reinterpret_cast<uintptr_t>(lw) & 31) == 0;
  if (aligned) {
    // 2 x LDG.128 per thread per iteration (4 doubles), warp-uniform trip count
    const int64_t iters = (n4 + nthreads - 1) / nthreads;
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t q = it * nthreads + tid;
      const bool v = q < n4;
      double2 a = make_double2(0, 0), b = make_double2(0, 0);
      if (v) {
        a = __ldcs(reinterpret_cast<const double2*>(lw) + 2 * q);
        b = __ldcs(reinterpret_cast<const double2*>(lw) + 2 * q + 1);
      }
      handle(a.x, 4 * q, v);
      handle(a.y, 4 * q + 1, v);
      handle(b.x, 4 * q + 2, v);
      handle(b.y, 4 * q + 3, v);
    }
    const int64_t rem0 = n4 * 4;
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      const int64_t i = rem0 + lane;
      handle(i < n ? lw[i] : 0.0, i, i < n);
    }
  } else {
    const int64_t iters = (n + nthreads - 1) / nthreads;
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t i = it * nthreads + tid;
      handle(i < n ? lw[i] : 0.0, i, i < n);
    }
  }
  // block reductions
  const int w = threadIdx.x >> 5;
  acc = warp_sum(acc);
  unsigned long long mk = dkey(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, mk, o);
    mk = t > mk ? t : mk;
  }
  if (lane == 0) { red[w] = acc; redk[w] = mk; }
  __syncthreads();
  if (w == 0) {
    double a = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    unsigned long long k = lane < (blockDim.x >> 5) ? redk[lane] : 0ull;
    a = warp_sum(a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long t = __shfl_xor_sync(0xffffffffu, k, o);
      k = t > k ? t : k;
    }
    if (lane == 0) {
      blk_sum[blockIdx.x] = a;
      atomicMax(&sc->maxkey, k);
    }
  }
}

// select the cutoff among the candidates, derive max / cutoff / exp(cutoff), fold the block sums
__global__ void __launch_bounds__(kSelThreads) psis_cutoff_kernel(PsisScalars* sc, const unsigned long long* __restrict__ cand_k,
                                                                   unsigned int cap, const double* __restrict__ blk_sum,
                                                                   int nblk) {
  __shared__ unsigned int hist[kBins];
  __shared__ unsigned long long sh[2];
  __shared__ double red[32];
  const unsigned int C = sc->ncand;
  const int M = sc->M;
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) s += blk_sum[b];
  s = block_sum(s, red);
  const double maxv = dkey_inv(sc->maxkey);
  unsigned long long cutkey;
  if (sc->strict) {
    cutkey = sc->t0key;                        // exact mode: t0 IS the (M+1)-th largest value
    if (C > cap) { if (threadIdx.x == 0) sc->status = 2; return; }
  } else {
    if (C > cap || C < (unsigned)(M + 1)) {
      if (threadIdx.x == 0) sc->status = 1;
      return;
    }
    cutkey = cta_select_kth_largest(cand_k, C, (unsigned long long)(M + 1), hist, sh);
  }
  if (threadIdx.x == 0) {
    const double cutraw = dkey_inv(cutkey);
    const double cutoff = fmax(cutraw - maxv, log(DBL_MIN));       // _psis.py:159, :170-173
    sc->cutkey = cutkey;
    sc->maxv = maxv;
    sc->cutoff = cutoff;
    sc->expcut = exp(cutoff);
    sc->body_below = s;
  }
}

// split candidates into tail (shifted value > cutoff) and body; accumulate the body part of the LSE
__global__ void psis_tail_compact_kernel(PsisScalars* sc, const double* __restrict__ cand_x,
                                         const int64_t* __restrict__ cand_i, double* __restrict__ tail_x,
                                         int64_t* __restrict__ tail_i, unsigned int tail_cap) {
  __shared__ double red[32];
  if (sc->status) return;
  const unsigned int C = sc->ncand;
  const double maxv = sc->maxv, cutoff = sc->cutoff;
  double acc = 0.0;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < C; i += gridDim.x * blockDim.x) {
    const double v = cand_x[i] - maxv;
    if (v > cutoff) {
      const unsigned int slot = atomicAdd(&sc->ntail, 1u);
      if (slot < tail_cap) {
        tail_x[slot] = v;
        tail_i[slot] = cand_i[i];
      }
    } else {
      acc += exp(v);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && acc != 0.0) atomicAdd(&sc->body_cand, acc);
}

// rank tail entries by (value, index) and by index; build the ascending array for the GPD fit
__global__ void __launch_bounds__(256) psis_tail_rank_kernel(PsisScalars* sc, const double* __restrict__ tail_x,
                                                             const int64_t* __restrict__ tail_i,
                                                             double* __restrict__ sorted_x, int* __restrict__ rank_of,
                                                             int64_t* __restrict__ idx_sorted, int* __restrict__ order_rank) {
  __shared__ double sx[256];
  __shared__ int64_t si[256];
  if (sc->status) return;
  const unsigned int n2 = sc->ntail;
  const double expcut = sc->expcut;
  for (unsigned int i0 = blockIdx.x * blockDim.x; i0 < n2; i0 += gridDim.x * blockDim.x) {
    const unsigned int i = i0 + threadIdx.x;
    const bool valid = i < n2;
    const double xi = valid ? tail_x[i] : 0.0;
    const int64_t ii = valid ? tail_i[i] : 0;
    int r = 0, p = 0;
    for (unsigned int j0 = 0; j0 < n2; j0 += 256) {
      __syncthreads();
      const unsigned int j = j0 + threadIdx.x;
      sx[threadIdx.x] = j < n2 ? tail_x[j] : INFINITY;
      si[threadIdx.x] = j < n2 ? tail_i[j] : INT64_MAX;
      __syncthreads();
      const int lim = (n2 - j0) < 256 ? (int)(n2 - j0) : 256;
      for (int q = 0; q < lim; ++q) {
        const double xj = sx[q];
        const int64_t ij = si[q];
        r += (xj < xi) || (xj == xi && ij < ii);
        p += ij < ii;
      }
    }
    if (valid) {
      rank_of[i] = r;
      sorted_x[r] = exp(xi) - expcut;                  // _psis.py:185-186
      idx_sorted[p] = ii;                              // tail indices in ascending index order
      order_rank[p] = r;                               // rank (0 = smallest tail value) of that index
    }
  }
}

// k_j = mean_i log1p(-b_j x_i) on the quadrature grid (_psis.py:264-286); one CTA per grid point
__global__ void __launch_bounds__(256) psis_gpd_grid_kernel(PsisScalars* sc, const double* __restrict__ sorted_x,
                                                            double* __restrict__ bs, double* __restrict__ ks) {
  __shared__ double red[32];
  if (sc->status) return;
  const int N = (int)sc->ntail;
  if (N <= 4) return;
  const int m = 30 + (int)sqrt((double)N);
  const double xq = sorted_x[(int)(N / 4.0 + 0.5) - 1];
  const double xmax = sorted_x[N - 1];
  for (int j = blockIdx.x; j < m; j += gridDim.x) {
    double b = 1.0 - sqrt((double)m / ((double)(j + 1) - 0.5));
    b /= 3.0 * xq;
    b += 1.0 / xmax;
    const double nb = -b;
    double acc = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) acc += log1p(nb * sorted_x[i]);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
      bs[j] = b;
      ks[j] = acc / (double)N;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double gpinv1(double p, double k, double sigma) {
  // _psis.py:335-377 for 0 < p < 1
  double q;
  if (fabs(k) < DBL_EPSILON)
    q = -log1p(-p);
  else
    q = expm1(-k * log1p(-p)) / k;
  return q * sigma;
}

// posterior weights, k-hat, sigma, smoothing decision, LSE  (_psis.py:288-324, :188-201)
__global__ void __launch_bounds__(kSelThreads) psis_gpd_finish_kernel(PsisScalars* sc, const double* __restrict__ sorted_x,
                                                                       const double* __restrict__ tail_x,
                                                                       const double* __restrict__ bs, const double* __restrict__ ks,
                                                                       double* __restrict__ Ls, double* __restrict__ result) {
  __shared__ double red[32];
  if (sc->status) {
    if (threadIdx.x == 0) result[R_STATUS] = (double)sc->status;
    return;
  }
  const int N = (int)sc->ntail;
  double k = INFINITY, sigma = 0.0;
  if (N > 4) {
    const int m = 30 + (int)sqrt((double)N);
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      double L = bs[j] / ks[j];
      L = log(-L);
      L -= ks[j];
      L -= 1.0;
      Ls[j] = L * (double)N;
    }
    __syncthreads();
    // w_j = 1 / sum_i exp(L_i - L_j); negligible weights dropped; b = sum b_j w_j / sum w_j
    double wsum = 0.0, bsum = 0.0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      const double Lj = Ls[j];
      double den = 0.0;
      for (int i = 0; i < m; ++i) den += exp(Ls[i] - Lj);
      const double w = 1.0 / den;
      if (w >= 10.0 * DBL_EPSILON) {
        wsum += w;
        bsum += bs[j] * w;
      }
    }
    wsum = block_sum(wsum, red);
    bsum = block_sum(bsum, red);
    const double b = bsum / wsum;
    double acc = 0.0;
    const double nb = -b;
    for (int i = threadIdx.x; i < N; i += blockDim.x) acc += log1p(nb * sorted_x[i]);
    acc = block_sum(acc, red);
    k = acc / (double)N;
    sigma = -k / b;
    k = k * (double)N / ((double)N + 10.0) + 5.0 / ((double)N + 10.0);
  }
  const bool smooth = (k >= 1.0 / 3.0) && !isinf(k);
  // LSE pieces: sum over tail of exp(v), v = smoothed (clamped at 0) or raw shifted value
  double ts = 0.0;
  const double expcut = sc->expcut;
  for (int r = threadIdx.x; r < N; r += blockDim.x) {
    double v;
    if (smooth) {
      v = log(gpinv1(((double)r + 0.5) / (double)N, k, sigma) + expcut);
      v = v > 0.0 ? 0.0 : v;
    } else {
      v = tail_x[r];                         // any order: only the sum matters
    }
    ts += exp(v);
  }
  ts = block_sum(ts, red);
  if (threadIdx.x == 0) {
    const double below = sc->body_below > 0.0 ? sc->body_below * exp(sc->t0 - sc->maxv) : 0.0;
    const double total = below + sc->body_cand + ts;
    sc->k = k;
    sc->sigma = sigma;
    sc->smoothed = smooth ? 1 : 0;
    sc->tail_sum = ts;
    sc->lse = log(total);
    result[R_KHAT] = k;
    result[R_SIGMA] = sigma;
    result[R_N2] = (double)N;
    result[R_CUTOFF] = sc->cutoff;
    result[R_LSE] = sc->lse;
    result[R_MAX] = sc->maxv;
    result[R_STATUS] = 0.0;
    result[R_M] = (double)sc->M;
    result[R_NCAND] = (double)sc->ncand;
    result[R_SMOOTHED] = smooth ? 1.0 : 0.0;
  }
}

// pass B: out = (x - max) - lse for the body (tail entries are written by the scatter kernel when
// smoothing is on); moments of v = out + lse for the divergence bounds
__global__ void __launch_bounds__(256) psis_pass_b_kernel(const double* __restrict__ lw, double* __restrict__ out, int64_t n,
                                                          PsisScalars* sc, double* __restrict__ blk_mom) {
  __shared__ double red[32];
  if (sc->status) return;
  const double maxv = sc->maxv, lse = sc->lse, cutoff = sc->cutoff;
  const bool smoothed = sc->smoothed != 0;
  double sv = 0.0, se = 0.0;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  auto one = [&](double x, bool& skip) -> double {
    const double v = x - maxv;
    skip = smoothed && (v > cutoff);
    if (!skip) {
      sv += v;
      se += exp(2.0 * v);
    }
    return v - lse;
  };
  const bool aligned = ((reinterpret_cast<uintptr_t>(lw) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (aligned) {
    const int64_t n2 = n / 2;
    for (int64_t q = tid; q < n2; q += nthreads) {
      const double2 a = __ldcs(reinterpret_cast<const double2*>(lw) + q);
      bool s0, s1;
      double2 o;
      o.x = one(a.x, s0);
      o.y = one(a.y, s1);
      if (!s0 && !s1) {
        __stcs(reinterpret_cast<double2*>(out) + q, o);
      } else {
        if (!s0) out[2 * q] = o.x;
        if (!s1) out[2 * q + 1] = o.y;
      }
    }
    if ((n & 1) && tid == 0) {
      bool s0;
      const double o = one(lw[n - 1], s0);
      if (!s0) out[n - 1] = o;
    }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) {
      bool s0;
      const double o = one(lw[i], s0);
      if (!s0) out[i] = o;
    }
  }
  sv = block_sum(sv, red);
  se = block_sum(se, red);
  if (threadIdx.x == 0) {
    blk_mom[2 * blockIdx.x] = sv;
    blk_mom[2 * blockIdx.x + 1] = se;
  }
}

// smoothed tail values into place (_psis.py:190-199) and final moment reduction
__global__ void __launch_bounds__(256) psis_tail_scatter_kernel(double* __restrict__ out, PsisScalars* sc,
                                                                const int64_t* __restrict__ tail_i,
                                                                const int* __restrict__ rank_of,
                                                                const double* __restrict__ blk_mom, int nblk,
                                                                double* __restrict__ result) {
  __shared__ double red[32];
  if (sc->status) return;
  const int N = (int)sc->ntail;
  const bool smoothed = sc->smoothed != 0;
  const double k = sc->k, sigma = sc->sigma, expcut = sc->expcut, lse = sc->lse;
  double sv = 0.0, se = 0.0;
  if (smoothed) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
      double v = log(gpinv1(((double)rank_of[i] + 0.5) / (double)N, k, sigma) + expcut);
      v = v > 0.0 ? 0.0 : v;
      out[tail_i[i]] = v - lse;
      sv += v;
      se += exp(2.0 * v);
    }
  }
  if (blockIdx.x == 0) {
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
      sv += blk_mom[2 * b];
      se += blk_mom[2 * b + 1];
    }
  }
  sv = block_sum(sv, red);
  se = block_sum(se, red);
  if (threadIdx.x == 0) {
    atomicAdd(&result[R_SUMV], sv);
    atomicAdd(&result[R_SUMEXP2V], se);
  }
}

// ---- exact mode: full-data radix select of the (M+1)-th largest key ----------------------------
__global__ void __launch_bounds__(256) psis_exact_hist_kernel(const double* __restrict__ lw, int64_t n, PsisScalars* sc,
                                                              int shift, int bits, unsigned long long mask,
                                                              unsigned int* __restrict__ ghist) {
  __shared__ unsigned int hist[kBins];
  const unsigned int nb = 1u << bits;
  for (unsigned int b = threadIdx.x; b < nb; b += blockDim.x) hist[b] = 0;
  __syncthreads();
  const unsigned long long prefix = sc->prefix;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (n + nthreads - 1) / nthreads;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t i = it * nthreads + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long key = 0;
    bool ok = false;
    if (i < n) {
      key = dkey(lw[i]);
      ok = (key & mask) == prefix;
    }
    const unsigned int dig = ok ? (unsigned int)((key >> shift) & (nb - 1)) : 0xffffffffu;
    const unsigned int peers = __match_any_sync(0xffffffffu, dig);
    if (ok && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[dig], __popc(peers));
  }
  __syncthreads();
  for (unsigned int b = threadIdx.x; b < nb; b += blockDim.x)
    if (hist[b]) atomicAdd(&ghist[b], hist[b]);
}

__global__ void psis_exact_scan_kernel(PsisScalars* sc, int shift, int bits, unsigned int* __restrict__ ghist, int last) {
  if (threadIdx.x != 0) return;
  const unsigned int nb = 1u << bits;
  unsigned long long K = sc->kth, cum = 0;
  for (int b = (int)nb - 1; b >= 0; --b) {
    const unsigned long long c = ghist[b];
    if (cum + c >= K) {
      sc->prefix |= (unsigned long long)b << shift;
      sc->kth = K - cum;
      break;
    }
    cum += c;
  }
  for (unsigned int b = 0; b < nb; ++b) ghist[b] = 0;
  if (last) {
    sc->t0key = sc->prefix;
    sc->t0 = dkey_inv(sc->prefix);
    sc->strict = 1;
  }
}

__global__ void psis_exact_begin_kernel(PsisScalars* sc) { sc->kth = (unsigned long long)sc->M + 1; sc->prefix = 0; }

// ---- stand-alone divergence-bound moments (diagnostics.py:148-186) --------------------------------
// pass 1: max; pass 2: sum exp(alpha (x - max)), sum x.  out[0]=max, out[1]=sum exp, out[2]=sum x
__global__ void __launch_bounds__(256) dbound_max_kernel(const double* __restrict__ x, int64_t n, unsigned long long* mk) {
  double mx = -INFINITY;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mx = fmax(mx, x[i]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomicMax(mk, dkey(mx));
}
__global__ void __launch_bounds__(256) dbound_sum_kernel(const double* __restrict__ x, int64_t n, double alpha,
                                                         const unsigned long long* mk, double* __restrict__ out) {
  __shared__ double red[32];
  const double mx = dkey_inv(*mk);
  double se = 0.0, sx = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    se += pow(exp(v - mx), alpha);        // np.exp(lw - max) ** alpha
    sx += v;
  }
  se = block_sum(se, red);
  sx = block_sum(sx, red);
  if (threadIdx.x == 0) {
    atomicAdd(&out[1], se);
    atomicAdd(&out[2], sx);
    if (blockIdx.x == 0) out[0] = mx;
  }
}

// ---------------------------------------------------------------------------------------------
struct PsisPlan {
  int M, m_sample, grid, tail_cap, mgrid;
  unsigned int cap, R;
  int64_t stride;
  size_t off_sc, off_skeys, off_candx, off_candi, off_candk, off_blk, off_tailx, off_taili, off_sorted, off_rank,
      off_idxsorted, off_orderrank, off_bs, off_ks, off_Ls, off_mom, off_ghist, total;
};

static void psis_plan(int64_t n, double reff, PsisPlan& p) {
  p.M = (int)ceil(fmin(0.2 * (double)n, 3.0 * sqrt((double)n / reff)));     // _psis.py:158
  if (p.M < 0) p.M = 0;
  p.m_sample = (int)(n < kSampleMax ? n : kSampleMax);
  p.stride = n / p.m_sample;
  const double f = (double)(p.M + 1) * (double)p.m_sample / (double)n;      // expected sample hits above the cutoff
  double R = ceil(f + 8.0 * sqrt(f) + 16.0);
  if (p.m_sample == n) R = p.M + 1;                                          // sample is everything: exact
  p.R = (unsigned int)fmin(R, 4.0e9);
  double expect = (double)p.R * (double)n / (double)p.m_sample;
  double cap = 4.0 * expect + 65536.0;
  if (cap > (double)n) cap = (double)n;
  if (cap < (double)(p.M + 1)) cap = (double)(p.M + 1);
  p.cap = (unsigned int)cap;
  p.tail_cap = p.M + 1;
  p.grid = sm_count() * 8;
  const int64_t need = (n / 4 + 255) / 256;
  if (need < p.grid) p.grid = (int)(need < 1 ? 1 : need);
  p.mgrid = 30 + (int)sqrt((double)p.tail_cap) + 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  p.off_sc = take(sizeof(PsisScalars));
  p.off_skeys = take(sizeof(unsigned long long) * p.m_sample);
  p.off_candx = take(sizeof(double) * p.cap);
  p.off_candi = take(sizeof(int64_t) * p.cap);
  p.off_candk = take(sizeof(unsigned long long) * p.cap);
  p.off_blk = take(sizeof(double) * p.grid);
  p.off_tailx = take(sizeof(double) * p.tail_cap);
  p.off_taili = take(sizeof(int64_t) * p.tail_cap);
  p.off_sorted = take(sizeof(double) * p.tail_cap);
  p.off_rank = take(sizeof(int) * p.tail_cap);
  p.off_idxsorted = take(sizeof(int64_t) * p.tail_cap);
  p.off_orderrank = take(sizeof(int) * p.tail_cap);
  p.off_bs = take(sizeof(double) * p.mgrid);
  p.off_ks = take(sizeof(double) * p.mgrid);
  p.off_Ls = take(sizeof(double) * p.mgrid);
  p.off_mom = take(sizeof(double) * 2 * p.grid);
  p.off_ghist = take(sizeof(unsigned int) * kBins);
  p.total = off;
}

}  // namespace vb
using namespace vb;

extern "C" size_t vb_psis_workspace_bytes(int64_t n, double reff) {
  if (n <= 1 || !(reff > 0)) return 0;
  PsisPlan p;
  psis_plan(n, reff, p);
  return p.total;
}

extern "C" int64_t vb_psis_tail_capacity(int64_t n, double reff) {
  if (n <= 1 || !(reff > 0)) return 0;
  PsisPlan p;
  psis_plan(n, reff, p);
  return p.tail_cap;
}

extern "C" int vb_psislw_f64(const double* lw, double* out, int64_t n, double reff, int exact, double* result,
                             int64_t* tail_idx, int32_t* tail_rank, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
  if (n <= 1) return set_error(VB_ERR_INVALID_ARG, "More than one log-weight needed.");       // _psis.py:144-145
  if (!(reff > 0)) return set_error(VB_ERR_INVALID_ARG, "psislw: Reff must be positive");
  if (!lw || !result) return set_error(VB_ERR_INVALID_ARG, "psislw: null pointer");
  PsisPlan p;
  psis_plan(n, reff, p);
  if (!workspace || workspace_bytes < p.total) return set_error(VB_ERR_WORKSPACE, "psislw: workspace too small");
  char* ws = static_cast<char*>(workspace);
  PsisScalars* sc = reinterpret_cast<PsisScalars*>(ws + p.off_sc);
  auto skeys = reinterpret_cast<unsigned long long*>(ws + p.off_skeys);
  auto candx = reinterpret_cast<double*>(ws + p.off_candx);
  auto candi = reinterpret_cast<int64_t*>(ws + p.off_candi);
  auto candk = reinterpret_cast<unsigned long long*>(ws + p.off_candk);
  auto blk = reinterpret_cast<double*>(ws + p.off_blk);
  auto tailx = reinterpret_cast<double*>(ws + p.off_tailx);
  auto taili = reinterpret_cast<int64_t*>(ws + p.off_taili);
  auto sorted = reinterpret_cast<double*>(ws + p.off_sorted);
  auto rankof = reinterpret_cast<int*>(ws + p.off_rank);
  auto idxsorted = tail_idx ? tail_idx : reinterpret_cast<int64_t*>(ws + p.off_idxsorted);
  auto orderrank = tail_rank ? tail_rank : reinterpret_cast<int*>(ws + p.off_orderrank);
  auto bs = reinterpret_cast<double*>(ws + p.off_bs);
  auto ks = reinterpret_cast<double*>(ws + p.off_ks);
  auto Ls = reinterpret_cast<double*>(ws + p.off_Ls);
  auto mom = reinterpret_cast<double*>(ws + p.off_mom);
  auto ghist = reinterpret_cast<unsigned int*>(ws + p.off_ghist);

  VB_CUDA(cudaMemsetAsync(result, 0, sizeof(double) * R_COUNT, stream));
  psis_init_kernel<<<1, 1, 0, stream>>>(sc, p.M);
  VB_CHECK_LAUNCH();
  if (!exact) {
    psis_sample_kernel<<<(p.m_sample + 255) / 256, 256, 0, stream>>>(lw, n, p.stride, p.m_sample, skeys);
    VB_CHECK_LAUNCH();
    psis_sample_select_kernel<<<1, kSelThreads, 0, stream>>>(skeys, p.m_sample, p.R, sc);
    VB_CHECK_LAUNCH();
  } else {
    VB_CUDA(cudaMemsetAsync(ghist, 0, sizeof(unsigned int) * kBins, stream));
    psis_exact_begin_kernel<<<1, 1, 0, stream>>>(sc);
    VB_CHECK_LAUNCH();
    int shift = 64;
    unsigned long long mask = 0;
    while (shift > 0) {
      const int bits = shift >= kDigitBits ? kDigitBits : shift;
      shift -= bits;
      psis_exact_hist_kernel<<<p.grid, 256, 0, stream>>>(lw, n, sc, shift, bits, mask, ghist);
      VB_CHECK_LAUNCH();
      psis_exact_scan_kernel<<<1, 32, 0, stream>>>(sc, shift, bits, ghist, shift == 0);
      VB_CHECK_LAUNCH();
      mask |= (unsigned long long)((1u << bits) - 1) << shift;
    }
  }
  psis_pass_a_kernel<<<p.grid, 256, 0, stream>>>(lw, n, sc, candx, candi, candk, p.cap, blk);
  VB_CHECK_LAUNCH();
  psis_cutoff_kernel<<<1, kSelThreads, 0, stream>>>(sc, candk, p.cap, blk, p.grid);
  VB_CHECK_LAUNCH();
  psis_tail_compact_kernel<<<sm_count(), 256, 0, stream>>>(sc, candx, candi, tailx, taili, (unsigned)p.tail_cap);
  VB_CHECK_LAUNCH();
  {
    int blocks = (p.tail_cap + 255) / 256;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    psis_tail_rank_kernel<<<blocks, 256, 0, stream>>>(sc, tailx, taili, sorted, rankof, idxsorted, orderrank);
    VB_CHECK_LAUNCH();
  }
  psis_gpd_grid_kernel<<<p.mgrid, 256, 0, stream>>>(sc, sorted, bs, ks);
  VB_CHECK_LAUNCH();
  psis_gpd_finish_kernel<<<1, kSelThreads, 0, stream>>>(sc, sorted, tailx, bs, ks, Ls, result);
  VB_CHECK_LAUNCH();
  if (out) {
    psis_pass_b_kernel<<<p.grid, 256, 0, stream>>>(lw, out, n, sc, mom);
    VB_CHECK_LAUNCH();
    int blocks = (p.tail_cap + 255) / 256;
    if (blocks > sm_count()) blocks = sm_count();
    psis_tail_scatter_kernel<<<blocks, 256, 0, stream>>>(out, sc, taili, rankof, mom, p.grid, result);
    VB_CHECK_LAUNCH();
  }
  return VB_OK;
}

extern "C" int vb_divergence_moments_f64(const double* lw, int64_t n, double alpha, double* out3, cudaStream_t stream) {
  if (n <= 0 || !lw || !out3) return set_error(VB_ERR_INVALID_ARG, "divergence_moments: bad arguments");
  if (!(alpha > 1.0)) return set_error(VB_ERR_INVALID_ARG, "alpha must be greater than 1");      // diagnostics.py:172-173
  // out3[3] doubles as scratch for the max key
  VB_CUDA(cudaMemsetAsync(out3, 0, sizeof(double) * 4, stream));
  int grid = sm_count() * 8;
  const int64_t need = (n + 255) / 256;
  if (need < grid) grid = (int)need;
  dbound_max_kernel<<<grid, 256, 0, stream>>>(lw, n, reinterpret_cast<unsigned long long*>(out3 + 3));
  VB_CHECK_LAUNCH();
  dbound_sum_kernel<<<grid, 256, 0, stream>>>(lw, n, alpha, reinterpret_cast<unsigned long long*>(out3 + 3), out3);
  VB_CHECK_LAUNCH();
  return VB_OK;
}
