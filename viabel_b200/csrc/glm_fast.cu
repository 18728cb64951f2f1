// Tensor-core "fast path" of the GLM model plugin (logistic link), sm_100a only:
// TMA -> shared memory -> tcgen05.mma (fp16 operands, fp32 accumulators in TMEM).
//
// Per 128-observation tile, one persistent CTA per SM:
//   GEMM1  Z[n,s]  = sum_j Xy[n,j] Theta[s,j]        M=128 (n), N=256 (s), K=d
//          three fp16 passes  Xh.Th + Xl.Th + Xh.Tl  (Xy = y*X and Theta split hi+lo: 22-bit operands)
//   E1     link epilogue out of TMEM: ll[s] += -softplus(-z), R[n,s] = w_s*sigmoid(-z) -> fp16 tile in
//          shared memory (the N x S logits / residuals never touch HBM), rbar[n] = sum_s R[n,s]
//   GEMM2  Tt[j,n] = sum_s E[s,j] R[n,s]             M=128 (j), N=128 (n), K=256 (s); A = E (MN-major)
//   E2     thread owns j: ge[j] += sum_n Xy[n,j] Tt[j,n],  gmu[j] += sum_n Xy[n,j] rbar[n]
//          (Xy chunk re-read through TMA, an L2 hit)
// Replaces the same reference code as glm_f64.cu (user log_density under autograd:
// models.py:27-39, objectives.py:161-167).  Tolerance of this path: 1e-4 relative (BASELINE.json).
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 = epilogue.
// Shared memory: 5 x 32 KB operand ring, 64 KB R tile, barriers.  TMEM: Z 256 cols, Tt 2 x 128 cols.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace vb {
namespace fast {

constexpr int kBM = 128;
constexpr int kSP = 256;                 // padded sample count (UMMA N of GEMM1, K of GEMM2)
constexpr int kSlotBytes = 32768;
constexpr int kNumSlots = 5;
constexpr int kRingBytes = kSlotBytes * kNumSlots;      // 163840
constexpr int kRBytes = kBM * kSP * 2;                  // 65536
constexpr int kMiscBytes = 3072;
constexpr int kSmemBytes = kRingBytes + kRBytes + kMiscBytes;   // 232448 = 227 KB
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;           // 320
constexpr uint32_t kTmemCols = 512;

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 columns of 32-bit accumulators -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors -------------------------------------------------------------------------------
// shared-memory matrix descriptor, 128-byte swizzle (layout_type = 2), descriptor version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // version
  d |= (uint64_t)2 << 61;     // SWIZZLE_128B
  return d;
}
// instruction descriptor: f16 x f16 -> f32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) /*c=f32*/ | (0u << 7) /*a=f16*/ | (0u << 10) /*b=f16*/ | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
  int64_t N;
  int d_pad;       // multiple of 128
  int S;           // valid samples (<= 256)
  int numTiles;
  int want_grad;
  const float* w;  // [256] sample weights, 0 beyond S
  double* ll_part;   // [grid][256]
  double* gmu_part;  // [grid][d_pad]
  double* ge_part;   // [grid][d_pad]
  float* dbg;        // optional: Z of this CTA's first tile [128][256], then Tt of j-block 0 [128][128]
};

struct Misc {
  uint64_t full[kNumSlots], empty[kNumSlots];
  uint64_t z_full, r_full, t_full[2], t_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float w[kSP];
  float rbar[2][kBM];
};
static_assert(sizeof(Misc) <= kMiscBytes, "misc smem overflow");

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
glm_fast_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                const __grid_constant__ CUtensorMap tmTh, const __grid_constant__ CUtensorMap tmTl,
                const __grid_constant__ CUtensorMap tmE, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint8_t* rtile = smem + kRingBytes;
  Misc* misc = reinterpret_cast<Misc*>(smem + kRingBytes + kRBytes);
  const uint32_t ring_u = smem_u32(ring), rtile_u = smem_u32(rtile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = p.d_pad / 64, JB = p.d_pad / 128;
  const int fillsPerTile = 3 * KC + (p.want_grad ? 4 * JB : 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumSlots; ++i) {
      mbar_init(smem_u32(&misc->full[i]), 1);
      mbar_init(smem_u32(&misc->empty[i]), 1);
    }
    mbar_init(smem_u32(&misc->z_full), 1);
    mbar_init(smem_u32(&misc->r_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&misc->t_full[i]), 1);
      mbar_init(smem_u32(&misc->t_empty[i]), 1);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kSP; i += kThreads) misc->w[i] = p.w[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_base;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      prefetch_tmap(&tmXh); prefetch_tmap(&tmXl); prefetch_tmap(&tmTh); prefetch_tmap(&tmTl); prefetch_tmap(&tmE);
      uint32_t fill = 0;
      auto acquire = [&](uint32_t bytes, uint32_t& dst, uint32_t& bar) {
        const uint32_t slot = fill % kNumSlots, par = (fill / kNumSlots) & 1;
        mbar_wait(smem_u32(&misc->empty[slot]), par ^ 1);
        bar = smem_u32(&misc->full[slot]);
        mbar_expect_tx(bar, bytes);
        dst = ring_u + slot * kSlotBytes;
        ++fill;
      };
      for (int tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x) {
        const int n0 = tile * kBM;
        uint32_t dst, bar;
        for (int kc = 0; kc < KC; ++kc) {
          acquire(32768, dst, bar);
          tma_load_2d(dst, &tmXh, kc * 64, n0, bar);
          tma_load_2d(dst + 16384, &tmXl, kc * 64, n0, bar);
          acquire(32768, dst, bar);
          tma_load_2d(dst, &tmTh, kc * 64, 0, bar);
          tma_load_2d(dst + 16384, &tmTh, kc * 64, 128, bar);
          acquire(32768, dst, bar);
          tma_load_2d(dst, &tmTl, kc * 64, 0, bar);
          tma_load_2d(dst + 16384, &tmTl, kc * 64, 128, bar);
        }
        if (p.want_grad) {
          for (int jb = 0; jb < JB; ++jb) {
            const int j0 = jb * 128;
            for (int half = 0; half < 2; ++half) {      // E for K-groups (2*half, 2*half+1)
              acquire(32768, dst, bar);
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const int s0 = (2 * half + g) * 64;
                tma_load_2d(dst + g * 16384, &tmE, j0, s0, bar);
                tma_load_2d(dst + g * 16384 + 8192, &tmE, j0 + 64, s0, bar);
              }
            }
            acquire(32768, dst, bar);
            tma_load_2d(dst, &tmXh, j0, n0, bar);
            tma_load_2d(dst + 16384, &tmXh, j0 + 64, n0, bar);
            acquire(32768, dst, bar);
            tma_load_2d(dst, &tmXl, j0, n0, bar);
            tma_load_2d(dst + 16384, &tmXl, j0 + 64, n0, bar);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(128, 256, 0, 0);   // Z: A = X (K-major), B = Theta (K-major)
      constexpr uint32_t idesc2 = make_idesc(128, 128, 1, 0);   // Tt: A = E (MN-major), B = R (K-major)
      uint32_t ltile = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x, ++ltile) {
        const uint32_t fbase = ltile * fillsPerTile;
        // ---- GEMM1 ----
        for (int kc = 0; kc < KC; ++kc) {
          uint32_t sa[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const uint32_t f = fbase + 3 * kc + i, slot = f % kNumSlots, par = (f / kNumSlots) & 1;
            mbar_wait(smem_u32(&misc->full[slot]), par);
            sa[i] = ring_u + slot * kSlotBytes;
          }
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t xh = make_desc(sa[0] + k * 32, 16, 1024);
            const uint64_t xl = make_desc(sa[0] + 16384 + k * 32, 16, 1024);
            const uint64_t th = make_desc(sa[1] + k * 32, 16, 1024);
            const uint64_t tl = make_desc(sa[2] + k * 32, 16, 1024);
            umma_f16(tmem, xh, th, idesc1, (kc | k) ? 1u : 0u);
            umma_f16(tmem, xl, th, idesc1, 1u);
            umma_f16(tmem, xh, tl, idesc1, 1u);
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const uint32_t f = fbase + 3 * kc + i;
            umma_commit(smem_u32(&misc->empty[f % kNumSlots]));
          }
        }
        umma_commit(smem_u32(&misc->z_full));
        // ---- GEMM2 ----  (r_full also means "Z has been read": the next tile may overwrite it)
        mbar_wait(smem_u32(&misc->r_full), ltile & 1);
        tc_fence_after();
        if (p.want_grad) {
          for (int jb = 0; jb < JB; ++jb, ++tcount) {
            const uint32_t buf = tcount & 1;
            mbar_wait(smem_u32(&misc->t_empty[buf]), ((tcount >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t dT = tmem + 256 + 128 * buf;
            for (int half = 0; half < 2; ++half) {
              const uint32_t f = fbase + 3 * KC + 4 * jb + half, slot = f % kNumSlots, par = (f / kNumSlots) & 1;
              mbar_wait(smem_u32(&misc->full[slot]), par);
              tc_fence_after();
              const uint32_t es = ring_u + slot * kSlotBytes;
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const int kg = 2 * half + g;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  // A = E^T tile: 64 j per 128-byte row, 16 s rows per K step, second 64-j group at +8 KB
                  const uint64_t ea = make_desc(es + g * 16384 + k * 2048, 8192, 1024);
                  const uint64_t rb = make_desc(rtile_u + kg * 16384 + k * 32, 16, 1024);
                  umma_f16(dT, ea, rb, idesc2, (kg | k) ? 1u : 0u);
                }
              }
              umma_commit(smem_u32(&misc->empty[slot]));
            }
            umma_commit(smem_u32(&misc->t_full[buf]));
          }
        }
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int ew = warp - 2;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int h = ew >> 2;             // column half
    const int row = 32 * q + lane;     // TMEM lane = tile row (E1) or j within block (E2)
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    double ll_acc[4] = {0.0, 0.0, 0.0, 0.0};
    double ge_acc[16], gmu_acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ge_acc[i] = gmu_acc[i] = 0.0;
    uint32_t ltile = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < p.numTiles; tile += gridDim.x, ++ltile) {
      const uint32_t fbase = ltile * fillsPerTile;
      const int64_t n = (int64_t)tile * kBM + row;
      const float rowvalid = n < p.N ? 1.0f : 0.0f;
      // -------- E1: link epilogue on Z --------
      mbar_wait(smem_u32(&misc->z_full), ltile & 1);
      tc_fence_after();
      float rsum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = 128 * h + 32 * c;
        float v[32];
        tmem_ld32(lane_addr + col0, v);
        if (p.dbg && ltile == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) p.dbg[(size_t)blockIdx.x * 49152 + row * 256 + col0 + i] = v[i];
        }
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float r2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float a = v[i + u];
            const float t = fast_exp2(-fabsf(a) * 1.4426950408889634f);
            const float den = 1.0f + t;
            const float l1p = __log2f(den) * 0.6931471805599453f;
            v[i + u] = (fmaxf(-a, 0.0f) + l1p) * rowvalid;                 // softplus(-a)
            const float sg = __fdividef(a >= 0.0f ? t : 1.0f, den);        // sigmoid(-a)
            r2[u] = sg * misc->w[col0 + i + u] * rowvalid;
            rsum += r2[u];
          }
          const __half2 hh = __floats2half2_rn(r2[0], r2[1]);
          packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        // R tile: K-group g (64 samples, 16 KB), row-major 128-byte rows, 128B swizzle
        {
          const int g = col0 >> 6;
          const int cbase = (col0 & 63) >> 3;          // first 16-byte chunk of this 32-sample run
          uint8_t* rowp = rtile + g * 16384 + row * 128;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int chunk = (cbase + u) ^ (row & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) =
                make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
          }
        }
        // column sums of softplus over this warp's 32 rows (transpose-reduce): lane l ends with column l
#pragma unroll
        for (int o = 16, cnt = 16; o >= 1; o >>= 1, cnt >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < cnt; ++i) {
            const float send = up ? v[i] : v[i + cnt];
            const float keep = up ? v[i + cnt] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        ll_acc[c] += (double)v[0];
      }
      misc->rbar[h][row] = rsum;
      fence_proxy_async();       // R tile writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      epi_barrier();
      if (threadIdx.x == 64) mbar_arrive(smem_u32(&misc->r_full));
      if (!p.want_grad) continue;
      // -------- E2: thread owns column j = 128*jb + row --------
#pragma unroll 1
      for (int jb = 0; jb < JB; ++jb, ++tcount) {
        const uint32_t buf = tcount & 1;
        const uint32_t fh = fbase + 3 * KC + 4 * jb + 2, fl = fh + 1;
        const uint32_t sh = fh % kNumSlots, sl = fl % kNumSlots;
        mbar_wait(smem_u32(&misc->t_full[buf]), (tcount >> 1) & 1);
        mbar_wait(smem_u32(&misc->full[sh]), (fh / kNumSlots) & 1);
        mbar_wait(smem_u32(&misc->full[sl]), (fl / kNumSlots) & 1);
        tc_fence_after();
        const uint8_t* xh = ring + sh * kSlotBytes + (row >> 6) * 16384;
        const uint8_t* xl = ring + sl * kSlotBytes + (row >> 6) * 16384;
        const int cj = row & 63;
        const int cchunk = (cj * 2) >> 4, cbyte = (cj * 2) & 15;
        float ge = 0.0f, gm = 0.0f;
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
          const int nb = 64 * h + 32 * part;
          float tv[32];
          tmem_ld32(lane_addr + 256 + 128 * buf + nb, tv);
          if (p.dbg && ltile == 0 && jb == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.dbg[(size_t)blockIdx.x * 49152 + 32768 + row * 128 + nb + i] = tv[i];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int nn = nb + i;
            const int off = nn * 128 + ((cchunk ^ (nn & 7)) << 4) + cbyte;
            const float x = __half2float(*reinterpret_cast<const __half*>(xh + off)) +
                            __half2float(*reinterpret_cast<const __half*>(xl + off));
            ge = fmaf(x, tv[i], ge);
            gm = fmaf(x, misc->rbar[0][nn] + misc->rbar[1][nn], gm);
          }
        }
        ge_acc[jb & 15] += (double)ge;
        gmu_acc[jb & 15] += (double)gm;
        tc_fence_before();
        epi_barrier();
        if (threadIdx.x == 64) {
          mbar_arrive(smem_u32(&misc->t_empty[buf]));
          mbar_arrive(smem_u32(&misc->empty[sh]));
          mbar_arrive(smem_u32(&misc->empty[sl]));
        }
      }
    }
    // -------- per-CTA partial sums (the operand ring is idle by now: use it as scratch) --------
    double (*red)[kBM] = reinterpret_cast<double (*)[kBM]>(ring);
    // ll: lane l of (q,h) holds column 128h + 32c + l summed over rows of quarter q -> sum the 4 quarters
    for (int c = 0; c < 4; ++c) {
      epi_barrier();
      if (q != 0) red[h][(q - 1) * 32 + lane] = ll_acc[c];       // 3 x 32 slots per column half
      epi_barrier();
      if (q == 0) {
        const double s = ll_acc[c] + red[h][lane] + red[h][32 + lane] + red[h][64 + lane];
        p.ll_part[(size_t)blockIdx.x * kSP + 128 * h + 32 * c + lane] = s;
      }
    }
    if (p.want_grad) {
      for (int jb = 0; jb < JB; ++jb) {
        epi_barrier();
        if (h == 1) {
          red[0][row] = ge_acc[jb & 15];
          red[1][row] = gmu_acc[jb & 15];
        }
        epi_barrier();
        if (h == 0) {
          p.ge_part[(size_t)blockIdx.x * p.d_pad + jb * 128 + row] = ge_acc[jb & 15] + red[0][row];
          p.gmu_part[(size_t)blockIdx.x * p.d_pad + jb * 128 + row] = gmu_acc[jb & 15] + red[1][row];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

// ---- operand preparation ------------------------------------------------------------------------
// Xy = y*X split into fp16 hi + lo, zero padded to [N_pad][d_pad]
__global__ void fast_prepare_x_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ y, int64_t N,
                                      int d, int64_t N_pad, int d_pad, __half* __restrict__ Xh, __half* __restrict__ Xl,
                                      float* __restrict__ absmax) {
  const int64_t total = N_pad * d_pad;
  float mx = 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / d_pad;
    const int j = (int)(i - n * d_pad);
    float v = 0.0f;
    if (n < N && j < d) v = (float)(X[n * ldx + j] * y[n]);
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    Xh[i] = hi;
    Xl[i] = lo;
    mx = fmaxf(mx, fabsf(v));
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(absmax), __float_as_int(mx));
}

// Theta hi/lo, E (fp16) and weights, zero padded to [256][d_pad]
__global__ void fast_prepare_theta_kernel(const double* __restrict__ theta, const double* __restrict__ base,
                                          const double* __restrict__ w, int64_t S, int d, int d_pad,
                                          __half* __restrict__ Th, __half* __restrict__ Tl, __half* __restrict__ E,
                                          float* __restrict__ wf) {
  const int64_t total = (int64_t)kSP * d_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / d_pad;
    const int j = (int)(i - s * d_pad);
    float t = 0.0f, e = 0.0f;
    if (s < S && j < d) {
      t = (float)theta[s * d + j];
      if (base) e = (float)base[s * d + j];
    }
    const __half hi = __float2half_rn(t);
    Th[i] = hi;
    Tl[i] = __float2half_rn(t - __half2float(hi));
    E[i] = __float2half_rn(e);
  }
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < kSP; s += (int64_t)gridDim.x * blockDim.x)
    wf[s] = s < S ? (w ? (float)w[s] : 1.0f) : 0.0f;
}

__global__ void reduce_partials_kernel(const double* __restrict__ part, int nblk, int64_t stride, int64_t n,
                                       double sign, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(size_t)b * stride + i];
    out[i] = sign * s;
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp16 row-major [rows][cols] tensor, box = [box_rows][64 cols], 128-byte swizzle
static bool encode_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct FastModel {
  int64_t N, N_pad;
  int d, d_pad;
  const __half* Xh;
  const __half* Xl;
  CUtensorMap tmXh, tmXl;
};

struct FastLayout {
  size_t off_Th, off_Tl, off_E, off_w, off_ll, off_gmu, off_ge, off_dbg, total;
  int grid;
};

static void fast_layout(int64_t N, int d_pad, FastLayout& L) {
  const int64_t tiles = ceil_div(N, kBM);
  int sms = sm_count();
  L.grid = (int)(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  L.off_Th = take((size_t)kSP * d_pad * 2);
  L.off_Tl = take((size_t)kSP * d_pad * 2);
  L.off_E = take((size_t)kSP * d_pad * 2);
  L.off_w = take(kSP * sizeof(float));
  L.off_ll = take((size_t)L.grid * kSP * sizeof(double));
  L.off_gmu = take((size_t)L.grid * d_pad * sizeof(double));
  L.off_ge = take((size_t)L.grid * d_pad * sizeof(double));
  L.total = off;
}

}  // namespace fast
}  // namespace vb

using namespace vb;
using namespace vb::fast;

extern "C" size_t vb_glm_fast_model_bytes(int64_t N, int d) {
  if (N <= 0 || d <= 0) return 0;
  const int64_t N_pad = ceil_div(N, kBM) * kBM;
  const int64_t d_pad = ceil_div(d, 128) * 128;
  return (size_t)(2 * N_pad * d_pad * 2) + 1024;
}

extern "C" size_t vb_glm_fast_workspace_bytes(int64_t N, int d, int64_t S) {
  if (N <= 0 || d <= 0 || S <= 0 || S > kSP) return 0;
  FastLayout L;
  fast_layout(N, (int)(ceil_div(d, 128) * 128), L);
  return L.total;
}

extern "C" int vb_glm_fast_create(void** handle, const double* X, int64_t ldx, const double* y, int64_t N, int d,
                                  int link, void* model_mem, size_t model_bytes, float* absmax_host,
                                  cudaStream_t stream) {
  if (!handle || !X || !y || N <= 0 || d <= 0 || ldx < d || !model_mem)
    return set_error(VB_ERR_INVALID_ARG, "glm_fast_create: bad arguments");
  if (link != VB_LINK_LOGISTIC) return set_error(VB_ERR_UNSUPPORTED, "glm_fast: only the logistic link has a tensor-core path");
  if (d > 2048) return set_error(VB_ERR_UNSUPPORTED, "glm_fast: d > 2048 not supported");
  if (model_bytes < vb_glm_fast_model_bytes(N, d)) return set_error(VB_ERR_WORKSPACE, "glm_fast_create: model buffer too small");
  if ((reinterpret_cast<uintptr_t>(model_mem) & 1023) != 0)
    return set_error(VB_ERR_INVALID_ARG, "glm_fast_create: model buffer must be 1024-byte aligned");
  FastModel* m = new FastModel();
  m->N = N;
  m->d = d;
  m->N_pad = ceil_div(N, kBM) * kBM;
  m->d_pad = (int)(ceil_div(d, 128) * 128);
  __half* Xh = static_cast<__half*>(model_mem);
  __half* Xl = Xh + m->N_pad * m->d_pad;
  float* absmax = reinterpret_cast<float*>(Xl + m->N_pad * m->d_pad);
  m->Xh = Xh;
  m->Xl = Xl;
  cudaError_t e = cudaMemsetAsync(absmax, 0, sizeof(float), stream);
  if (e != cudaSuccess) { delete m; return set_cuda_error(e); }
  fast_prepare_x_kernel<<<sm_count() * 8, 256, 0, stream>>>(X, ldx, y, N, d, m->N_pad, m->d_pad, Xh, Xl, absmax);
  e = cudaGetLastError();
  if (e != cudaSuccess) { delete m; return set_cuda_error(e); }
  if (absmax_host) {
    e = cudaMemcpyAsync(absmax_host, absmax, sizeof(float), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { delete m; return set_cuda_error(e); }
    if (!(*absmax_host < 3.0e4f)) {
      delete m;
      return set_error(VB_ERR_UNSUPPORTED, "glm_fast: |y*X| exceeds the fp16 operand range; use the float64 path");
    }
  }
  if (!encode_2d(&m->tmXh, Xh, m->N_pad, m->d_pad, 128) || !encode_2d(&m->tmXl, Xl, m->N_pad, m->d_pad, 128)) {
    delete m;
    return set_error(VB_ERR_CUDA, "glm_fast_create: cuTensorMapEncodeTiled failed");
  }
  *handle = m;
  return VB_OK;
}

extern "C" int vb_glm_fast_destroy(void* handle) {
  delete static_cast<FastModel*>(handle);
  return VB_OK;
}

extern "C" int vb_glm_fast_sweep(void* handle, const double* theta, const double* base, const double* w, int64_t S,
                                 int want_grad, double* out_ll, double* out_gmu, double* out_ge, void* workspace,
                                 size_t workspace_bytes, float* debug, cudaStream_t stream) {
  FastModel* m = static_cast<FastModel*>(handle);
  if (!m || !theta || !out_ll || S <= 0) return set_error(VB_ERR_INVALID_ARG, "glm_fast_sweep: bad arguments");
  if (S > kSP) return set_error(VB_ERR_UNSUPPORTED, "glm_fast_sweep: at most 256 samples per sweep");
  if (want_grad && (!base || !out_gmu || !out_ge)) return set_error(VB_ERR_INVALID_ARG, "glm_fast_sweep: want_grad needs base, out_gmu, out_ge");
  FastLayout L;
  fast_layout(m->N, m->d_pad, L);
  if (!workspace || workspace_bytes < L.total) return set_error(VB_ERR_WORKSPACE, "glm_fast_sweep: workspace too small");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0)
    return set_error(VB_ERR_INVALID_ARG, "glm_fast_sweep: workspace must be 1024-byte aligned");
  char* ws = static_cast<char*>(workspace);
  __half* Th = reinterpret_cast<__half*>(ws + L.off_Th);
  __half* Tl = reinterpret_cast<__half*>(ws + L.off_Tl);
  __half* E = reinterpret_cast<__half*>(ws + L.off_E);
  float* wf = reinterpret_cast<float*>(ws + L.off_w);

  static thread_local const void* cached_ws = nullptr;
  static thread_local int cached_dpad = 0;
  static thread_local CUtensorMap tmTh, tmTl, tmE;
  if (cached_ws != workspace || cached_dpad != m->d_pad) {
    if (!encode_2d(&tmTh, Th, kSP, m->d_pad, 128) || !encode_2d(&tmTl, Tl, kSP, m->d_pad, 128) ||
        !encode_2d(&tmE, E, kSP, m->d_pad, 64))
      return set_error(VB_ERR_CUDA, "glm_fast_sweep: cuTensorMapEncodeTiled failed");
    cached_ws = workspace;
    cached_dpad = m->d_pad;
  }

  fast_prepare_theta_kernel<<<128, 256, 0, stream>>>(theta, want_grad ? base : nullptr, w, S, m->d, m->d_pad, Th, Tl, E, wf);
  VB_CHECK_LAUNCH();

  Params p;
  p.N = m->N;
  p.d_pad = m->d_pad;
  p.S = (int)S;
  p.numTiles = (int)ceil_div(m->N, kBM);
  p.want_grad = want_grad;
  p.w = wf;
  p.ll_part = reinterpret_cast<double*>(ws + L.off_ll);
  p.gmu_part = reinterpret_cast<double*>(ws + L.off_gmu);
  p.ge_part = reinterpret_cast<double*>(ws + L.off_ge);
  p.dbg = debug;
  static bool attr_set = false;
  if (!attr_set) {
    VB_CUDA(cudaFuncSetAttribute(glm_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  glm_fast_kernel<<<L.grid, kThreads, kSmemBytes, stream>>>(m->tmXh, m->tmXl, tmTh, tmTl, tmE, p);
  VB_CHECK_LAUNCH();
  // ll = -sum softplus
  reduce_partials_kernel<<<1, 256, 0, stream>>>(p.ll_part, L.grid, kSP, S, -1.0, out_ll);
  VB_CHECK_LAUNCH();
  if (want_grad) {
    reduce_partials_kernel<<<(m->d + 255) / 256, 256, 0, stream>>>(p.gmu_part, L.grid, m->d_pad, m->d, 1.0, out_gmu);
    VB_CHECK_LAUNCH();
    reduce_partials_kernel<<<(m->d + 255) / 256, 256, 0, stream>>>(p.ge_part, L.grid, m->d_pad, m->d, 1.0, out_ge);
    VB_CHECK_LAUNCH();
  }
  return VB_OK;
}
