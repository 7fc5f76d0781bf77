// lpmx_pair_kernel.cuh -- device code of the O(N^2) pair-sum kernel family (sm_100a).
// Included by lpmx_kernels.cu (the product instances) and tools/tune_pair_sum.cu (the tuning
// harness that times alternative shapes on the GPU box).  See lpmx_kernels.cu for the design notes.
#ifndef LPMX_PAIR_KERNEL_CUH
#define LPMX_PAIR_KERNEL_CUH

#include "lpmx_fast_log.h"
#include "lpmx_internal.h"

namespace lpmx {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait for a phase of a ring barrier.  Steady state: the first try_wait succeeds or suspends for the hardware's time limit.
// A wait that keeps failing backs off (nanosleep after 1024 tries) and has a deadline: a bulk copy that faulted or a ring
// protocol error would otherwise spin forever and take the GPU with it; after kMbarDeadlineNs the CTA traps, the launch
// fails with a sticky CUDA error and the next lpmx_* call reports it.  A legitimate wait is microseconds (one 256-record
// chunk of another warp's work).
constexpr unsigned long long kMbarDeadlineNs = 4000000000ull;
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 1023u) == 0) {
      __nanosleep(256);
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > kMbarDeadlineNs) asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
// global -> shared bulk copy (TMA, SASS UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 20-bit reciprocal seed (MUFU.RCP64H)
__device__ __forceinline__ double rcp_seed(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
}
// ------------------------------------------------------------------------------------------------
// log(d) for the stream-function kinds: lpmx_fast_log.h (a 256-entry shared-memory table behind the source ring and a
// degree-5 polynomial, 9 FP64-pipe instructions).  The table {c_i, -log c_i} is copied from global memory by every CTA.
// ------------------------------------------------------------------------------------------------
__device__ const double2 kLogTable[kLogMEntries] = {
#if LPMX_LOG_MBITS == 10
#include "log_table_10.inc"
#elif LPMX_LOG_MBITS == 8
#include "log_table_8.inc"
#else
#include "log_table_7.inc"
#endif
};

// kinds whose chunk sums are formed from zero and added to the running total once per chunk (see pair_sum_kernel)
__host__ __device__ constexpr bool kind_two_level(int k) { return k == kVel || k == kVelPsi || k == kPsi; }
__host__ __device__ constexpr bool kind_has_log(int k) { return k == kVelPsi || k == kPsi || k == kPlaneVelPsi || k == kPlaneSwe; }

// the two tables of one CTA: mtab (16-byte aligned) then ktab
struct LogTables {
  const double2* m;
  const double* k;
};
__host__ __device__ constexpr size_t log_tables_bytes() {
  return kLogMEntries * sizeof(double2) + ((kLogKEntries * sizeof(double) + 15) / 16) * 16;
}
__device__ __forceinline__ double fast_log(double d, const LogTables& t) { return fast_log(d, t.m, t.k); }

// exp(-x) for x >= 0 (the PSE kernel, lpm_pse.hpp:66-73): k = rint(-x log2 e) by the magic-number add, r = -x - k ln2
// in two FMAs (hi/lo split of ln2), |r| <= ln2/2, degree-12 Taylor polynomial (|r|^13/13! < 2e-16), scale by 2^k
// with an integer add on the exponent field.  x is clamped to 700 (exp(-700) ~ 1e-304 stays normal; the PSE term
// it multiplies is below any tolerance there).  17 FP64-pipe instructions; relative error < 4e-16.
__device__ __forceinline__ double fast_exp_neg(double x) {
  x = fmin(x, 700.0);
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, -1.4426950408889634074, kMagic);
  const int k = __double2loint(t);
  const double kf = t - kMagic;
  double r = fma(kf, -6.93147180369123816490e-01, -x);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double p = fma(r, 1.0 / 479001600.0, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// ------------------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------------------
struct SumArgs {
  Vec3View tgt;
  const int* tgt_map;  // optional: target tg of this launch is element tgt_map[tg] of `tgt` / `self_idx` (an index list of a
                       // sharded solver: its leaf faces, or its vertices and divided faces); null: element tg itself
  const int* self_idx;
  const double* packed;
  double* part;
  int n_tgt;
  int n_tb;
  int n_sc;
  long n_tgt_pad;
  double kappa;  // 1 + eps^2 (sphere) / eps^2 (plane)
  double aux;    // 1 / pse_eps^2 (kPlaneSwe)
};

// dynamic shared memory of one CTA: source ring + full/empty barriers (+ the log table) (+ the running totals of the two-level
// summation: T * NACC doubles per compute thread, see pair_sum_kernel)
__host__ __device__ constexpr size_t pair_smem_bytes(int kind, int T, int lanes) {
  return (size_t)kStages * kChunk * kind_rec(kind) * sizeof(double) + 2 * kStages * sizeof(uint64_t) +
         (kind_has_log(kind) ? log_tables_bytes() : 0) +
         (kind_two_level(kind) ? (size_t)T * kind_nacc(kind) * lanes * sizeof(double) : 0);
}

__host__ __device__ __forceinline__ int cta_of_item(long item, int grid, long n_items) {
  return (int)(((item + 1) * (long)grid - 1) / n_items);
}

// Per-kind pair bodies.  x = target, (y, g..) = source record, acc = this target's accumulators.
// CHECK: compare the source's global compact index with the target's own (self) index.
template <int KIND, bool CHECK>
struct Pair;

// BVE / IC2D record: s = {y0, y1, y2, G*y0, G*y1, G*y2, G, 0}.
// Velocity moment: 9 FP64-pipe instructions per pair -- 3 (d), 3 (r = 1/d from the MUFU seed), 3 (M += r * G*y).
template <bool CHECK>
struct Pair<kVel, CHECK> {
  static constexpr int NLOAD = 6;  // doubles of the record this kind reads
  __device__ __forceinline__ static void apply(const double* x, const double* /*kx*/, double kappa, double /*aux*/, const double* s,
                                               int j, int self, double* acc, const LogTables& /*tbl*/) {
    const double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    const double r0 = rcp_seed(d);
    const double e = fma(-d, r0, 1.0);
    const double p = fma(e, e, e);
    double r = fma(r0, p, r0);
    if (CHECK) r = (j == self) ? 0.0 : r;
    acc[0] = fma(r, s[3], acc[0]);
    acc[1] = fma(r, s[4], acc[1]);
    acc[2] = fma(r, s[5], acc[2]);
  }
};

template <bool CHECK>
struct Pair<kVelPsi, CHECK> {
  static constexpr int NLOAD = 8;
  __device__ __forceinline__ static void apply(const double* x, const double*, double kappa, double, const double* s, int j,
                                               int self, double* acc, const LogTables& tbl) {
    double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    double gam = s[6];
    if (CHECK) {
      const bool me = (j == self);
      d = me ? 1.0 : d;
      gam = me ? 0.0 : gam;
    }
    const double r0 = rcp_seed(d);
    const double e = fma(-d, r0, 1.0);
    const double p = fma(e, e, e);
    double r = fma(r0, p, r0);
    if (CHECK) r = (j == self) ? 0.0 : r;
    acc[0] = fma(r, s[3], acc[0]);
    acc[1] = fma(r, s[4], acc[1]);
    acc[2] = fma(r, s[5], acc[2]);
    acc[3] = fma(gam, fast_log(d, tbl), acc[3]);
  }
};

template <bool CHECK>
struct Pair<kPsi, CHECK> {
  static constexpr int NLOAD = 8;
  __device__ __forceinline__ static void apply(const double* x, const double*, double kappa, double, const double* s, int j,
                                               int self, double* acc, const LogTables& tbl) {
    double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    double gam = s[6];
    if (CHECK) {
      const bool me = (j == self);
      d = me ? 1.0 : d;
      gam = me ? 0.0 : gam;
    }
    acc[0] = fma(gam, fast_log(d, tbl), acc[0]);
  }
};

// kSwe accumulators: [0..2] Mz = sum Gz y/d, [3..5] Ms = sum Gs y/d, [6..14] G (row-major):
//   G_ab += (Gz/d^2) c_a q_b - (Gs/d^2) q_a p_b,  c = x cross y, q = kappa x - y, p = y - (x.y) x
// The 1/d parts of the coded gradient polynomials ([y]x / d and (x.y) P / d) are linear in y and
// are rebuilt from Mz and Ms in the finalize kernel.
template <bool CHECK>
struct Pair<kSwe, CHECK> {
  static constexpr int NLOAD = 6;
  __device__ __forceinline__ static void apply(const double* x, const double* kx, double kappa, double, const double* s,
                                               int j, int self, double* acc, const LogTables& /*tbl*/) {
    double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    double gz = s[3], gs = s[4];
    if (CHECK) {
      const bool me = (j == self);
      d = me ? 1.0 : d;
      gz = me ? 0.0 : gz;
      gs = me ? 0.0 : gs;
    }
    const double xy = kappa - d;
    const double r0 = rcp_seed(d);
    const double e = fma(-d, r0, 1.0);
    const double pp = fma(e, e, e);
    const double r = fma(r0, pp, r0);
    const double wz = gz * r, ws = gs * r;
    acc[0] = fma(wz, s[0], acc[0]);
    acc[1] = fma(wz, s[1], acc[1]);
    acc[2] = fma(wz, s[2], acc[2]);
    acc[3] = fma(ws, s[0], acc[3]);
    acc[4] = fma(ws, s[1], acc[4]);
    acc[5] = fma(ws, s[2], acc[5]);
    const double wz2 = wz * r, ws2 = ws * r;
    double c[3], q[3], p[3];
    c[0] = fma(x[1], s[2], -(x[2] * s[1]));
    c[1] = fma(x[2], s[0], -(x[0] * s[2]));
    c[2] = fma(x[0], s[1], -(x[1] * s[0]));
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      q[b] = kx[b] - s[b];
      p[b] = fma(-xy, x[b], s[b]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double cz = wz2 * c[a];
      const double qs = ws2 * q[a];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        acc[6 + 3 * a + b] = fma(cz, q[b], acc[6 + 3 * a + b]);
        acc[6 + 3 * a + b] = fma(-qs, p[b], acc[6 + 3 * a + b]);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Planar kinds.  x = {x0, x1, surface height of the target}; kappa = eps^2.  The difference
// x - y is formed per pair (pulling the target out of the sum would cancel digits for targets far from the origin).
// ------------------------------------------------------------------------------------------------

// Incompressible2DKernels<PlaneGeometry>::kernel_vals (lpm_incompressible2d_kernels.hpp:70-84) with the source
// strength folded in: record {y0, y1, G, 0}, G = zeta A / (2 pi).  acc = {u0, u1, sum G log a}; psi = -acc[2] / 2.
template <bool CHECK>
struct Pair<kPlaneVelPsi, CHECK> {
  static constexpr int NLOAD = 4;
  __device__ __forceinline__ static void apply(const double* x, const double*, double kappa, double, const double* s,
                                               int j, int self, double* acc, const LogTables& tbl) {
    const double dx = x[0] - s[0], dy = x[1] - s[1];
    double a = fma(dx, dx, fma(dy, dy, kappa));
    double g = s[2];
    if (CHECK) {
      const bool me = (j == self);
      a = me ? 1.0 : a;
      g = me ? 0.0 : g;
    }
    const double r0 = rcp_seed(a);
    const double e = fma(-a, r0, 1.0);
    const double p = fma(e, e, e);
    const double r = fma(r0, p, r0);
    const double w = g * r;
    acc[0] = fma(-dy, w, acc[0]);
    acc[1] = fma(dx, w, acc[1]);
    acc[2] = fma(g, fast_log(a, tbl), acc[2]);
  }
};

// planar_swe_sums_rhs_pse (lpm_swe_kernels.hpp:393-445): record {y0, y1, Gz, Gs, s_j, Ap}, Gz = zeta A / (2 pi),
// Gs = sigma A / (2 pi), Ap = A / (pi pse_eps^2).  With a = |x-y|^2 + eps^2, r = 1/a, wz = Gz r, ws = Gs r and
// pA = 1 - 2 dx^2 r, pB = 2 dx dy r, pC = 1 - 2 dy^2 r the six velocity/gradient entries of the reference are
//   u      = (-dy wz + dx ws,  dx wz + dy ws)
//   du1dx1 =  wz pB + ws pA     du1dx2 = -wz pC - ws pB     du2dx1 = wz pA - ws pB     du2dx2 = -wz pB + ws pC
// (same per-pair combinations as the reference, so no cancellation is introduced), and
//   lap    = (s_j - s_i) Ap (40 (1 - q) + 10 q^2 - 2 q^3 / 3) exp(-q),  q = |x-y|^2 / pse_eps^2  (lpm_pse.hpp:66-73)
//   psi, phi accumulate G log a; the finalize kernel applies the factor -1/2.
// acc = {u0, u1, du1dx1, du1dx2, du2dx1, du2dx2, lap, sum Gz log a, sum Gs log a}.
// POT = false drops the two potentials (kPlaneSweNoPot: 7 accumulators, no log).
template <bool CHECK, bool POT>
__device__ __forceinline__ void plane_swe_pair(const double* x, double kappa, double inv_pe2, const double* s, int j, int self,
                                               double* acc, const LogTables& tbl) {
  const double dx = x[0] - s[0], dy = x[1] - s[1];
  const double a2 = dx * dx, c2 = dy * dy, b2 = dx * dy;
  const double rsq = a2 + c2;
  double a = rsq + kappa;
  double gz = s[2], gs = s[3], ap = s[5];
  if (CHECK) {
    const bool me = (j == self);
    a = me ? 1.0 : a;
    gz = me ? 0.0 : gz;
    gs = me ? 0.0 : gs;
    ap = me ? 0.0 : ap;
  }
  const double r0 = rcp_seed(a);
  const double e = fma(-a, r0, 1.0);
  const double pp = fma(e, e, e);
  const double r = fma(r0, pp, r0);
  const double wz = gz * r, ws = gs * r;
  acc[0] = fma(-dy, wz, acc[0]);
  acc[0] = fma(dx, ws, acc[0]);
  acc[1] = fma(dx, wz, acc[1]);
  acc[1] = fma(dy, ws, acc[1]);
  const double r2 = r + r;
  const double pA = fma(-r2, a2, 1.0), pC = fma(-r2, c2, 1.0), pB = r2 * b2;
  acc[2] = fma(wz, pB, acc[2]);
  acc[2] = fma(ws, pA, acc[2]);
  acc[3] = fma(-wz, pC, acc[3]);
  acc[3] = fma(-ws, pB, acc[3]);
  acc[4] = fma(wz, pA, acc[4]);
  acc[4] = fma(-ws, pB, acc[4]);
  acc[5] = fma(-wz, pB, acc[5]);
  acc[5] = fma(ws, pC, acc[5]);
  if (POT) {
    const double lg = fast_log(a, tbl);
    acc[7] = fma(gz, lg, acc[7]);
    acc[8] = fma(gs, lg, acc[8]);
  }
  const double q = rsq * inv_pe2;
  const double pre = fma(q, fma(q, fma(q, -2.0 / 3.0, 10.0), -40.0), 40.0);
  const double t = (s[4] - x[2]) * ap;
  acc[6] = fma(t * pre, fast_exp_neg(q), acc[6]);
}

template <bool CHECK>
struct Pair<kPlaneSwe, CHECK> {
  static constexpr int NLOAD = 6;
  __device__ __forceinline__ static void apply(const double* x, const double*, double kappa, double inv_pe2,
                                               const double* s, int j, int self, double* acc, const LogTables& tbl) {
    plane_swe_pair<CHECK, true>(x, kappa, inv_pe2, s, j, self, acc, tbl);
  }
};

template <bool CHECK>
struct Pair<kPlaneSweNoPot, CHECK> {
  static constexpr int NLOAD = 6;
  __device__ __forceinline__ static void apply(const double* x, const double*, double kappa, double inv_pe2,
                                               const double* s, int j, int self, double* acc, const LogTables& tbl) {
    plane_swe_pair<CHECK, false>(x, kappa, inv_pe2, s, j, self, acc, tbl);
  }
};

template <int KIND, int T, int UNROLL, bool CHECK>
__device__ __forceinline__ void chunk_loop(const double (*x)[3], const double (*kx)[3], double kappa, double aux,
                                           const double* __restrict__ sp, int j0, const int* self,
                                           double (*acc)[kind_nacc(KIND)], const LogTables& tbl) {
  constexpr int REC = kind_rec(KIND);
  static_assert(kChunk % UNROLL == 0, "source-loop unroll must divide the chunk");
#pragma unroll 1
  for (int jj = 0; jj < kChunk; jj += UNROLL) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int j = jj + u;
      constexpr int NLOAD = Pair<KIND, CHECK>::NLOAD;
      double s[NLOAD];
      const double2* s2 = reinterpret_cast<const double2*>(sp + (size_t)j * REC);
#pragma unroll
      for (int v = 0; v < NLOAD / 2; ++v) {
        const double2 t = s2[v];
        s[2 * v] = t.x;
        s[2 * v + 1] = t.y;
      }
#pragma unroll
      for (int t = 0; t < T; ++t) Pair<KIND, CHECK>::apply(x[t], kx[t], kappa, aux, s, j0 + j, self[t], acc[t], tbl);
    }
  }
}

// blocks of the two-level summation: 4 chunks = 1024 terms (sqrt(b) + sqrt(N / b) is flat around b = sqrt(N): 313 .. 2290 for
// the BASELINE meshes)
constexpr int kFlushChunks = 4;
template <int T, int NACC>
__device__ __forceinline__ void flush_block(double (*acc)[NACC], double* tot, int lanes, int tid, bool first) {
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int q = 0; q < NACC; ++q) {
      double* p = tot + (size_t)(t * NACC + q) * lanes + tid;
      *p = first ? acc[t][q] : (*p + acc[t][q]);
      acc[t][q] = 0.0;
    }
}

// Compile-time shape of one kernel instance.
//   KIND    pair body (PairKind)          T      targets per thread (register blocking)
//   NW      compute warps per CTA (+1 producer warp)     MINB   CTAs per SM the register budget allows
//   UNROLL  unroll factor of the source loop
template <int KIND_, int T_, int NW_, int MINB_, int UNROLL_>
struct PairCfg {
  static constexpr int KIND = KIND_, T = T_, NW = NW_, MINB = MINB_, UNROLL = UNROLL_;
  static constexpr int THREADS = (NW_ + 1) * 32;
  static constexpr int LANES = NW_ * 32;
  static constexpr int TB = T_ * NW_ * 32;
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB) pair_sum_kernel(const SumArgs a) {
  constexpr int KIND = C::KIND;
  constexpr int T = C::T;
  constexpr int kComputeWarps = C::NW;
  constexpr int kLanesPerCta = C::LANES;
  constexpr int REC = kind_rec(KIND);
  constexpr int NACC = kind_nacc(KIND);
  constexpr uint32_t kStageBytes = kChunk * REC * sizeof(double);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  // log tables, 16-byte aligned: 2 * kStages * 8 bytes after the ring
  double2* mtab = reinterpret_cast<double2*>(empty + kStages);
  double* ktab = reinterpret_cast<double*>(mtab + kLogMEntries);
  const LogTables tbl{mtab, ktab};
  // running totals of the two-level summation, [t * NACC + q][compute thread] (conflict-free), behind the tables / barriers
  double* tot = reinterpret_cast<double*>(smem_raw + (size_t)kStages * kStageBytes + 2 * kStages * sizeof(uint64_t) +
                                          (kind_has_log(KIND) ? log_tables_bytes() : 0));
  if (kind_has_log(KIND)) {
    for (int i = threadIdx.x; i < kLogMEntries; i += C::THREADS) mtab[i] = kLogTable[i];
    for (int i = threadIdx.x; i < kLogKEntries; i += C::THREADS) ktab[i] = fast_log_ktab_entry(i);
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kComputeWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const long n_items = (long)a.n_tb * a.n_sc;
  const int grid = gridDim.x;
  const long it0 = ((long)blockIdx.x * n_items) / grid;
  const long it1 = ((long)(blockIdx.x + 1) * n_items) / grid;

  if (warp == kComputeWarps) {
    // ---- producer warp: one lane streams source chunks through the ring ----
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int sc = (int)(it0 % a.n_sc);
      for (long it = it0; it < it1; ++it) {
        mbar_wait(empty + s, ph ^ 1u);
        mbar_arrive_expect_tx(full + s, kStageBytes);
        tma_load_1d(stage + (size_t)s * kChunk * REC, a.packed + (size_t)sc * kChunk * REC, kStageBytes, full + s);
        if (++sc == a.n_sc) sc = 0;
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    return;
  }

  // ---- compute warps ----
  const int tid = threadIdx.x;  // 0 .. kLanesPerCta-1
  constexpr int TB = C::TB;
  int s = 0;
  uint32_t ph = 0;
  long it = it0;
  while (it < it1) {
    const int tb = (int)(it / a.n_sc);
    int sc = (int)(it - (long)tb * a.n_sc);
    const long it_end = min(it1, (long)(tb + 1) * a.n_sc);

    double x[T][3], kx[T][3], acc[T][NACC];
    int self[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const long tg = (long)tb * TB + t * kLanesPerCta + tid;
      const bool valid = tg < a.n_tgt;
      const long ge = (valid && a.tgt_map) ? (long)a.tgt_map[tg] : tg;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        x[t][k] = valid ? a.tgt(ge, k) : 0.0;
        kx[t][k] = a.kappa * x[t][k];
      }
      self[t] = (valid && a.self_idx) ? a.self_idx[ge] : -1;
#pragma unroll
      for (int q = 0; q < NACC; ++q) acc[t][q] = 0.0;
    }

    // two-level summation (kind_two_level): the register accumulators hold one BLOCK of kFlushChunks * 256 terms (they restart
    // from zero) and are added to this thread's running totals in shared memory once per block.  The factored velocity sum
    // u = x cross sum(G y / d) carries a component of M along x that the cross product cancels (|M| / |u| ~ 2..50), so the
    // rounding of a single running sum over N sources shows up amplified in u: 4.3e-13 of max|u| at cubed-7 and 1.2e-12 at
    // icos-8 against a long-double sum, where the reference's own sequential sum has 1.8e-13 (profiles/r2e_parity_errors.jsonl).
    // With block partials the accumulated rounding scales with sqrt(b) + sqrt(N / b) instead of sqrt(N): 3.0e-13 at icos-8
    // (r2f).  The totals live in shared memory so that the inner loop keeps its registers: holding them in registers cost 12 %
    // of the kernel's speed (r2f: 54.9 -> 62.5 ms per BVERK4 step at cubed-7); a flush per 256-term chunk cost 2 % (r2g).
    bool first_block = true;
    while (it < it_end) {
      const long blk_end = kind_two_level(KIND) ? min(it_end, it + (long)kFlushChunks) : it_end;
      for (; it < blk_end; ++it, ++sc) {
        mbar_wait(full + s, ph);
        const double* sp = stage + (size_t)s * kChunk * REC;
        const int j0 = sc * kChunk;
        bool hit = false;
#pragma unroll
        for (int t = 0; t < T; ++t) hit |= (unsigned)(self[t] - j0) < (unsigned)kChunk;
        if (__any_sync(0xffffffffu, hit))
          chunk_loop<KIND, T, C::UNROLL, true>(x, kx, a.kappa, a.aux, sp, j0, self, acc, tbl);
        else
          chunk_loop<KIND, T, C::UNROLL, false>(x, kx, a.kappa, a.aux, sp, j0, self, acc, tbl);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (kind_two_level(KIND)) {  // after the ring slot has been handed back
        flush_block<T, NACC>(acc, tot, kLanesPerCta, tid, first_block);
        first_block = false;
      }
    }

    // flush this CTA's contribution to target block tb into its slot
    const int slot = blockIdx.x - cta_of_item((long)tb * a.n_sc, grid, n_items);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const long tg = (long)tb * TB + t * kLanesPerCta + tid;
#pragma unroll
      for (int q = 0; q < NACC; ++q)
        a.part[((long)slot * NACC + q) * a.n_tgt_pad + tg] =
            kind_two_level(KIND) ? tot[(size_t)(t * NACC + q) * kLanesPerCta + tid] : acc[t][q];
    }
  }
}


}  // namespace lpmx
#endif
