// lpmx_const_stream_body.h -- body of pair_sum_const_kernel (lpmx_const_stream.cu) over a small platform interface, so that the
// CPU suite runs the kernel's own indexing (targets per thread, padding, accumulator layout, first-launch flag, self-pair
// exclusion by compact index) with every CUDA thread a loop iteration (tests/cpp/const_stream_model.cpp).
//
// Platform P: int tid(), bid(), n_threads();  bool any_sync(bool);  double src(int i)  (record storage of the launch's bank);
//             double rcp_seed(double)  (>= 19 correct bits);  void accumulate(double* p, double v, bool first)  (*p = v or *p += v;
//             every address belongs to one thread of one launch at a time);  void launch_dependents(), wait_prior()  (programmatic
//             dependent launch on the GPU; no-ops elsewhere).
#ifndef LPMX_CONST_STREAM_BODY_H
#define LPMX_CONST_STREAM_BODY_H

#include <cmath>

#if defined(__CUDACC__)
#define LPMX_CS_HD __host__ __device__ __forceinline__
#else
#define LPMX_CS_HD inline
#endif

namespace lpmx {
namespace cs {

constexpr int kBatch = 1280;  // records per launch = one whole constant bank (61 440 of its 65 536 bytes); two banks alternate
constexpr int kRec = 6;       // doubles per record {y0, y1, y2, G*y0, G*y1, G*y2}
constexpr int kBankDoubles = (kBatch + 1) * kRec;  // + one record the pipelined loop reads ahead into and never uses

struct CsArgs {
  const double* tgt;    // target coordinates, element (i, k) at tgt[i * tgt_si + k * tgt_sk], indexed from 0
  long tgt_si, tgt_sk;
  const int* tgt_map;   // optional: target tg of this launch is element tgt_map[tg] of `tgt` / `self_idx` (a sharded solver's
                        // index list); null: element tg itself.  The accumulators are indexed by tg either way.
  const int* self_idx;  // compact source index of each target's own particle, or -1 (may be null)
  double* acc;          // [3][n_tgt_pad]
  long n_tgt_pad;
  int n_tgt;
  int j0;     // compact index of the bank's first record
  int n_rec;  // records of the bank this launch sums: kBatch, or kBatch / 2 (the kernel instance has it as a constant)
  int first;  // start from zero instead of the stored accumulators
  int prefetch_stride;  // doubles between the loads of the CTA's prefetch warp (GPU only; 0: no prefetch)
  double kappa;
};

// Same arithmetic per pair as Pair<kVel>::apply (lpmx_pair_kernel.cuh): 3 (d) + 3 (1/d from the seed) + 3 (M += r * G*y).
// The record of source j + 1 is loaded BEFORE the pairs of source j are evaluated (software pipelining: on the GPU a
// constant load to a uniform register has a latency that two warps per scheduler do not hide when it is issued right in
// front of its first use -- r2q: short-scoreboard stalls, 77 % of the FP64 pipe).  The last iteration loads record kBatch:
// the bank is declared one record longer (its content is never used).
// NREC: the trip count as a compile-time constant (a run-time bound costs the pipelined loop 2 %, r2z)
template <int T, bool CHECK, int NREC, class P>
LPMX_CS_HD void loop(P& pf, const double (&x)[T][3], const int (&self)[T], double (&acc)[T][3], int j0, double kappa) {
  double s[kRec], sn[kRec];
#pragma unroll
  for (int k = 0; k < kRec; ++k) sn[k] = pf.src(k);
#pragma unroll 2
  for (int j = 0; j < NREC; ++j) {
#pragma unroll
    for (int k = 0; k < kRec; ++k) {
      s[k] = sn[k];
      sn[k] = pf.src(kRec * (j + 1) + k);
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const double d = fma(-x[t][0], s[0], fma(-x[t][1], s[1], fma(-x[t][2], s[2], kappa)));
      const double r0 = pf.rcp_seed(d);
      const double e = fma(-d, r0, 1.0);
      const double p = fma(e, e, e);
      double r = fma(r0, p, r0);
      if (CHECK) r = (j0 + j == self[t]) ? 0.0 : r;
      acc[t][0] = fma(r, s[3], acc[t][0]);
      acc[t][1] = fma(r, s[4], acc[t][1]);
      acc[t][2] = fma(r, s[5], acc[t][2]);
    }
  }
}

template <int T, int NREC = kBatch, class P>
LPMX_CS_HD void body(P& pf, const CsArgs& a) {
  pf.launch_dependents();
  const int lanes = pf.n_threads();
  const long base_t = (long)pf.bid() * ((long)T * lanes) + pf.tid();
  double x[T][3], acc[T][3];
  int self[T];
  bool hit = false;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base_t + (long)t * lanes;  // < n_tgt_pad by construction of the grid
    const bool valid = tg < a.n_tgt;
    const long ge = (valid && a.tgt_map) ? (long)a.tgt_map[tg] : tg;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      x[t][k] = valid ? a.tgt[ge * a.tgt_si + k * a.tgt_sk] : 0.0;  // a zero target sees d = kappa: finite, never read back
      acc[t][k] = 0.0;  // two-level summation: this launch's 1 280 terms start from zero (see the store below)
    }
    self[t] = (valid && a.self_idx) ? a.self_idx[ge] : -1;
    hit |= (unsigned)(self[t] - a.j0) < (unsigned)NREC;
  }
  if (pf.any_sync(hit))
    loop<T, true, NREC>(pf, x, self, acc, a.j0, a.kappa);
  else
    loop<T, false, NREC>(pf, x, self, acc, a.j0, a.kappa);
  pf.wait_prior();  // the accumulators take the launches' sums in launch order
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base_t + (long)t * lanes;
    // the launch's partial sum is added to the running total once: the accumulated rounding of the factored sum scales with
    // sqrt(1280) + sqrt(N / 1280) instead of sqrt(N) (lpmx_pair_kernel.cuh, "two-level summation")
#pragma unroll
    for (int k = 0; k < 3; ++k) pf.accumulate(a.acc + (long)k * a.n_tgt_pad + tg, acc[t][k], a.first != 0);
  }
}

}  // namespace cs
}  // namespace lpmx

#endif
