// lpmx_const_stream_body.h -- body of pair_sum_const_kernel (lpmx_const_stream.cu) over a small platform interface, so that the
// CPU suite runs the kernel's own indexing (targets per thread, padding, accumulator layout, first-launch flag, self-pair
// exclusion by compact index) with every CUDA thread a loop iteration (tests/cpp/const_stream_model.cpp).
//
// Platform P: int tid(), bid(), n_threads();  bool any_sync(bool);  double src(int i)  (record storage of the two halves);
//             double rcp_seed(double)  (>= 19 correct bits).
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

constexpr int kHalf = 640;  // records per half of the constant bank
constexpr int kRec = 6;     // doubles per record {y0, y1, y2, G*y0, G*y1, G*y2}

struct CsArgs {
  const double* tgt;    // this launch's targets, element (i, k) at tgt[i * tgt_si + k * tgt_sk], indexed from 0
  long tgt_si, tgt_sk;
  const int* self_idx;  // compact source index of each target's own particle, or -1 (may be null)
  double* acc;          // [3][n_tgt_pad]
  long n_tgt_pad;
  int n_tgt;
  int half;   // which half of the bank this launch reads
  int j0;     // compact index of the half's first record
  int first;  // start from zero instead of the stored accumulators
  double kappa;
};

// Same arithmetic per pair as Pair<kVel>::apply (lpmx_pair_kernel.cuh): 3 (d) + 3 (1/d from the seed) + 3 (M += r * G*y).
template <int T, bool CHECK, class P>
LPMX_CS_HD void loop(P& pf, const double (&x)[T][3], const int (&self)[T], double (&acc)[T][3], int base, int j0, double kappa) {
#pragma unroll 2
  for (int j = 0; j < kHalf; ++j) {
    const double s0 = pf.src(base + kRec * j), s1 = pf.src(base + kRec * j + 1), s2 = pf.src(base + kRec * j + 2);
    const double s3 = pf.src(base + kRec * j + 3), s4 = pf.src(base + kRec * j + 4), s5 = pf.src(base + kRec * j + 5);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const double d = fma(-x[t][0], s0, fma(-x[t][1], s1, fma(-x[t][2], s2, kappa)));
      const double r0 = pf.rcp_seed(d);
      const double e = fma(-d, r0, 1.0);
      const double p = fma(e, e, e);
      double r = fma(r0, p, r0);
      if (CHECK) r = (j0 + j == self[t]) ? 0.0 : r;
      acc[t][0] = fma(r, s3, acc[t][0]);
      acc[t][1] = fma(r, s4, acc[t][1]);
      acc[t][2] = fma(r, s5, acc[t][2]);
    }
  }
}

template <int T, class P>
LPMX_CS_HD void body(P& pf, const CsArgs& a) {
  const int lanes = pf.n_threads();
  const long base_t = (long)pf.bid() * ((long)T * lanes) + pf.tid();
  double x[T][3], acc[T][3];
  int self[T];
  bool hit = false;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base_t + (long)t * lanes;  // < n_tgt_pad by construction of the grid
    const bool valid = tg < a.n_tgt;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      x[t][k] = valid ? a.tgt[tg * a.tgt_si + k * a.tgt_sk] : 0.0;  // a zero target sees d = kappa: finite, never read back
      acc[t][k] = 0.0;  // two-level summation: this launch's 640 terms start from zero (see the store below)
    }
    self[t] = (valid && a.self_idx) ? a.self_idx[tg] : -1;
    hit |= (unsigned)(self[t] - a.j0) < (unsigned)kHalf;
  }
  const int base = a.half * (kHalf * kRec);
  if (pf.any_sync(hit))
    loop<T, true>(pf, x, self, acc, base, a.j0, a.kappa);
  else
    loop<T, false>(pf, x, self, acc, base, a.j0, a.kappa);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base_t + (long)t * lanes;
    // the launch's partial sum is added to the running total once: the accumulated rounding of the factored sum scales with
    // sqrt(640) + sqrt(N / 640) instead of sqrt(N) (lpmx_pair_kernel.cuh, "two-level summation")
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double* p = a.acc + (long)k * a.n_tgt_pad + tg;
      *p = a.first ? acc[t][k] : (*p + acc[t][k]);
    }
  }
}

}  // namespace cs
}  // namespace lpmx

#endif
