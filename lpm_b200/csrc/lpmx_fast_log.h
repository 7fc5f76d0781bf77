// lpmx_fast_log.h -- log(d) for the stream-function kinds of the pair-sum kernel; __host__ __device__ so the CPU suite checks
// the very arithmetic the kernel runs (tests/cpp/fast_log_check.cpp).
//
// d = kappa - x.y (sphere) or |x - y|^2 + eps^2 (plane) is a positive normal double.  The general-purpose libdevice log
// (special cases, denormals, ~45 FP64-pipe instructions) is replaced by one table lookup and a short polynomial:
//   d = 2^k m, m in [1, 2);  i = top B mantissa bits;  c_i ~ 1 / (1 + (i + 1/2) / 2^B);  t = m c_i - 1 (one FMA, |t| <= 2^-(B+1))
//   log d = k ln2 + (-log c_i) + log1p(t),   log1p(t) = t - t^2/2 + ... +- t^n/n
// The table {c_i, -log c_i} sits in shared memory behind the source ring (a divergent LDS.128 per pair).  Its size is a
// measured trade (profiles/r1af_tune_log_variants.txt, 229 376 x 98 304 pairs on a B200; kVelPsi / kPsi in ms):
//   B = 7, degree 7 (first version)  35.4 / 26.4        B = 8, degree 5, exponent table  34.3 / 31.1
//   B = 7, degree 6                  34.7 / 25.6        B = 10, degree 4, exponent table 36.6 / 31.4
//   B = 8, degree 5  (adopted)       33.2 / 26.2        B = 7, degree 6, exponent table  33.9 / 30.9
// A wider table shortens the polynomial (the kernels are issue bound: an FP64 instruction holds the dispatch port for two
// cycles) but neighbouring lanes stop sharing entries, so the lookup costs more shared-memory wavefronts; 256 entries is
// the optimum.  Taking k ln2 (and the special cases) from a second table indexed by the exponent field removes the
// int -> double conversion and the selects, but the second lookup makes the log-only kernel LDS bound (+5 ms): the
// variant is kept behind LPMX_LOG_KTAB for the record and is off.
// Absolute error < 3e-16 max(1, |log d|) over (1e-16, 4) (tests/test_fast_log.py); non-finite for d <= 0 like std::log, so a
// target sitting exactly on a source (a divided icosahedral panel on its centre child at eps = 0) stays non-finite.
#ifndef LPMX_FAST_LOG_H
#define LPMX_FAST_LOG_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#if defined(__CUDACC__)
#define LPMX_HD __host__ __device__ __forceinline__
#else
#define LPMX_HD inline
#endif

namespace lpmx {

// Compile-time shape of the log (chosen by measurement, see above): mantissa-table bits, polynomial degree of log1p,
// and whether k ln2 comes from the exponent table (1) or from an int -> double conversion with an explicit non-finite
// select (0).  Truncation |t|^(deg+1) / (deg+1) with |t| <= 2^-(bits+1): bits 7 -> degree 6 (2e-18), 8 -> 5 (9e-18),
// 10 -> 4 (6e-18).
#ifndef LPMX_LOG_MBITS
#define LPMX_LOG_MBITS 8
#endif
#ifndef LPMX_LOG_DEG
#define LPMX_LOG_DEG (LPMX_LOG_MBITS >= 10 ? 4 : LPMX_LOG_MBITS >= 8 ? 5 : 6)
#endif
#ifndef LPMX_LOG_KTAB
#define LPMX_LOG_KTAB 0
#endif

constexpr int kLogMBits = LPMX_LOG_MBITS;
constexpr int kLogMEntries = 1 << kLogMBits;                 // mantissa table, {c_i, -log c_i}
constexpr int kLogKEntries = LPMX_LOG_KTAB ? 2049 : 0;       // exponent table, index min(exponent field, 2048)

struct LogPair {  // layout of double2
  double x, y;
};

LPMX_HD int log_hi(double d) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(d);
#else
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return (int)(u >> 32);
#endif
}
LPMX_HD int log_lo(double d) {
#if defined(__CUDA_ARCH__)
  return __double2loint(d);
#else
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return (int)(u & 0xffffffffu);
#endif
}
LPMX_HD double log_make(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  std::memcpy(&d, &u, 8);
  return d;
#endif
}

// entry e of the exponent table: (e - 1023) ln2 to within an ulp (ln2 split hi/lo), and the special cases
LPMX_HD double fast_log_ktab_entry(int e) {
  if (e == 0) return -std::numeric_limits<double>::infinity();          // d = 0 (and denormals): std::log(0)
  if (e >= 2047) return std::numeric_limits<double>::quiet_NaN();       // inf / NaN (2047), d < 0 (2048)
  const double k = (double)(e - 1023);
  return fma(k, 6.93147180369123816490e-01, k * 1.90821492927058770002e-10);
}

// log1p(t) = t + t^2 q(t), q by Horner from the highest kept term
LPMX_HD double fast_log1p_poly(double t) {
#if LPMX_LOG_DEG == 7
  double q = fma(t, 1.0 / 7.0, -1.0 / 6.0);
  q = fma(q, t, 0.2);
  q = fma(q, t, -0.25);
  q = fma(q, t, 1.0 / 3.0);
#elif LPMX_LOG_DEG == 6
  double q = fma(t, -1.0 / 6.0, 0.2);
  q = fma(q, t, -0.25);
  q = fma(q, t, 1.0 / 3.0);
#elif LPMX_LOG_DEG == 5
  double q = fma(t, 0.2, -0.25);
  q = fma(q, t, 1.0 / 3.0);
#else
  double q = fma(t, -0.25, 1.0 / 3.0);
#endif
  q = fma(q, t, -0.5);
  return fma(q, t * t, t);
}

template <typename MTab>
LPMX_HD double fast_log(double d, const MTab* __restrict__ mtab, const double* __restrict__ ktab) {
  const int hi = log_hi(d);
  const MTab cl = mtab[(hi >> (20 - kLogMBits)) & (kLogMEntries - 1)];
  const double m = log_make((hi & 0x000fffff) | 0x3ff00000, log_lo(d));
  const double t = fma(m, cl.x, -1.0);
  const double l1p = fast_log1p_poly(t);
#if LPMX_LOG_KTAB
  const unsigned e = (unsigned)hi >> 20;
  const double kl = ktab[e < 2048u ? e : 2048u];
  return kl + (cl.y + l1p);  // the two small terms first: one rounding at the magnitude of the result
#else
  (void)ktab;
  // d <= 0: the reference's std::log returns -inf / NaN there; stay non-finite (integer compare + select)
  const int k = (hi >> 20) - 1023;
  const double kd = hi > 0 ? (double)k : std::numeric_limits<double>::quiet_NaN();
  return fma(kd, 0.693147180559945309417232121458, cl.y) + l1p;
#endif
}

}  // namespace lpmx
#endif
