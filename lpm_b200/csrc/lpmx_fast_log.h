// lpmx_fast_log.h -- log(d) for the stream-function kinds of the pair-sum kernel; __host__ __device__ so the CPU suite checks
// the very arithmetic the kernel runs (tests/cpp/fast_log_check.cpp).
//
// d = kappa - x.y (sphere) or |x - y|^2 + eps^2 (plane) is a positive normal double.  The general-purpose libdevice log
// (special cases, denormals, ~45 FP64-pipe instructions) is replaced by two table lookups and a short polynomial:
//   d = 2^k m, m in [1, 2);  i = top 10 mantissa bits;  c_i ~ 1 / (1 + (i + 1/2) / 1024);  t = m c_i - 1 (one FMA, |t| <= 2^-11)
//   log d = k ln2 + (-log c_i) + log1p(t),   log1p(t) = t - t^2/2 + t^3/3 - t^4/4 + t^5/5   (|t|^6 / 6 < 2.3e-21)
// 8 FP64-pipe instructions (the 128-entry version needed degree 7 and an int -> double conversion: 11), no I2F and no
// selects: k ln2 comes from a second table indexed by the exponent field, whose entries also carry the special cases --
// -inf for d = 0 (std::log(0)), NaN for d < 0, inf and NaN -- so a target sitting exactly on a source (a divided
// icosahedral panel on its centre child at eps = 0) stays non-finite like the reference's std::log, with no compare in the
// loop.  The pair kernels are issue bound (every FP64 instruction holds the dispatch port for two cycles, every other one
// for one), so both the three FP64 and the four integer/select instructions saved per pair count.
// Tables (shared memory, 32.8 KB): mtab[1024] = {c_i, -log c_i} (lpmx log_table.inc, -log c_i from 60-digit arithmetic),
// ktab[2049] = (e - 1023) ln2 built at kernel start (fast_log_ktab_entry).  Absolute error < 3e-16 max(1, |log d|).
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

constexpr int kLogMEntries = 1024;  // mantissa table, {c_i, -log c_i}
constexpr int kLogKEntries = 2049;  // exponent table, index min(exponent field, 2048)

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

template <typename MTab>
LPMX_HD double fast_log(double d, const MTab* __restrict__ mtab, const double* __restrict__ ktab) {
  const int hi = log_hi(d);
  const unsigned e = (unsigned)hi >> 20;
  const double kl = ktab[e < 2048u ? e : 2048u];
  const MTab cl = mtab[(hi >> 10) & 1023];
  const double m = log_make((hi & 0x000fffff) | 0x3ff00000, log_lo(d));
  const double t = fma(m, cl.x, -1.0);
  double q = fma(t, 0.2, -0.25);
  q = fma(q, t, 1.0 / 3.0);
  q = fma(q, t, -0.5);
  const double l1p = fma(q, t * t, t);
  return kl + (cl.y + l1p);  // the two small terms first: one rounding at the magnitude of the result
}

}  // namespace lpmx
#endif
