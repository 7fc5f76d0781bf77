// fast_log_check.cpp -- host build of the kernel's log (lpm_b200/csrc/lpmx_fast_log.h, the same source the device compiles)
// against logl over the range a mesh can produce.  Prints the largest |fast_log - log| / max(1, |log|) and the special cases.
// Compiled with -ffp-contract=off: every fma in the header is explicit, nothing else may be fused.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../lpm_b200/csrc/lpmx_fast_log.h"

using namespace lpmx;

static const LogPair kM[kLogMEntries] = {
#if LPMX_LOG_MBITS == 10
#include "../../lpm_b200/csrc/log_table_10.inc"
#elif LPMX_LOG_MBITS == 8
#include "../../lpm_b200/csrc/log_table_8.inc"
#else
#include "../../lpm_b200/csrc/log_table_7.inc"
#endif
};

int main() {
  std::vector<double> kt(kLogKEntries + 1);
  for (int e = 0; e < kLogKEntries; ++e) kt[e] = fast_log_ktab_entry(e);
  double worst = 0, worst_d = 0;
  long n = 0;
  auto check = [&](double d) {
    const double got = fast_log(d, kM, kt.data());
    const long double ref = logl((long double)d);
    const double err = (double)(fabsl((long double)got - ref) / fmaxl(1.0L, fabsl(ref)));
    if (err > worst) worst = err, worst_d = d;
    ++n;
  };
  // log-uniform sweep 1e-16 .. 4, every table boundary and its neighbours, powers of two, values next to 1
  for (int i = 0; i <= 2000000; ++i) check(std::pow(10.0, -16.0 + 16.60206 * i / 2000000.0));
  for (int i = 0; i < 1024; ++i) {
    const double b = 1.0 + i / 1024.0;  // every boundary of the widest table; the narrower ones are subsets
    check(b), check(std::nextafter(b, 0.0)), check(std::nextafter(b, 4.0)), check(b + 0.5 / 1024), check(0.5 * b), check(2 * b);
    check(b * 1e-9);
  }
  for (int k = -60; k <= 2; ++k) check(std::ldexp(1.0, k));
  check(1 - std::ldexp(1.0, -53)), check(1 + std::ldexp(1.0, -52));
  srand48(20261017);
  for (int i = 0; i < 2000000; ++i) check(std::ldexp(1.0 + drand48(), (int)(lrand48() % 45) - 43));
  const double z = fast_log(0.0, kM, kt.data()), neg = fast_log(-1.0, kM, kt.data());
  const double inf = fast_log(std::numeric_limits<double>::infinity(), kM, kt.data());
  const double nan = fast_log(std::numeric_limits<double>::quiet_NaN(), kM, kt.data());
  (void)inf, (void)nan;  // inf / NaN arguments cannot occur (d is a finite difference of finite numbers)
  std::printf("%ld %.3e %.17g %d %d\n", n, worst, worst_d, (int)!std::isfinite(z), (int)std::isnan(neg));
  return 0;
}
