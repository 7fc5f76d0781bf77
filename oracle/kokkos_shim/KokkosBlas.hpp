// KokkosBlas.hpp -- the two BLAS-1 calls the reference's steppers make (lpm_bve_rk4_impl.hpp, lpm_incompressible2d_rk2_impl.hpp,
// lpm_swe_rk2_impl.hpp), with KokkosKernels' documented semantics, elementwise over whole views of any rank:
//   scal(R, a, X):                     R = a * X
//   update(alpha, X, beta, Y, gamma, Z):  Z = gamma * Z + alpha * X + beta * Y
// TEST INFRASTRUCTURE (oracle/_ref only).  Our own code, not derived from KokkosKernels.  Each product is rounded separately
// and the sums are formed left to right without FMA contraction (KokkosKernels' generic kernels are plain C++ expressions; a
// compiler may contract them, which is why the stepper comparisons carry a tolerance of a few ulp rather than bit equality).
#ifndef ORACLE_KOKKOS_SHIM_BLAS_HPP
#define ORACLE_KOKKOS_SHIM_BLAS_HPP
#include "Kokkos_Core.hpp"

namespace KokkosBlas {
template <class RV, class A, class XV>
void scal(const RV& r, const A& a, const XV& x) {
  const std::size_t n = r.size();
  auto* rp = r.data();
  const auto* xp = x.data();
#pragma omp parallel for schedule(static)
  for (std::size_t i = 0; i < n; ++i) rp[i] = a * xp[i];
}
template <class XV, class YV, class ZV>
void update(const double alpha, const XV& x, const double beta, const YV& y, const double gamma, const ZV& z) {
  const std::size_t n = z.size();
  auto* zp = z.data();
  const auto* xp = x.data();
  const auto* yp = y.data();
  if (gamma == 0.0) {
#pragma omp parallel for schedule(static)
    for (std::size_t i = 0; i < n; ++i) zp[i] = alpha * xp[i] + beta * yp[i];
  } else {
#pragma omp parallel for schedule(static)
    for (std::size_t i = 0; i < n; ++i) zp[i] = gamma * zp[i] + alpha * xp[i] + beta * yp[i];
  }
}
}  // namespace KokkosBlas
#endif
