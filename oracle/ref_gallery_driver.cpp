// ref_gallery_driver.cpp -- C entry points around the REFERENCE's own initial-condition functors
// (/root/reference/src/lpm_vorticity_gallery.hpp, lpm_velocity_gallery.hpp, lpm_surface_gallery.hpp, lpm_coriolis.hpp and the
// Bessel functions of util/lpm_math.hpp; compiled in place, never copied) against oracle/kokkos_shim.  Linked into
// oracle/_ref/liblpm_ref.so.  TEST INFRASTRUCTURE: generates tests/golden/ref_gallery.npz
// (tests/golden/make_ref_gallery_golden.py), which pins lpm_b200/gallery.py (what bench.py and the parity tests feed the
// engine) and the C++ shim's gallery (include/lpm/lpm_gallery.hpp, lpm_plane.hpp: what the example drivers use).
// "The same initial conditions" is part of the parity claim, so the functors are checked like the kernels are.
#include <cstdint>

#include "lpm_coriolis.hpp"
#include "lpm_surface_gallery.hpp"
#include "lpm_velocity_gallery.hpp"
#include "lpm_vorticity_gallery.hpp"

using namespace Lpm;

namespace {
template <typename F>
void eval3(const F& f, int n, const double* pts, double* out) {
  SphereGeometry::crd_view_type x(const_cast<double*>(pts), n);
  for (int i = 0; i < n; ++i) out[i] = f(Kokkos::subview(x, i, Kokkos::ALL));
}
template <typename F>
void eval2(const F& f, int n, const double* pts, double* out) {
  PlaneGeometry::crd_view_type x(const_cast<double*>(pts), n);
  for (int i = 0; i < n; ++i) out[i] = f(Kokkos::subview(x, i, Kokkos::ALL));
}
}  // namespace

extern "C" {

// id: 0 SolidBodyRotation, 1 GaussianVortexSphere() with set_gauss_const(p0), 2 RossbyHaurwitz54(p0, p1),
//     3 SphereTestCase2Vorticity, 4 SphereTestCase2InitialSurface, 5 CoriolisSphere(p0).f  -- points are n x 3
//     10 PlanarGaussianMountain, 11 PlanarGaussianMountain::laplacian, 12 PlanarGaussianSurfacePerturbation,
//     13 CollidingDipolePairPlane(), 14 CoriolisBetaPlane(p0, p1).f                          -- points are n x 2
int ref_gallery_scalar(int id, int n, const double* pts, double p0, double p1, double* out) {
  switch (id) {
    case 0: eval3(SolidBodyRotation(), n, pts, out); return 0;
    case 1: {
      GaussianVortexSphere g;
      g.set_gauss_const(p0);
      eval3(g, n, pts, out);
      return 0;
    }
    case 2: eval3(RossbyHaurwitz54(p0, p1), n, pts, out); return 0;
    case 3: eval3(SphereTestCase2Vorticity(), n, pts, out); return 0;
    case 4: eval3(SphereTestCase2InitialSurface(), n, pts, out); return 0;
    case 5: {
      const CoriolisSphere c(p0);
      eval3([&](const auto& x) { return c.f(x); }, n, pts, out);
      return 0;
    }
    case 10: eval2(PlanarGaussianMountain(), n, pts, out); return 0;
    case 11: {
      const PlanarGaussianMountain m;
      eval2([&](const auto& x) { return m.laplacian(x); }, n, pts, out);
      return 0;
    }
    case 12: eval2(PlanarGaussianSurfacePerturbation(), n, pts, out); return 0;
    case 13: eval2(CollidingDipolePairPlane(), n, pts, out); return 0;
    case 14: {
      const CoriolisBetaPlane c(p0, p1);
      eval2([&](const auto& x) { return c.f(x); }, n, pts, out);
      return 0;
    }
    default: return 1;
  }
}

// RossbyWave54Velocity(u0, amp)(x, t): out is n x 3
void ref_gallery_rh54_velocity(int n, const double* pts, double u0, double amp, double* out) {
  const RossbyWave54Velocity vel(u0, amp);
  SphereGeometry::crd_view_type x(const_cast<double*>(pts), n);
  for (int i = 0; i < n; ++i) {
    const auto u = vel(Kokkos::subview(x, i, Kokkos::ALL), 0.0);
    for (int k = 0; k < 3; ++k) out[3 * i + k] = u[k];
  }
}

// RossbyHaurwitz54::set_stationary_wave_speed(Omega) -> u0
double ref_gallery_rh54_stationary_u0(double Omega) {
  RossbyHaurwitz54 f;
  f.set_stationary_wave_speed(Omega);
  return f.u0;
}

// bessel_j0 / bessel_j1 of util/lpm_math.hpp (the Lamb dipole's), order = 0 or 1
void ref_bessel_j(int order, int n, const double* x, double* out) {
  for (int i = 0; i < n; ++i) out[i] = order == 0 ? bessel_j0(x[i]) : bessel_j1(x[i]);
}

// atan4 (util/lpm_math.hpp:66-104)
void ref_atan4(int n, const double* y, const double* x, double* out) {
  for (int i = 0; i < n; ++i) out[i] = atan4(y[i], x[i]);
}

}  // extern "C"
