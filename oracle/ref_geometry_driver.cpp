// ref_geometry_driver.cpp -- TEST INFRASTRUCTURE.  The reference's geometry functions (src/lpm_geometry.hpp: SphereGeometry /
// PlaneGeometry polygon_area, barycenter, midpoint, distance), compiled IN PLACE from /root/reference against oracle/kokkos_shim
// and exported with plain-pointer signatures, so that tests/test_mesh.py can check the host mesh generator's arithmetic
// (lpm_b200/csrc/lpmx_mesh.cpp) against them bit for bit.  Built WITHOUT FMA contraction like the generator (the reference's
// values depend on its compiler flags; the formula is what is pinned).  Output: oracle/_ref/liblpm_ref_geometry.so (git-ignored).
// Nothing under lpm_b200/ or include/ uses it.
#include "lpm_geometry.hpp"
#include <cstdio>
using namespace Lpm;
extern "C" {
// ctr[3], verts[n][3]
double ref_sphere_polygon_area(const double* ctr, const double* verts, int n) {
  Kokkos::View<Real[3], Kokkos::HostSpace> c("c");
  Kokkos::View<Real**, Kokkos::HostSpace> v("v", n, 3);
  for (int k = 0; k < 3; ++k) c(k) = ctr[k];
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) v(i, k) = verts[3 * i + k];
  return SphereGeometry::polygon_area(c, v, n);
}
double ref_plane_polygon_area(const double* ctr, const double* verts, int n) {
  Kokkos::View<Real[2], Kokkos::HostSpace> c("c");
  Kokkos::View<Real**, Kokkos::HostSpace> v("v", n, 2);
  for (int k = 0; k < 2; ++k) c(k) = ctr[k];
  for (int i = 0; i < n; ++i) for (int k = 0; k < 2; ++k) v(i, k) = verts[2 * i + k];
  return PlaneGeometry::polygon_area(c, v, n);
}
void ref_sphere_midpoint(double* out, const double* a, const double* b) { SphereGeometry::midpoint(out, a, b); }
void ref_plane_midpoint(double* out, const double* a, const double* b) { PlaneGeometry::midpoint(out, a, b); }
void ref_sphere_barycenter(double* out, const double* verts, int n) {
  Kokkos::View<Real**, Kokkos::HostSpace> v("v", n, 3);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) v(i, k) = verts[3 * i + k];
  SphereGeometry::barycenter(out, v, n);
}
void ref_plane_barycenter(double* out, const double* verts, int n) {
  Kokkos::View<Real**, Kokkos::HostSpace> v("v", n, 2);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 2; ++k) v(i, k) = verts[2 * i + k];
  PlaneGeometry::barycenter(out, v, n);
}
double ref_sphere_distance(const double* a, const double* b) { return SphereGeometry::distance(a, b); }
}
