// compose/siqk_sqr.hpp -- stand-in for the one COMPOSE header the reference's mesh class includes
// (/root/reference/src/mesh/lpm_polymesh2d.hpp:13).  COMPOSE is absent from /root/reference and from the image.
// The only use is PolyMesh2d::quad_ref (lpm_polymesh2d.hpp:665, reference coordinates of a point in a spherical quadrilateral
// for the bilinear interpolation of lpm_bivar_remesh), which nothing under oracle/ calls: the stand-in aborts if it ever is.
// TEST INFRASTRUCTURE (oracle/_ref only); our own code.
#ifndef ORACLE_SHIM_SIQK_SQR_HPP
#define ORACLE_SHIM_SIQK_SQR_HPP
#include <cstdio>
#include <cstdlib>
namespace siqk {
namespace sqr {
template <class V, class Q>
inline void calc_sphere_to_ref(const V&, const Q*, const double*, double&, double&) {
  std::fprintf(stderr, "siqk::sqr::calc_sphere_to_ref: COMPOSE is not available in the oracle build\n");
  std::abort();
}
}  // namespace sqr
}  // namespace siqk
#endif
