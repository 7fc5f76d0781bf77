// lpm/lpm_geometry.hpp -- SphereGeometry: static helpers of src/lpm_geometry.hpp:256-643 that the drivers use,
// and atan4 (src/util/lpm_math.hpp:66-104).
#ifndef LPM_SHIM_GEOMETRY_HPP
#define LPM_SHIM_GEOMETRY_HPP

#include "lpm_views.hpp"

namespace Lpm {

/// Longitude-valued arctangent in [0, 2 pi) (src/util/lpm_math.hpp:66-104)
inline Real atan4(const Real y, const Real x) {
  Real result = 0;
  const bool xz = std::abs(x) < constants::ZERO_TOL, yz = std::abs(y) < constants::ZERO_TOL;
  if (xz) {
    if (!yz) result = (y > 0 ? 0.5 * constants::PI : 1.5 * constants::PI);
  } else if (yz) {
    result = (x > 0 ? 0 : constants::PI);
  } else {
    const Real theta = std::atan2(std::abs(y), std::abs(x));
    if (x > 0 && y > 0) result = theta;
    else if (x < 0 && y > 0) result = constants::PI - theta;
    else if (x < 0 && y < 0) result = constants::PI + theta;
    else result = 2 * constants::PI - theta;
  }
  return result;
}

struct SphereGeometry {
  static constexpr Int ndim = 3;
  typedef vec3_view_type crd_view_type;
  typedef vec3_view_type vec_view_type;
  static std::string id_string() { return "SphereGeometry"; }
  template <typename A, typename B>
  static Real dot(const A& a, const B& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
  template <typename A>
  static Real norm2(const A& a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
  template <typename A>
  static Real mag(const A& a) { return std::sqrt(norm2(a)); }
  template <typename C, typename A, typename B>
  static void cross(C& c, const A& a, const B& b) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
  }
  template <typename A>
  static Real latitude(const A& x) { return std::atan2(x[2], std::sqrt(x[0] * x[0] + x[1] * x[1])); }
  template <typename A>
  static Real longitude(const A& x) { return atan4(x[1], x[0]); }
};

}  // namespace Lpm
#endif
