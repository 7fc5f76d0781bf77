// lpm/lpm_gallery.hpp -- the initial-condition functors the four benchmark configs use.
//   SolidBodyRotation, GaussianVortexSphere, RossbyHaurwitz54, SphereTestCase2Vorticity
//                                     src/lpm_vorticity_gallery.hpp:31-57,59-102,104-147,264-278
//   ZeroFunctor, UniformDepthSurface, SphereTestCase2InitialSurface   src/lpm_surface_gallery.hpp:92-134
//   RossbyWave54Velocity               src/lpm_velocity_gallery.hpp:250-288
#ifndef LPM_SHIM_GALLERY_HPP
#define LPM_SHIM_GALLERY_HPP

#include <array>

#include "lpm_geometry.hpp"

namespace Lpm {

struct SolidBodyRotation {
  typedef SphereGeometry geo;
  static constexpr Real OMEGA = 2 * constants::PI;
  static constexpr bool IsVorticity = true;
  Real operator()(const Real&, const Real&, const Real& z) const { return 2 * OMEGA * z; }
  template <typename PtView>
  Real operator()(const PtView& pt) const { return 2 * OMEGA * pt[2]; }
  std::string name() const { return "rotation"; }
  void init_velocity(Real& u, Real& v, Real& w, const Real& x, const Real& y, const Real&) const {
    u = -OMEGA * y;
    v = OMEGA * x;
    w = 0;
  }
};

struct GaussianVortexSphere {
  typedef SphereGeometry geo;
  static constexpr bool IsVorticity = true;
  Real gauss_const, vortex_strength, shape_parameter;
  std::array<Real, 3> xyz_ctr;
  GaussianVortexSphere(const Real str = 4 * constants::PI, const Real b = 4, const Real init_lon = 0,
                       const Real init_lat = constants::PI / 20)
      : gauss_const(0), vortex_strength(str), shape_parameter(b),
        xyz_ctr{std::cos(init_lon) * std::cos(init_lat), std::sin(init_lon) * std::cos(init_lat), std::sin(init_lat)} {}
  void set_gauss_const(const Real vorticity_sum) { gauss_const = vorticity_sum / (4 * constants::PI); }
  Real operator()(const Real& x, const Real& y, const Real& z) const {
    const Real distsq = 1.0 - x * xyz_ctr[0] - y * xyz_ctr[1] - z * xyz_ctr[2];
    return vortex_strength * std::exp(-square(shape_parameter) * distsq) - gauss_const;
  }
  template <typename PtType>
  Real operator()(const PtType& xyz) const { return (*this)(xyz[0], xyz[1], xyz[2]); }
  std::string name() const { return "SphericalGaussianVortex"; }
};

struct RossbyHaurwitz54 {
  typedef SphereGeometry geo;
  static constexpr bool IsVorticity = true;
  Real u0, rh54_amplitude;
  RossbyHaurwitz54(const Real zonal_background_velocity = 0, const Real wave_amp = 1)
      : u0(zonal_background_velocity), rh54_amplitude(wave_amp) {}
  std::string name() const { return "RossbyHaurwitz54"; }
  void set_stationary_wave_speed(const Real& Omega = 2 * constants::PI) { u0 = Omega / 14; }
  Real legendreP54(const Real z) const { return z * square(square(z) - 1); }
  Real operator()(const Real& x, const Real& y, const Real& z) const {
    return 2 * u0 * z + 30 * rh54_amplitude * std::cos(4 * atan4(y, x)) * legendreP54(z);
  }
  template <typename PtType>
  Real operator()(const PtType& xyz) const { return (*this)(xyz[0], xyz[1], xyz[2]); }
};

struct RossbyWave54Velocity {
  typedef SphereGeometry geo;
  static constexpr Int ndim = 3;
  Real background_rotation, rh54_amplitude;
  RossbyWave54Velocity(const Real u0 = 0, const Real amp = 1) : background_rotation(u0), rh54_amplitude(amp) {}
  RossbyWave54Velocity(const RossbyHaurwitz54& v) : background_rotation(v.u0), rh54_amplitude(v.rh54_amplitude) {}
  std::string name() const { return "RossbyWave54Velocity"; }
  template <typename CV>
  std::array<Real, 3> operator()(const CV x, const Real&) const {
    const Real lat = SphereGeometry::latitude(x), lon = SphereGeometry::longitude(x);
    const Real u = rh54_amplitude * 0.5 * std::cos(4 * lon) * cube(std::cos(lat)) * (5 * std::cos(2 * lat) - 3);
    const Real v = rh54_amplitude * 4 * cube(std::cos(lat)) * std::sin(lat) * std::sin(4 * lon);
    return {-background_rotation * x[1] - u * std::sin(lon) - v * std::sin(lat) * std::cos(lon),
            background_rotation * x[0] + u * std::cos(lon) - v * std::sin(lat) * std::sin(lon), v * std::cos(lat)};
  }
};

struct SphereTestCase2Vorticity {
  static constexpr Real sphere_radius = 1.0;
  static constexpr Real u0 = 2 * constants::PI / 12;
  static constexpr bool IsVorticity = true;
  template <typename CV>
  Real operator()(const CV& xyz) const { return 2 * u0 * xyz[2]; }
  std::string name() const { return "SphereTestCase2Vorticity"; }
};

struct ZeroFunctor {
  template <typename CV>
  Real operator()(const CV) const { return 0; }
  template <typename CV>
  Real laplacian(const CV) const { return 0; }
  std::string name() const { return "ZeroFunctor"; }
};

struct UniformDepthSurface {
  Real H0;
  explicit UniformDepthSurface(const Real h0 = 1) : H0(h0) {}
  template <typename CV>
  Real operator()(const CV) const { return H0; }
};

struct SphereTestCase2InitialSurface {
  static constexpr Real h0 = 10;
  static constexpr Real g = 1.0;
  static constexpr Real Omega = 2 * constants::PI;
  static constexpr Real u0 = 2 * constants::PI / 12;
  template <typename CV>
  Real operator()(const CV xyz) const { return h0 + Omega * u0 * (1 - square(xyz[2])) / g; }
  std::string name() const { return "SphereTestCase2InitialSurface"; }
};

}  // namespace Lpm
#endif
