// gallery_check.cpp -- host-only: evaluates the C++ shim's initial-condition functors (include/lpm/lpm_gallery.hpp,
// lpm_plane.hpp, lpm_coriolis.hpp) on points read from a raw float64 file and writes the values, in a fixed order, to another.
// Usage: gallery_check <sphere_points.bin> <n_sphere> <plane_points.bin> <n_plane> <out.bin>
// tests/test_gallery.py compares the output with tests/golden/ref_gallery.npz (the reference's functors compiled in place).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lpm/lpm.hpp"

using namespace Lpm;

static std::vector<double> read_all(const char* path, size_t n) {
  std::vector<double> v(n);
  FILE* f = std::fopen(path, "rb");
  if (!f || std::fread(v.data(), sizeof(double), n, f) != n) std::exit(2);
  std::fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 6) return 1;
  const int ns = std::atoi(argv[2]), np = std::atoi(argv[4]);
  const std::vector<double> s = read_all(argv[1], 3 * (size_t)ns), p = read_all(argv[3], 2 * (size_t)np);
  std::vector<double> out;
  auto sphere = [&](auto f) {
    for (int i = 0; i < ns; ++i) out.push_back(f(&s[3 * (size_t)i]));
  };
  auto plane = [&](auto f) {
    for (int i = 0; i < np; ++i) out.push_back(f(&p[2 * (size_t)i]));
  };
  // order must match ORDER in tests/test_gallery.py
  sphere([](const double* x) { return SolidBodyRotation()(x); });
  sphere([](const double* x) { return GaussianVortexSphere()(x); });
  sphere([](const double* x) {
    GaussianVortexSphere g;
    g.set_gauss_const(0.37);
    return g(x);
  });
  sphere([](const double* x) { return RossbyHaurwitz54(0.0, 1.0)(x); });
  sphere([](const double* x) {
    RossbyHaurwitz54 f;
    f.set_stationary_wave_speed();
    return f(x);
  });
  sphere([](const double* x) { return RossbyHaurwitz54(0.3, 0.25)(x); });
  sphere([](const double* x) { return SphereTestCase2Vorticity()(x); });
  sphere([](const double* x) { return SphereTestCase2InitialSurface()(x); });
  sphere([](const double* x) { return CoriolisSphere(2 * constants::PI).f(x); });
  plane([](const double* x) { return PlanarGaussianMountain()(x); });
  plane([](const double* x) { return PlanarGaussianMountain().laplacian(x); });
  plane([](const double* x) { return PlanarGaussianSurfacePerturbation()(x); });
  plane([](const double* x) { return CollidingDipolePairPlane()(x); });
  plane([](const double* x) { return CoriolisBetaPlane(0.7, 0.2).f(x); });
  {
    RossbyHaurwitz54 f;
    f.set_stationary_wave_speed();
    const RossbyWave54Velocity vel(f);
    for (int i = 0; i < ns; ++i) {
      const auto u = vel(&s[3 * (size_t)i], 0.0);
      for (int k = 0; k < 3; ++k) out.push_back(u[k]);
    }
  }
  sphere([](const double* x) { return atan4(x[1], x[0]); });
  FILE* f = std::fopen(argv[5], "wb");
  if (!f || std::fwrite(out.data(), sizeof(double), out.size(), f) != out.size()) return 3;
  std::fclose(f);
  return 0;
}
