// sphere_gaussian_vortex -- a Gaussian vortex on the rotating sphere with Incompressible2D + Incompressible2DRK2
// (reference: examples/sphere_gaussian_vortex.cpp; the vorticity is shifted so that its integral vanishes,
// set_gauss_const(total_vorticity), :83-86).
//   usage: sphere_gaussian_vortex [-s cubed|icos] [-d depth] [-tf tfinal] [-n nsteps] [-eps smoothing]
#include "sphere_ic2d.hpp"

using namespace Lpm;

template <typename seed_type>
int run(const Options& opt) {
  CoriolisSphere coriolis;
  GaussianVortexSphere gauss_vort;
  auto setup = [&](Incompressible2D<seed_type>& s, GaussianVortexSphere& v) {
    const Real total_vort0 = s.total_vorticity();
    v.set_gauss_const(total_vort0);
    s.init_vorticity(v);
    std::printf("gauss_const = %.12e; total vorticity after the shift = %.3e\n", v.gauss_const, s.total_vorticity());
  };
  auto nothing = [](Incompressible2D<seed_type>&, GaussianVortexSphere&) {};
  return example::run_ic2d<seed_type>("sphere_gaussian_vortex", opt, gauss_vort, coriolis, setup, nothing);
}

int main(int argc, char* argv[]) {
  Options opt(argc, argv);
  try {
    return opt.get_str("-s", "cubed") == "icos" ? run<IcosTriSphereSeed>(opt) : run<CubedSphereSeed>(opt);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "sphere_gaussian_vortex: %s\n", e.what());
    return 2;
  }
}
