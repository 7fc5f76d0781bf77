// sphere_rh54 -- Rossby-Haurwitz wave 4 with Incompressible2D + Incompressible2DRK2 (reference:
// examples/sphere_rh54.cpp; RH54 vorticity with the stationary wave speed u0 = Omega/14, :108-111).
//   usage: sphere_rh54 [-s cubed|icos] [-d depth] [-tf tfinal] [-n nsteps] [-eps smoothing]
#include "sphere_ic2d.hpp"

using namespace Lpm;

template <typename seed_type>
int run(const Options& opt) {
  CoriolisSphere coriolis;
  RossbyHaurwitz54 vorticity;
  vorticity.set_stationary_wave_speed(coriolis.Omega);
  RossbyWave54Velocity velocity(vorticity);
  auto report = [&](Incompressible2D<seed_type>& s, RossbyHaurwitz54&) {
    // velocity error against the exact RH54 velocity (examples/sphere_rh54.cpp:150-170)
    const Index nf = s.mesh.n_faces_host();
    vec3_view_type exact("velocity_exact", nf), err("vel_error", nf);
    for (Index i = 0; i < nf; ++i) {
      const auto u = velocity(s.mesh.faces.phys_crds.view.row(i), s.t);
      for (int k = 0; k < 3; ++k) {
        exact(i, k) = u[k];
        err(i, k) = s.mesh.faces.mask(i) ? 0 : s.velocity_active.view(i, k) - u[k];
      }
    }
    ErrNorms vel_err(err, exact, s.mesh.faces.area);
    std::printf("t = %g: velocity error vs exact RH54: %s\n", s.t, vel_err.info_string().c_str());
  };
  return example::run_ic2d<seed_type>("sphere_rh54", opt, vorticity, coriolis, report, report);
}

int main(int argc, char* argv[]) {
  Options opt(argc, argv);
  try {
    return opt.get_str("-s", "cubed") == "icos" ? run<IcosTriSphereSeed>(opt) : run<CubedSphereSeed>(opt);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "sphere_rh54: %s\n", e.what());
    return 2;
  }
}
