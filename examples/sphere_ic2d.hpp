// sphere_ic2d.hpp -- the driver shared by sphere_rh54 and sphere_gaussian_vortex: Incompressible2D +
// Incompressible2DRK2 on the sphere (reference: examples/sphere_rh54.cpp:56-380, examples/sphere_gaussian_vortex.cpp:36-300
// minus AMR / remeshing / VTK, which are outside the hot path).
#ifndef LPMX_EXAMPLE_SPHERE_IC2D_HPP
#define LPMX_EXAMPLE_SPHERE_IC2D_HPP
#include <memory>

#include "example_util.hpp"
#include "lpm/lpm.hpp"

namespace example {
using namespace Lpm;

/// Lat0 tracer of the reference examples: initial latitude
struct Lat0 {
  template <typename CV>
  Real operator()(const CV& x) const { return SphereGeometry::latitude(x); }
  std::string name() const { return "lat0"; }
};

template <typename seed_type, typename Vorticity, typename Setup, typename PerStep>
int run_ic2d(const char* example, const Options& opt, Vorticity& vorticity, const CoriolisSphere& coriolis, Setup setup,
             PerStep per_step) {
  const Int depth = opt.get_int("-d", 4);
  const Real tfinal = opt.get_real("-tf", 0.1);
  const Int nsteps = opt.get_int("-n", 10);
  const Real eps = opt.get_real("-eps", 0.0);
  const Real dt = tfinal / nsteps;
  Timer total;
  PolyMeshParameters<seed_type> mesh_params(depth, 1.0, 0, 0);
  auto sphere = std::make_unique<Incompressible2D<seed_type>>(mesh_params, coriolis, eps);
  sphere->init_vorticity(vorticity);
  setup(*sphere, vorticity);
  sphere->init_direct_sums();
  Lat0 lat0;
  sphere->allocate_tracer(lat0);
  sphere->init_tracer(lat0);
  std::printf("%s", sphere->info_string().c_str());
  const auto vel_range = sphere->velocity_active.range(sphere->mesh.n_faces_host());
  const Real cr = vel_range.second * dt / sphere->mesh.appx_mesh_size();
  std::printf("velocity magnitude (min, max) = (%g, %g); approximate Courant number = %g\n", vel_range.first,
              vel_range.second, cr);
  const Real vort0 = sphere->total_vorticity(), ke0 = sphere->total_kinetic_energy(), ens0 = sphere->total_enstrophy();
  auto solver = std::make_unique<Incompressible2DRK2<seed_type>>(dt, *sphere);
  Timer loop;
  Real max_ftle = 0;
  // examples/sphere_rh54.cpp:190-198,255-300: rebuild the particle set every remesh_interval steps (uniform meshes)
  const Int remesh_interval = opt.get_int("-rm", nsteps + 1);
  const bool remesh_direct = opt.get_str("-rs", "indirect") == "direct";
  const gmls::Params gmls_params(opt.get_int("-ro", 4));
  Int rm_counter = 0;
  for (Int t_idx = 0; t_idx < nsteps; ++t_idx) {
    if ((t_idx + 1) % remesh_interval == 0) {
      ++rm_counter;
      auto new_sphere = std::make_unique<Incompressible2D<seed_type>>(mesh_params, coriolis, eps);
      new_sphere->t = sphere->t;
      new_sphere->allocate_tracer(lat0);
      auto remesh = compadre_remesh(*new_sphere, *sphere, gmls_params);
      if (remesh_direct)
        remesh.uniform_direct_remesh();
      else
        remesh.uniform_indirect_remesh(vorticity, coriolis, lat0);
      sphere = std::move(new_sphere);
      solver.reset(new Incompressible2DRK2<seed_type>(dt, *sphere, solver->t_idx));
    }
    sphere->advance_timestep(*solver);
    if constexpr (std::is_same<typename seed_type::faceKind, QuadFace>::value) {
      // examples/sphere_rh54.cpp:308-318 (the reference's FTLE is a static_assert for triangular panels)
      ComputeFTLE<seed_type> ftle(sphere->ftle.view, sphere->mesh.vertices.phys_crds.view, sphere->ref_crds_passive.view,
                                  sphere->mesh.faces.phys_crds.view, sphere->ref_crds_active.view, sphere->mesh.faces.verts,
                                  sphere->mesh.faces.mask, sphere->t - sphere->t_ref);
      ftle.apply(sphere->mesh.n_faces_host());
      max_ftle = get_max_ftle(sphere->ftle.view, sphere->mesh.faces.mask, sphere->mesh.n_faces_host());
      if (max_ftle != ftle.max_ftle && !(std::isnan(max_ftle) || std::isnan(ftle.max_ftle)))
        throw std::runtime_error("device and host max_ftle disagree");
    }
    per_step(*sphere, vorticity);
  }
  std::printf("max_ftle = %.12e; remeshes: %d\n", max_ftle, rm_counter);
  const double loop_s = loop.seconds();
  const Real vort1 = sphere->total_vorticity(), ke1 = sphere->total_kinetic_energy(), ens1 = sphere->total_enstrophy();
  const Index nv = sphere->mesh.n_vertices_host(), nf = sphere->mesh.n_faces_host(), nl = sphere->mesh.faces.n_leaves_host();
  const double inter = 2.0 * ((double)(nv + nf) * nl - nl) * nsteps;
  std::printf("total vorticity %.12e -> %.12e; kinetic energy %.12e -> %.12e; enstrophy %.12e -> %.12e\n", vort0, vort1, ke0,
              ke1, ens0, ens1);
  std::printf("{\"example\": \"%s\", \"seed\": \"%s\", \"depth\": %d, \"steps\": %d, \"dt\": %g, \"t\": %g, \"loop_s\": %.6f, "
              "\"total_s\": %.6f, \"rk2_interactions_per_s\": %.6e, \"gpu_launches\": %ld, \"ke_drift\": %.3e, "
              "\"enstrophy_drift\": %.3e}\n",
              example, seed_type::id_string().c_str(), depth, nsteps, dt, sphere->t, loop_s, total.seconds(), inter / loop_s,
              Engine::launch_count(), std::abs(ke1 - ke0) / ke0, std::abs(ens1 - ens0) / ens0);
  // a Lagrangian particle method conserves enstrophy of the leaves exactly when Omega = 0 and to O(dt^2) otherwise
  return (std::abs(ke1 - ke0) / ke0 < 0.05 && std::isfinite(vort1)) ? 0 : 1;
}
}  // namespace example
#endif
