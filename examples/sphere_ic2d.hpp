// sphere_ic2d.hpp -- the driver shared by sphere_rh54 and sphere_gaussian_vortex: Incompressible2D +
// Incompressible2DRK2 on the sphere (reference: examples/sphere_rh54.cpp:56-380, examples/sphere_gaussian_vortex.cpp:36-300):
// adaptive refinement at start-up (-ab / -al / -amr, -c), remeshing on an interval or on the FTLE (-rm, -rs, -ro, -rt, -ftle),
// uniform or adaptive; -o <root> -of <n> writes .vtp frames of the model (vtk_mesh_interface).
#ifndef LPMX_EXAMPLE_SPHERE_IC2D_HPP
#define LPMX_EXAMPLE_SPHERE_IC2D_HPP
#include <limits>
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
  // examples/sphere_gaussian_vortex.cpp:58-76: -amr n sets both the memory buffer and the refinement limit
  Int amr_buffer = opt.get_int("-ab", 0), amr_limit = opt.get_int("-al", 0);
  if (opt.get_int("-amr", -1) > 0) amr_buffer = amr_limit = opt.get_int("-amr", -1);
  const bool amr = (amr_buffer > 0 && amr_limit > 0);
  Real max_circ_tol = opt.get_real("-c", std::numeric_limits<Real>::max());
  Logger logger(example, opt.has("-v") ? Log::debug : Log::info);
  PolyMeshParameters<seed_type> mesh_params(depth, 1.0, amr_buffer, amr_limit);
  auto sphere = std::make_unique<Incompressible2D<seed_type>>(mesh_params, coriolis, eps);
  sphere->init_vorticity(vorticity);
  setup(*sphere, vorticity);
  if (amr) {
    // :89-118 -- flag on the circulation |zeta| A, relative tolerance fixed on the uniform mesh; each pass looks at the
    // faces the previous one added, divides, and re-evaluates the initial vorticity on the new particles
    Refinement<seed_type> refiner(sphere->mesh);
    ScalarIntegralFlag max_circulation_flag(refiner.flags, sphere->rel_vort_active.view, sphere->mesh.faces.area,
                                            sphere->mesh.faces.mask, sphere->mesh.n_faces_host(), max_circ_tol);
    max_circulation_flag.set_tol_from_relative_value();
    max_circ_tol = max_circulation_flag.tol;
    logger.info("amr is enabled with limit {}, max_circ_tol = {}", amr_limit, max_circ_tol);
    Index face_start_idx = 0;
    for (int i = 0; i < amr_limit; ++i) {
      const Index face_end_idx = sphere->mesh.n_faces_host();
      refiner.iterate(face_start_idx, face_end_idx, max_circulation_flag);
      logger.info("amr iteration {}: initial circulation refinement count = {}", i, refiner.count[0]);
      sphere->mesh.divide_flagged_faces(refiner.flags, logger);
      sphere->update_device();
      sphere->init_vorticity(vorticity);
      face_start_idx = face_end_idx;
    }
    ko::deep_copy(sphere->ref_crds_passive.view, sphere->mesh.vertices.lag_crds.view);
    ko::deep_copy(sphere->ref_crds_active.view, sphere->mesh.faces.lag_crds.view);
  } else {
    logger.info("amr is not enabled; using uniform meshes.");
  }
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
  // -o <root> [-of n]: a .vtp frame of the whole model at t = 0 and after every n-th step (reference: LPM_USE_VTK blocks)
  const std::string vtk_root = opt.has("-o") ? opt.get_str("-o", "") + "_" + seed_type::id_string() + std::to_string(depth) + "_" : "";
  const Int write_frequency = opt.get_int("-of", 1);
  int frame_counter = 0;
  if (!vtk_root.empty()) vtk_mesh_interface(*sphere).write(vtk_frame_name(vtk_root, frame_counter));
  Timer loop;
  Real max_ftle = 0;
  // examples/sphere_rh54.cpp:190-198,255-300: rebuild the particle set every remesh_interval steps (uniform meshes)
  const Int remesh_interval = opt.get_int("-rm", nsteps + 1);
  const bool remesh_direct = opt.get_str("-rs", "indirect") == "direct";
  const gmls::Params gmls_params(opt.get_int("-ro", 4));
  const bool use_ftle = opt.get_str("-rt", "interval") == "ftle";  // :147,205-215
  const Real ftle_tol = opt.get_real("-ftle", 2.0);
  Int rm_counter = 0;
  for (Int t_idx = 0; t_idx < nsteps; ++t_idx) {
    const bool ftle_trigger = (use_ftle && max_ftle > ftle_tol);
    const bool interval_trigger = ((t_idx + 1) % remesh_interval == 0);
    if (ftle_trigger || interval_trigger) {
      ++rm_counter;
      if (ftle_trigger) logger.info("remesh {} triggered by ftle", rm_counter);
      auto new_sphere = std::make_unique<Incompressible2D<seed_type>>(mesh_params, coriolis, eps);
      new_sphere->t = sphere->t;
      new_sphere->allocate_tracer(lat0);
      auto remesh = compadre_remesh(*new_sphere, *sphere, gmls_params);
      if (amr) {
        // :224-235 -- the new mesh is refined where the interpolated circulation exceeds the start-up tolerance
        Refinement<seed_type> refiner(new_sphere->mesh);
        ScalarIntegralFlag max_circulation_flag(refiner.flags, new_sphere->rel_vort_active.view, new_sphere->mesh.faces.area,
                                                new_sphere->mesh.faces.mask, new_sphere->mesh.n_faces_host(), max_circ_tol);
        if (remesh_direct)
          remesh.adaptive_direct_remesh(refiner, max_circulation_flag);
        else
          remesh.adaptive_indirect_remesh(refiner, max_circulation_flag, vorticity, coriolis, lat0);
        // compadre_remesh() resets the reference coordinates BEFORE the adaptive passes add particles
        // (src/lpm_incompressible2d_impl.hpp:394-397); the reference leaves those of the added particles at zero, which makes
        // their FTLE NaN.  Deviation, flagged: they are set here, after the passes.
        ko::deep_copy(new_sphere->ref_crds_passive.view, new_sphere->mesh.vertices.phys_crds.view);
        ko::deep_copy(new_sphere->ref_crds_active.view, new_sphere->mesh.faces.phys_crds.view);
      } else if (remesh_direct) {
        remesh.uniform_direct_remesh();
      } else {
        remesh.uniform_indirect_remesh(vorticity, coriolis, lat0);
      }
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
    if (!vtk_root.empty() && (t_idx + 1) % write_frequency == 0)
      vtk_mesh_interface(*sphere).write(vtk_frame_name(vtk_root, ++frame_counter));
  }
  std::printf("max_ftle = %.12e; remeshes: %d\n", max_ftle, rm_counter);
  const double loop_s = loop.seconds();
  const Real vort1 = sphere->total_vorticity(), ke1 = sphere->total_kinetic_energy(), ens1 = sphere->total_enstrophy();
  const Index nv = sphere->mesh.n_vertices_host(), nf = sphere->mesh.n_faces_host(), nl = sphere->mesh.faces.n_leaves_host();
  const double inter = 2.0 * ((double)(nv + nf) * nl - nl) * nsteps;
  std::printf("total vorticity %.12e -> %.12e; kinetic energy %.12e -> %.12e; enstrophy %.12e -> %.12e\n", vort0, vort1, ke0,
              ke1, ens0, ens1);
  Index max_level = 0;
  for (Index i = 0; i < nf; ++i) max_level = std::max(max_level, sphere->mesh.faces.level(i));
  std::printf("{\"example\": \"%s\", \"seed\": \"%s\", \"depth\": %d, \"steps\": %d, \"dt\": %g, \"t\": %g, \"loop_s\": %.6f, "
              "\"total_s\": %.6f, \"rk2_interactions_per_s\": %.6e, \"gpu_launches\": %ld, \"ke_drift\": %.3e, "
              "\"enstrophy_drift\": %.3e, \"n_verts\": %d, \"n_faces\": %d, \"n_leaves\": %d, \"max_level\": %d}\n",
              example, seed_type::id_string().c_str(), depth, nsteps, dt, sphere->t, loop_s, total.seconds(), inter / loop_s,
              Engine::launch_count(), std::abs(ke1 - ke0) / ke0, std::abs(ens1 - ens0) / ens0, nv, nf, nl, max_level);
  // a Lagrangian particle method conserves enstrophy of the leaves exactly when Omega = 0 and to O(dt^2) otherwise
  return (std::abs(ke1 - ke0) / ke0 < 0.05 && std::isfinite(vort1)) ? 0 : 1;
}
}  // namespace example
#endif
