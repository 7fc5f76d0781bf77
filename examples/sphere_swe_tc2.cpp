// sphere_swe_tc2 -- Williamson shallow-water test case 2 (steady zonal flow) with SWE + SWERK2 (reference:
// examples/sphere_swe_tc2.cpp:60-251).  The reference evaluates the surface Laplacian with Compadre GMLS; that step is
// outside the direct-sum path, so this driver supplies the Laplacian of the TC2 surface in closed form, which is what the
// reference's own tc2_exact_sol computes as `slap_exact` (:243-244).
//   usage: sphere_swe_tc2 [-s cubed|icos] [-d depth] [-tf tfinal] [-n nsteps] [-eps smoothing]
#include <memory>

#include "example_util.hpp"
#include "lpm/lpm.hpp"

using namespace Lpm;

template <typename seed_type>
int run(const Options& opt) {
  typedef ZeroFunctor topography_type;
  typedef SphereTestCase2InitialSurface init_sfc_type;
  typedef SphereTestCase2Vorticity vorticity_type;
  const Int depth = opt.get_int("-d", 4);
  const Int nsteps = opt.get_int("-n", 10);
  const Real dt = opt.get_real("-tf", 0.05) / nsteps;
  constexpr Real gravity = init_sfc_type::g, omega = 2 * constants::PI, u0 = constants::PI / 6;
  Timer total;
  PolyMeshParameters<seed_type> mesh_params(depth, 1.0, 0);
  CoriolisSphere coriolis(omega);
  auto sphere = std::make_unique<SWE<seed_type>>(mesh_params, coriolis);
  sphere->g = gravity;
  std::printf("Courant number for this problem is appx. %g\n", constants::PI / 6 * dt / sphere->mesh.appx_mesh_size());
  topography_type topo;
  init_sfc_type sfc;
  sphere->init_surface(topo, sfc);
  vorticity_type vorticity;
  sphere->init_vorticity(vorticity, true);
  sphere->set_kernel_parameters(opt.get_real("-eps", 0.0), 0);
  sphere->init_direct_sums(true);

  // slap_exact of examples/sphere_swe_tc2.cpp:243-244
  SurfaceLaplacian slap = [=](int, Index nv, const Real* vx, const Real*, Real* vlaps, Index nf, const Real* fx, const Real*,
                              const unsigned char*, Real* flaps) {
    auto f = [=](const Real* x) {
      const Real sin_sq = square(x[2]), cos_sq = 1 - sin_sq;
      return (square(u0) + 2 * omega * u0) * (2 * sin_sq - cos_sq) / gravity;
    };
    for (Index i = 0; i < nv; ++i) vlaps[i] = f(vx + 3 * i);
    for (Index i = 0; i < nf; ++i) flaps[i] = f(fx + 3 * i);
  };
  // examples/sphere_swe_tc2.cpp:136-139: GMLS of order 4 for the surface Laplacian (here on the device); -lap exact
  // swaps in the closed form above
  const gmls::Params gmls_params(opt.get_int("-gmls", 4), SphereGeometry::ndim);
  auto solver = opt.get_str("-lap", "gmls") == "exact"
                    ? std::make_unique<SWERK2<seed_type, topography_type>>(dt, *sphere, topo, slap)
                    : std::make_unique<SWERK2<seed_type, topography_type>>(dt, *sphere, topo, gmls_params);
  std::printf("%s%s", solver->info_string().c_str(), sphere->info_string().c_str());

  const Index nv = sphere->mesh.n_vertices_host(), nf = sphere->mesh.n_faces_host(), nl = sphere->mesh.faces.n_leaves_host();
  scalar_view_type depth0("depth0", nv), zeta0("zeta0", nf);
  for (Index i = 0; i < nv; ++i) depth0(i) = sphere->depth_passive.view(i);
  for (Index i = 0; i < nf; ++i) zeta0(i) = sphere->rel_vort_active.view(i);
  // -o <root> [-of n]: .vtp frames of the model at t = 0 and after every n-th step (examples/sphere_swe_tc2.cpp:157-168,199-210)
  const std::string vtk_root = opt.has("-o") ? opt.get_str("-o", "") + "_" + seed_type::id_string() + "_" : "";
  const int write_frequency = opt.get_int("-of", 1);
  int frame_counter = 0;
  if (!vtk_root.empty()) vtk_mesh_interface(*sphere).write(vtk_frame_name(vtk_root, frame_counter));
  Timer loop;
  for (int t_idx = 0; t_idx < nsteps; ++t_idx) {
    sphere->advance_timestep(*solver);
    if (!vtk_root.empty() && (t_idx + 1) % write_frequency == 0)
      vtk_mesh_interface(*sphere).write(vtk_frame_name(vtk_root, ++frame_counter));
  }
  const double loop_s = loop.seconds();

  // TC2 is steady: report how far the fields drifted (the reference writes these as VTK error fields)
  scalar_view_type derr("depth_err", nv), zerr("zeta_err", nf), wt_v("unit", nv);
  for (Index i = 0; i < nv; ++i) wt_v(i) = 1;
  ErrNorms depth_err(derr, sphere->depth_passive.view, depth0, wt_v);
  ErrNorms zeta_err(zerr, sphere->rel_vort_active.view, zeta0, sphere->mesh.faces.area, sphere->mesh.faces.mask);
  Real max_div = 0;
  for (Index i = 0; i < nf; ++i)
    if (!sphere->mesh.faces.mask(i)) max_div = std::max(max_div, std::abs(sphere->div_active.view(i)));
  std::printf("t = %g: depth drift %s\n         vorticity drift %s\n         max |divergence| on leaves %.3e\n", sphere->t,
              depth_err.info_string().c_str(), zeta_err.info_string().c_str(), max_div);
  const double inter = 2.0 * ((double)(nv + nf) * nl - nl) * nsteps;
  std::printf("{\"example\": \"sphere_swe_tc2\", \"seed\": \"%s\", \"depth\": %d, \"steps\": %d, \"dt\": %g, \"loop_s\": %.6f, "
              "\"total_s\": %.6f, \"rk2_interactions_per_s\": %.6e, \"gpu_launches\": %ld, \"depth_l2\": %.3e, \"zeta_l2\": %.3e}\n",
              seed_type::id_string().c_str(), depth, nsteps, dt, loop_s, total.seconds(), inter / loop_s, Engine::launch_count(),
              depth_err.l2, zeta_err.l2);
  return (depth_err.l2 < 1e-2 && zeta_err.l2 < 0.1 && std::isfinite(max_div)) ? 0 : 1;
}

int main(int argc, char* argv[]) {
  Options opt(argc, argv);
  try {
    return opt.get_str("-s", "cubed") == "icos" ? run<IcosTriSphereSeed>(opt) : run<CubedSphereSeed>(opt);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "sphere_swe_tc2: %s\n", e.what());
    return 2;
  }
}
