// plane_gravity_wave -- examples/plane_gravity_wave.cpp of the reference against the shim: a Gaussian surface perturbation
// over a Gaussian mountain on a free-boundary planar mesh, SWE<QuadRectSeed> + SWERK4 (the planar 9-tuple direct sums with
// the PSE Laplacian, 4 evaluations per step).  Options as in the reference: -tf -n -d -f -b -r -ab -al -eps -pse; -s
// quad|tri chooses the seed (the reference hard-codes QuadRectSeed and mentions TriHexSeed in a comment, :30).
#include <cstdio>

#include "example_util.hpp"
#include "lpm/lpm.hpp"

using namespace Lpm;

template <typename seed_type>
int run(const Options& opt) {
  using topography_type = PlanarGaussianMountain;
  using init_sfc_type = PlanarGaussianSurfacePerturbation;
  using coriolis_type = CoriolisBetaPlane;
  using pse_type = pse::BivariateOrder8;
  (void)sizeof(pse_type);
  const int nsteps = opt.get_int("-n", 5);
  const Real dt = opt.get_real("-tf", 0.5) / nsteps;
  Timer total;
  // the reference passes (depth, radius, amr_limit): the third argument is the AMR buffer (:121-124)
  PolyMeshParameters<seed_type> mesh_params(opt.get_int("-d", 4), opt.get_real("-r", 6.0), opt.get_int("-al", 0));
  coriolis_type coriolis(opt.get_real("-f", 0.0), opt.get_real("-b", 0.0));
  auto plane = std::make_unique<SWE<seed_type>>(mesh_params, coriolis);
  topography_type topo;
  init_sfc_type sfc;
  plane->init_surface(topo, sfc);
  constexpr bool do_velocity = true;
  plane->set_kernel_parameters(opt.get_real("-eps", 0.0),
                               pse::PSEKernel<PlaneGeometry>::get_epsilon(plane->mesh.appx_mesh_size(), opt.get_real("-pse", 11.0 / 20)));
  plane->init_direct_sums(do_velocity);
  std::printf("%s", plane->info_string().c_str());
  const Real mass0 = plane->total_mass();
  const auto s0 = plane->surf_active.range(plane->mesh.n_faces_host());
  auto solver = std::make_unique<SWERK4<seed_type, topography_type>>(dt, *plane, topo);
  std::printf("%s", solver->info_string().c_str());
  // -o <root> [-of n]: .vtp frames at t = 0 and after every n-th step (:139-150,160-170)
  const std::string vtk_root =
      opt.has("-o") ? opt.get_str("-o", "") + "_" + seed_type::id_string() + std::to_string(mesh_params.init_depth) + "_" : "";
  const int write_frequency = opt.get_int("-of", 1);
  int frame_counter = 0;
  if (!vtk_root.empty()) vtk_mesh_interface(*plane).write(vtk_frame_name(vtk_root, frame_counter));
  Timer loop;
  for (int t_idx = 0; t_idx < nsteps; ++t_idx) {
    plane->advance_timestep(*solver);
    if (!vtk_root.empty() && (t_idx + 1) % write_frequency == 0)
      vtk_mesh_interface(*plane).write(vtk_frame_name(vtk_root, ++frame_counter));
  }
  const double loop_s = loop.seconds();
  const Real mass1 = plane->total_mass();
  const Index nv = plane->mesh.n_vertices_host(), nf = plane->mesh.n_faces_host(), nl = plane->mesh.faces.n_leaves_host();
  // the surface stays between the flat level and the initial crest while the wave spreads; depth stays positive
  Real smin = 1e300, smax = -1e300, hmin = 1e300, umax = 0;
  bool finite = true;
  for (Index i = 0; i < nf; ++i) {
    if (plane->mesh.faces.mask(i)) continue;
    const Real s = plane->surf_active.view(i), h = plane->depth_active.view(i);
    smin = std::min(smin, s), smax = std::max(smax, s), hmin = std::min(hmin, h);
    umax = std::max(umax, PlaneGeometry::mag(plane->velocity_active.view.row(i)));
    finite = finite && std::isfinite(s) && std::isfinite(h);
  }
  const double inter = 4.0 * ((double)(nv + nf) * nl - nl) * nsteps;
  std::printf("surface range (%.6f, %.6f) -> (%.6f, %.6f); min depth %.6f; max speed %.6f; total mass %.12e -> %.12e\n", s0.first,
              s0.second, smin, smax, hmin, umax, mass0, mass1);
  std::printf("{\"example\": \"plane_gravity_wave\", \"seed\": \"%s\", \"depth\": %d, \"steps\": %d, \"dt\": %g, \"t\": %g, "
              "\"loop_s\": %.6f, \"total_s\": %.6f, \"rk4_interactions_per_s\": %.6e, \"gpu_launches\": %ld, \"n_verts\": %d, "
              "\"n_faces\": %d, \"n_leaves\": %d, \"pse_eps\": %.6f, \"surf_min\": %.8f, \"surf_max\": %.8f, \"surf_max0\": %.8f, "
              "\"min_depth\": %.8f, \"max_speed\": %.8f, \"mass_drift\": %.3e}\n",
              seed_type::id_string().c_str(), mesh_params.init_depth, nsteps, dt, plane->t, loop_s, total.seconds(), inter / loop_s,
              Engine::launch_count(), nv, nf, nl, plane->pse_eps, smin, smax, s0.second, hmin, umax,
              std::abs(mass1 - mass0) / mass0);
  return (finite && hmin > 0 && smax <= s0.second + 1e-3) ? 0 : 1;
}

int main(int argc, char** argv) {
  const Options opt(argc, argv);
  if (opt.has("help")) {
    std::printf("plane_gravity_wave [-s quad|tri] [-d depth] [-r radius] [-tf tfinal] [-n nsteps] [-f f0] [-b beta] [-eps eps] "
                "[-pse power] [-al amr]\n");
    return 0;
  }
  try {
    if (opt.get_str("-s", "quad") == "tri") return run<TriHexSeed>(opt);
    return run<QuadRectSeed>(opt);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "plane_gravity_wave: %s\n", e.what());
    return std::string(e.what()).find("lpmx_create failed") != std::string::npos ? 2 : 3;
  }
}
