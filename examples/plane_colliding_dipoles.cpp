// plane_colliding_dipoles -- examples/plane_colliding_dipoles.cpp of the reference against the shim: two Lamb dipoles on a
// free-boundary planar mesh, Incompressible2D<QuadRectSeed> + Incompressible2DRK2, adaptive refinement at start-up with the
// circulation and vorticity-variation flags (:94-139), FTLE every step (:269-279).  Options as in the reference: -tf -n -d -r
// -f -b -eps -ab -al -amr -c (max circulation tol) -zv (vorticity variation tol) -fv (flow map variation tol).
// Not provided: the reference's planar remesh is bivar_remesh (BIVAR scattered-data interpolation, src/mesh/lpm_bivar_remesh*),
// a third-party algorithm outside this path; -rm is refused.
#include <cstdio>
#include <limits>

#include "example_util.hpp"
#include "lpm/lpm.hpp"

using namespace Lpm;

int main(int argc, char** argv) {
  const Options opt(argc, argv);
  if (opt.has("help")) {
    std::printf("plane_colliding_dipoles [-d depth] [-r radius] [-tf tfinal] [-n nsteps] [-eps eps] [-amr n] [-c tol] [-zv tol] [-fv tol]\n");
    return 0;
  }
  try {
    using seed_type = QuadRectSeed;
    using coriolis_type = CoriolisBetaPlane;
    using vorticity_type = CollidingDipolePairPlane;
    LPM_REQUIRE_MSG(!opt.has("-rm"), "planar remeshing (bivar_remesh) is not part of this engine");
    const int nsteps = opt.get_int("-n", 5);
    const Real dt = opt.get_real("-tf", 0.1) / nsteps;
    Int amr_buffer = opt.get_int("-ab", 0), amr_limit = opt.get_int("-al", 0);
    if (opt.get_int("-amr", -1) > 0) amr_buffer = amr_limit = opt.get_int("-amr", -1);
    Timer total;
    Logger logger("plane_colliding_dipoles", opt.has("-v") ? Log::debug : Log::info);
    PolyMeshParameters<seed_type> mesh_params(opt.get_int("-d", 4), opt.get_real("-r", 6.0), amr_buffer, amr_limit);
    const Real big = std::numeric_limits<Real>::max();
    Real max_circ_tol = opt.get_real("-c", big), flow_map_var_tol = opt.get_real("-fv", big), zeta_var_tol = opt.get_real("-zv", big);
    const bool amr = (mesh_params.amr_limit > 0 && (max_circ_tol < 0.5 * big || flow_map_var_tol < 0.5 * big || zeta_var_tol < 0.5 * big));
    coriolis_type coriolis(opt.get_real("-f", 0.0), opt.get_real("-b", 0.0));
    const Real epsilon = opt.get_real("-eps", 0.0);
    auto plane = std::make_unique<Incompressible2D<seed_type>>(mesh_params, coriolis, epsilon);
    vorticity_type vorticity;
    plane->init_vorticity(vorticity);
    if (amr) {
      Refinement<seed_type> refiner(plane->mesh);
      ScalarIntegralFlag max_circulation_flag(refiner.flags, plane->rel_vort_active.view, plane->mesh.faces.area,
                                              plane->mesh.faces.mask, plane->mesh.n_faces_host(), max_circ_tol);
      FlowMapVariationFlag<seed_type> flow_map_variation_flag(refiner.flags, plane->mesh, flow_map_var_tol);
      ScalarVariationFlag zeta_var_flag(refiner.flags, plane->rel_vort_active.view, plane->rel_vort_passive.view,
                                        plane->mesh.faces.verts, plane->mesh.faces.mask, plane->mesh.n_faces_host(), zeta_var_tol);
      max_circulation_flag.set_tol_from_relative_value();
      max_circ_tol = max_circulation_flag.tol;
      flow_map_variation_flag.set_tol_from_relative_value();
      flow_map_var_tol = flow_map_variation_flag.tol;
      zeta_var_flag.set_tol_from_relative_value();
      zeta_var_tol = zeta_var_flag.tol;
      logger.info("amr is enabled with limit {}, max_circ_tol = {}, flow_map_var_tol = {}, zeta_var_tol = {}", mesh_params.amr_limit,
                  max_circ_tol, flow_map_var_tol, zeta_var_tol);
      Index face_start_idx = 0;
      for (int i = 0; i < amr_limit; ++i) {
        const Index face_end_idx = plane->mesh.n_faces_host();
        refiner.iterate(face_start_idx, face_end_idx, max_circulation_flag, zeta_var_flag);
        logger.info("amr iteration {}: initial circulation refinement count = {}", i, refiner.count[0]);
        logger.info("amr iteration {}: vorticity variation refinement count = {}", i, refiner.count[1]);
        plane->mesh.divide_flagged_faces(refiner.flags, logger);
        plane->update_device();
        plane->init_vorticity(vorticity);
        face_start_idx = face_end_idx;
      }
      // the reference leaves ref_crds of the added particles at zero here; set them (see sphere_ic2d.hpp)
      ko::deep_copy(plane->ref_crds_passive.view, plane->mesh.vertices.phys_crds.view);
      ko::deep_copy(plane->ref_crds_active.view, plane->mesh.faces.phys_crds.view);
    } else {
      logger.info("amr is not enabled; using uniform meshes.");
    }
    plane->init_direct_sums();
    const auto vel_range = plane->velocity_active.range(plane->mesh.n_faces_host());
    const Real cr = vel_range.second * dt / plane->mesh.appx_min_mesh_size();
    std::printf("%s", plane->info_string().c_str());
    logger.info("velocity magnitude (min, max) = ({}, {}); approximate Courant number = {}", vel_range.first, vel_range.second, cr);
    if (cr > 0.5) logger.warn("Courant number {} may be too high.", cr);
    const Real vort0 = plane->total_vorticity(), ke0 = plane->total_kinetic_energy(), ens0 = plane->total_enstrophy();
    auto solver = std::make_unique<Incompressible2DRK2<seed_type>>(dt, *plane);
    Real max_ftle = 0;
    const Real tref = 0;
    const std::string vtk_root =
        opt.has("-o") ? opt.get_str("-o", "") + "_" + seed_type::id_string() + std::to_string(mesh_params.init_depth) + "_" : "";
    const int write_frequency = opt.get_int("-of", 1);
    int frame_counter = 0;
    if (!vtk_root.empty()) vtk_mesh_interface(*plane).write(vtk_frame_name(vtk_root, frame_counter));
    Timer loop;
    for (int t_idx = 0; t_idx < nsteps; ++t_idx) {
      plane->advance_timestep(*solver);
      ComputeFTLE<seed_type> ftle(plane->ftle.view, plane->mesh.vertices.phys_crds.view, plane->ref_crds_passive.view,
                                  plane->mesh.faces.phys_crds.view, plane->ref_crds_active.view, plane->mesh.faces.verts,
                                  plane->mesh.faces.mask, plane->t - tref);
      ftle.apply(plane->mesh.n_faces_host());
      max_ftle = get_max_ftle(plane->ftle.view, plane->mesh.faces.mask, plane->mesh.n_faces_host());
      if (!vtk_root.empty() && (t_idx + 1) % write_frequency == 0)
        vtk_mesh_interface(*plane).write(vtk_frame_name(vtk_root, ++frame_counter));
    }
    const double loop_s = loop.seconds();
    const Real vort1 = plane->total_vorticity(), ke1 = plane->total_kinetic_energy(), ens1 = plane->total_enstrophy();
    const Index nv = plane->mesh.n_vertices_host(), nf = plane->mesh.n_faces_host(), nl = plane->mesh.faces.n_leaves_host();
    Index max_level = 0;
    for (Index i = 0; i < nf; ++i) max_level = std::max(max_level, plane->mesh.faces.level(i));
    const double inter = 2.0 * ((double)(nv + nf) * nl - nl) * nsteps;
    std::printf("total vorticity %.12e -> %.12e; kinetic energy %.12e -> %.12e; enstrophy %.12e -> %.12e; max_ftle %.6e\n", vort0,
                vort1, ke0, ke1, ens0, ens1, max_ftle);
    std::printf("{\"example\": \"plane_colliding_dipoles\", \"seed\": \"%s\", \"depth\": %d, \"steps\": %d, \"dt\": %g, \"t\": %g, "
                "\"loop_s\": %.6f, \"total_s\": %.6f, \"rk2_interactions_per_s\": %.6e, \"gpu_launches\": %ld, \"n_verts\": %d, "
                "\"n_faces\": %d, \"n_leaves\": %d, \"max_level\": %d, \"total_vorticity\": %.3e, \"ke_drift\": %.3e, "
                "\"enstrophy_drift\": %.3e, \"max_ftle\": %.6e}\n",
                seed_type::id_string().c_str(), mesh_params.init_depth, nsteps, dt, plane->t, loop_s, total.seconds(), inter / loop_s,
                Engine::launch_count(), nv, nf, nl, max_level, vort1, std::abs(ke1 - ke0) / ke0, std::abs(ens1 - ens0) / ens0,
                max_ftle);
    return (std::isfinite(ke1) && std::abs(ke1 - ke0) / ke0 < 0.05) ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "plane_colliding_dipoles: %s\n", e.what());
    return std::string(e.what()).find("lpmx_create failed") != std::string::npos ? 2 : 3;
  }
}
