// lpm/lpm_incompressible2d.hpp -- Incompressible2D<Seed> and Incompressible2DRK2<Seed> on the sphere.
//   Incompressible2D<Seed>             src/lpm_incompressible2d.hpp:17-108, src/lpm_incompressible2d_impl.hpp
//   Incompressible2DRK2<Seed>          src/lpm_incompressible2d_rk2.hpp:15-63, _rk2_impl.hpp:75-172
// compadre_remesh(new, old, params) returns the CompadreRemesh hand-off (lpm_compadre_remesh.hpp); adaptive refinement goes
// through mesh.divide_flagged_faces + Refinement (lpm_refinement.hpp); the FTLE field is filled by ComputeFTLE (lpm_ftle.hpp),
// as in the reference's drivers.
#ifndef LPM_SHIM_INCOMPRESSIBLE2D_HPP
#define LPM_SHIM_INCOMPRESSIBLE2D_HPP

#include "lpm_compadre_remesh.hpp"
#include "lpm_coriolis.hpp"
#include "lpm_polymesh2d.hpp"

namespace Lpm {

template <typename SeedType>
class Incompressible2D {
 public:
  using geo = typename SeedType::geo;
  using Coriolis = CoriolisSphere;

  Coords<geo> ref_crds_passive, ref_crds_active;
  ScalarField<VertexField> rel_vort_passive;
  ScalarField<FaceField> rel_vort_active;
  ScalarField<VertexField> abs_vort_passive;
  ScalarField<FaceField> abs_vort_active;
  ScalarField<VertexField> stream_fn_passive;
  ScalarField<FaceField> stream_fn_active;
  VectorField<geo, VertexField> velocity_passive;
  VectorField<geo, FaceField> velocity_active;
  ScalarField<FaceField> ftle;
  std::map<std::string, ScalarField<VertexField>> tracer_passive;
  std::map<std::string, ScalarField<FaceField>> tracer_active;
  PolyMesh2d<SeedType> mesh;
  Coriolis coriolis;
  Real t;
  Real t_ref;
  Real eps;  ///< velocity kernel smoothing parameter

  Incompressible2D(const PolyMeshParameters<SeedType>& mesh_params, const Coriolis& coriolis, const Real velocity_eps)
      : ref_crds_passive(mesh_params.nmaxverts), ref_crds_active(mesh_params.nmaxfaces),
        rel_vort_passive("relative_vorticity", mesh_params.nmaxverts), rel_vort_active("relative_vorticity", mesh_params.nmaxfaces),
        abs_vort_passive("absolute_vorticity", mesh_params.nmaxverts), abs_vort_active("absolute_vorticity", mesh_params.nmaxfaces),
        stream_fn_passive("stream_function", mesh_params.nmaxverts), stream_fn_active("stream_function", mesh_params.nmaxfaces),
        velocity_passive("velocity", mesh_params.nmaxverts), velocity_active("velocity", mesh_params.nmaxfaces),
        ftle("ftle", mesh_params.nmaxfaces), mesh(mesh_params), coriolis(coriolis), t(0), t_ref(0), eps(velocity_eps) {
    // reference coordinates start as the physical coordinates (src/lpm_incompressible2d_impl.hpp:38-39)
    for (Index i = 0; i < mesh.n_vertices_host(); ++i)
      for (int k = 0; k < 3; ++k) ref_crds_passive.view(i, k) = mesh.vertices.phys_crds.view(i, k);
    for (Index i = 0; i < mesh.n_faces_host(); ++i)
      for (int k = 0; k < 3; ++k) ref_crds_active.view(i, k) = mesh.faces.phys_crds.view(i, k);
  }

  void update_host() {}
  void update_device() {}

  /// zeta = vorticity(x), abs = zeta + f(x) (src/lpm_incompressible2d_impl.hpp:139-173)
  template <typename VorticityType>
  void init_vorticity(const VorticityType& vorticity) {
    const auto vx = mesh.vertices.phys_crds.view;
    for (Index i = 0; i < mesh.n_vertices_host(); ++i) {
      const Real zeta = vorticity(vx.row(i));
      rel_vort_passive.view(i) = zeta;
      abs_vort_passive.view(i) = zeta + coriolis.f(vx.row(i));
    }
    const auto fx = mesh.faces.phys_crds.view;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) {
      const Real zeta = vorticity(fx.row(i));
      rel_vort_active.view(i) = zeta;
      abs_vort_active.view(i) = zeta + coriolis.f(fx.row(i));
    }
  }

  void allocate_tracer(const std::string& name) {
    tracer_passive.emplace(name, ScalarField<VertexField>(name, mesh.params.nmaxverts));
    tracer_active.emplace(name, ScalarField<FaceField>(name, mesh.params.nmaxfaces));
  }
  template <typename TracerType>
  void allocate_tracer(const TracerType& tracer, const std::string& tname = std::string()) {
    allocate_tracer(tname.empty() ? tracer.name() : tname);
  }
  /// tracer(x) at the Lagrangian coordinates
  template <typename TracerType>
  void init_tracer(const TracerType& tracer, const std::string& tname = std::string()) {
    const std::string name = tname.empty() ? tracer.name() : tname;
    LPM_REQUIRE_MSG(tracer_passive.count(name) == 1, "tracer not allocated: " + name);
    const auto vl = mesh.vertices.lag_crds.view;
    for (Index i = 0; i < mesh.n_vertices_host(); ++i) tracer_passive.at(name).view(i) = tracer(vl.row(i));
    const auto fl = mesh.faces.lag_crds.view;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) tracer_active.at(name).view(i) = tracer(fl.row(i));
  }
  Int n_tracers() const { return (Int)tracer_passive.size(); }

  /// Incompressible2D{Passive,Active}Sums at the current state (src/lpm_incompressible2d_impl.hpp:235-254)
  void init_direct_sums() {
    lpmx_handle_t h = Engine::get();
    const Index nv = mesh.n_vertices_host(), nf = mesh.n_faces_host();
    Engine::check(lpmx_ic2d_sums(h, mesh.vertices.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, nv, mesh.faces.phys_crds.view.data(),
                                 LPMX_LAYOUT_RIGHT, 0, rel_vort_active.view.data(), mesh.faces.area.data(), mesh.faces.mask.data(),
                                 nf, eps, 0, velocity_passive.view.data(), stream_fn_passive.view.data()),
                  "Incompressible2DPassiveSums");
    Engine::check(lpmx_ic2d_sums(h, nullptr, LPMX_LAYOUT_RIGHT, 0, nf, mesh.faces.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0,
                                 rel_vort_active.view.data(), mesh.faces.area.data(), mesh.faces.mask.data(), nf, eps, 1,
                                 velocity_active.view.data(), stream_fn_active.view.data()),
                  "Incompressible2DActiveSums");
  }

  /// solver.advance_timestep_impl(); t = t_idx * dt (src/lpm_incompressible2d_impl.hpp:273-278)
  template <typename SolverType>
  void advance_timestep(SolverType& solver) {
    solver.advance_timestep_impl();
    t = solver.t_idx * solver.dt;
  }

  /// conserved totals over the leaves (src/lpm_incompressible2d_impl.hpp:91-137)
  Real total_vorticity() const {
    Real s = 0;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) s += (mesh.faces.mask(i) ? 0 : rel_vort_active.view(i) * mesh.faces.area(i));
    return s;
  }
  Real total_enstrophy() const {
    Real s = 0;
    for (Index i = 0; i < mesh.n_faces_host(); ++i)
      s += (mesh.faces.mask(i) ? 0 : square(rel_vort_active.view(i)) * mesh.faces.area(i));
    return 0.5 * s;
  }
  Real total_kinetic_energy() const {
    Real s = 0;
    for (Index i = 0; i < mesh.n_faces_host(); ++i)
      if (!mesh.faces.mask(i)) s += geo::norm2(velocity_active.view.row(i)) * mesh.faces.area(i);
    return 0.5 * s;
  }

  std::string info_string(const int tab_level = 0) const {
    std::ostringstream ss;
    ss << "Incompressible2D<" << SeedType::id_string() << ">: t = " << t << ", eps = " << eps << ", Omega = " << coriolis.Omega
       << ", " << n_tracers() << " tracers\n" << mesh.info_string("", tab_level + 1);
    return ss.str();
  }
};

/// compadre_remesh(new_ic2d, old_ic2d, gmls_params) (src/lpm_incompressible2d_impl.hpp:389-456): the same field maps --
/// relative/absolute vorticity, stream function, every tracer; velocity as the vector field -- and t_ref / ref_crds reset.
template <typename SeedType>
CompadreRemesh<SeedType> compadre_remesh(Incompressible2D<SeedType>& new_ic2d, const Incompressible2D<SeedType>& old_ic2d,
                                         const gmls::Params& gmls_params) {
  typename CompadreRemesh<SeedType>::vert_scalar_field_map ps_old, ps_new;
  typename CompadreRemesh<SeedType>::face_scalar_field_map as_old, as_new;
  typename CompadreRemesh<SeedType>::vert_vector_field_map pv_old, pv_new;
  typename CompadreRemesh<SeedType>::face_vector_field_map av_old, av_new;
  for (Index i = 0; i < new_ic2d.mesh.n_vertices_host(); ++i)
    for (int k = 0; k < 3; ++k) new_ic2d.ref_crds_passive.view(i, k) = new_ic2d.mesh.vertices.phys_crds.view(i, k);
  for (Index i = 0; i < new_ic2d.mesh.n_faces_host(); ++i)
    for (int k = 0; k < 3; ++k) new_ic2d.ref_crds_active.view(i, k) = new_ic2d.mesh.faces.phys_crds.view(i, k);
  new_ic2d.t_ref = old_ic2d.t;
  ps_old.emplace("relative_vorticity", old_ic2d.rel_vort_passive), ps_new.emplace("relative_vorticity", new_ic2d.rel_vort_passive);
  ps_old.emplace("absolute_vorticity", old_ic2d.abs_vort_passive), ps_new.emplace("absolute_vorticity", new_ic2d.abs_vort_passive);
  ps_old.emplace("stream_function", old_ic2d.stream_fn_passive), ps_new.emplace("stream_function", new_ic2d.stream_fn_passive);
  as_old.emplace("relative_vorticity", old_ic2d.rel_vort_active), as_new.emplace("relative_vorticity", new_ic2d.rel_vort_active);
  as_old.emplace("absolute_vorticity", old_ic2d.abs_vort_active), as_new.emplace("absolute_vorticity", new_ic2d.abs_vort_active);
  as_old.emplace("stream_function", old_ic2d.stream_fn_active), as_new.emplace("stream_function", new_ic2d.stream_fn_active);
  for (const auto& t : old_ic2d.tracer_passive) ps_old.emplace(t.first, t.second);
  for (const auto& t : new_ic2d.tracer_passive) ps_new.emplace(t.first, t.second);
  for (const auto& t : old_ic2d.tracer_active) as_old.emplace(t.first, t.second);
  for (const auto& t : new_ic2d.tracer_active) as_new.emplace(t.first, t.second);
  pv_old.emplace("velocity", old_ic2d.velocity_passive), pv_new.emplace("velocity", new_ic2d.velocity_passive);
  av_old.emplace("velocity", old_ic2d.velocity_active), av_new.emplace("velocity", new_ic2d.velocity_active);
  return CompadreRemesh<SeedType>(new_ic2d.mesh, ps_new, as_new, pv_new, av_new, old_ic2d.mesh, ps_old, as_old, pv_old, av_old,
                                  gmls_params);
}

template <typename SeedType>
class Incompressible2DRK2 {
 public:
  using geo = typename SeedType::geo;
  Real dt;
  Incompressible2D<SeedType>& ic2d;
  Int t_idx;
  Index n_passive, n_active;
  Real eps;

  Incompressible2DRK2(const Real dt, Incompressible2D<SeedType>& ic2d, const Index t_idx = 0)
      : dt(dt), ic2d(ic2d), t_idx(t_idx), n_passive(ic2d.mesh.n_vertices_host()), n_active(ic2d.mesh.n_faces_host()),
        eps(ic2d.eps) {}

  /// Heun step (src/lpm_incompressible2d_rk2_impl.hpp:75-172) = lpmx_ic2d_rk2_step, in place on ic2d's views
  void advance_timestep_impl() {
    auto& m = ic2d.mesh;
    Engine::check(lpmx_ic2d_rk2_step(Engine::get(), dt, ic2d.coriolis.Omega, eps, n_passive, m.vertices.phys_crds.view.data(),
                                     ic2d.rel_vort_passive.view.data(), ic2d.velocity_passive.view.data(),
                                     ic2d.stream_fn_passive.view.data(), n_active, m.faces.phys_crds.view.data(),
                                     ic2d.rel_vort_active.view.data(), ic2d.velocity_active.view.data(),
                                     ic2d.stream_fn_active.view.data(), m.faces.area.data(), m.faces.mask.data(),
                                     LPMX_LAYOUT_RIGHT, 0, 0, 1),
                  "Incompressible2DRK2::advance_timestep_impl");
    ++t_idx;
  }

  std::string info_string(const int tab_level = 0) const {
    std::ostringstream ss;
    ss << std::string(tab_level, '\t') << "Incompressible2DRK2: dt = " << dt << ", t_idx = " << t_idx << ", eps = " << eps << "\n";
    return ss.str();
  }
};

}  // namespace Lpm
#endif
