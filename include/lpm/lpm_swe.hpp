// lpm/lpm_swe.hpp -- SWE<Seed> and SWERK2<Seed, Topo> on the sphere.
//   SWE<Seed>                          src/lpm_swe.hpp:17-150, src/lpm_swe_impl.hpp:59-445
//   SWERK2<Seed, Topo>                 src/lpm_swe_rk2.hpp:15-39, src/lpm_swe_rk2_impl.hpp:16-258
// Differences from the reference, both forced by scope:
//   * SWERK2 takes a SurfaceLaplacian provider where the reference takes gmls::Params: the GMLS Laplacian is
//     Compadre code outside the direct-sum path.  The provider is called for the predictor state (stage 1) and
//     the new state (stage 2), exactly where the reference runs GMLS (rk2_impl.hpp:134-154, :233-252), and by the
//     constructor for the initial state (:56-77).
//   * set_kernel_parameters stores eps on the sphere as well; the reference assigns it only for PlaneGeometry
//     (src/lpm_swe_impl.hpp:98-107), leaving SWE::eps indeterminate on the sphere (SURVEY.md quirk C-iii).
#ifndef LPM_SHIM_SWE_HPP
#define LPM_SHIM_SWE_HPP

#include <functional>

#include "lpm_coriolis.hpp"
#include "lpm_gallery.hpp"
#include "lpm_polymesh2d.hpp"

namespace Lpm {

template <typename SeedType>
class SWE {
 public:
  using geo = typename SeedType::geo;
  using Coriolis = CoriolisSphere;

  ScalarField<VertexField> rel_vort_passive, pot_vort_passive, div_passive, surf_passive, bottom_passive, surf_lap_passive,
      depth_passive, double_dot_passive, stream_fn_passive, potential_passive;
  ScalarField<FaceField> rel_vort_active, pot_vort_active, div_active, surf_active, bottom_active, surf_lap_active,
      depth_active, double_dot_active, stream_fn_active, potential_active, mass_active;
  VectorField<geo, VertexField> velocity_passive;
  VectorField<geo, FaceField> velocity_active;
  PolyMesh2d<SeedType> mesh;
  Coriolis coriolis;
  std::map<std::string, ScalarField<VertexField>> tracer_passive;
  std::map<std::string, ScalarField<FaceField>> tracer_active;
  Real g;
  Real t;
  Real eps;
  Real pse_eps;

  SWE(const PolyMeshParameters<SeedType>& mp, const Coriolis& coriolis)
      : rel_vort_passive("relative_vorticity", mp.nmaxverts), pot_vort_passive("potential_vorticity", mp.nmaxverts),
        div_passive("divergence", mp.nmaxverts), surf_passive("surface_height", mp.nmaxverts),
        bottom_passive("bottom_height", mp.nmaxverts), surf_lap_passive("surface_laplacian", mp.nmaxverts),
        depth_passive("depth", mp.nmaxverts), double_dot_passive("double_dot", mp.nmaxverts),
        stream_fn_passive("stream_function", mp.nmaxverts), potential_passive("potential", mp.nmaxverts),
        rel_vort_active("relative_vorticity", mp.nmaxfaces), pot_vort_active("potential_vorticity", mp.nmaxfaces),
        div_active("divergence", mp.nmaxfaces), surf_active("surface_height", mp.nmaxfaces),
        bottom_active("bottom_height", mp.nmaxfaces), surf_lap_active("surface_laplacian", mp.nmaxfaces),
        depth_active("depth", mp.nmaxfaces), double_dot_active("double_dot", mp.nmaxfaces),
        stream_fn_active("stream_function", mp.nmaxfaces), potential_active("potential", mp.nmaxfaces),
        mass_active("mass", mp.nmaxfaces), velocity_passive("velocity", mp.nmaxverts), velocity_active("velocity", mp.nmaxfaces),
        mesh(mp), coriolis(coriolis), g(1), t(0), eps(0), pse_eps(0) {}

  void set_kernel_parameters(const Real vel_eps, const Real pse) {
    LPM_REQUIRE(vel_eps >= 0);
    eps = vel_eps;
    pse_eps = pse;
  }
  void update_host() {}
  void update_device() {}

  /// bottom, surface, depth = s - b, mass = depth * area (src/lpm_swe_impl.hpp:301-336)
  template <typename BottomType, typename SurfaceType>
  void init_surface(const BottomType& topo, const SurfaceType& sfc) {
    const auto vx = mesh.vertices.phys_crds.view;
    for (Index i = 0; i < mesh.n_vertices_host(); ++i) {
      const Real b = topo(vx.row(i)), s = sfc(vx.row(i));
      bottom_passive.view(i) = b;
      surf_passive.view(i) = s;
      depth_passive.view(i) = s - b;
    }
    const auto fx = mesh.faces.phys_crds.view;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) {
      const Real b = topo(fx.row(i)), s = sfc(fx.row(i));
      bottom_active.view(i) = b;
      surf_active.view(i) = s;
      depth_active.view(i) = s - b;
      mass_active.view(i) = (s - b) * mesh.faces.area(i);
    }
  }

  /// (src/lpm_swe_impl.hpp:338-370)
  template <typename VorticityType>
  void init_vorticity(const VorticityType& vorticity, const bool depth_set = true) {
    const auto vx = mesh.vertices.phys_crds.view;
    for (Index i = 0; i < mesh.n_vertices_host(); ++i) {
      const Real zeta = vorticity(vx.row(i));
      rel_vort_passive.view(i) = zeta;
      if (depth_set) pot_vort_passive.view(i) = (zeta + coriolis.f(vx.row(i))) / depth_passive.view(i);
    }
    const auto fx = mesh.faces.phys_crds.view;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) {
      const Real zeta = vorticity(fx.row(i));
      rel_vort_active.view(i) = zeta;
      if (depth_set) pot_vort_active.view(i) = (zeta + coriolis.f(fx.row(i))) / depth_active.view(i);
    }
  }

  /// (src/lpm_swe_impl.hpp:372-392; evaluated at the Lagrangian coordinates)
  template <typename DivergenceType>
  void init_divergence(const DivergenceType& divergence) {
    const auto vl = mesh.vertices.lag_crds.view;
    for (Index i = 0; i < mesh.n_vertices_host(); ++i) div_passive.view(i) = divergence(vl.row(i));
    const auto fl = mesh.faces.lag_crds.view;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) div_active.view(i) = divergence(fl.row(i));
  }

  void allocate_scalar_tracer(const std::string& name) {
    tracer_passive.emplace(name, ScalarField<VertexField>(name, mesh.params.nmaxverts));
    tracer_active.emplace(name, ScalarField<FaceField>(name, mesh.params.nmaxfaces));
  }

  /// SphereVertexSums / SphereFaceSums on the LAGRANGIAN coordinates (src/lpm_swe_impl.hpp:428-443, quirk C-iv)
  void init_direct_sums(const bool do_velocity = true) {
    lpmx_handle_t h = Engine::get();
    const Index nv = mesh.n_vertices_host(), nf = mesh.n_faces_host();
    Engine::check(lpmx_swe_sphere_sums(h, mesh.vertices.lag_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, nv,
                                       mesh.faces.lag_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, rel_vort_active.view.data(),
                                       div_active.view.data(), mesh.faces.area.data(), mesh.faces.mask.data(), nf, eps, 0,
                                       do_velocity, velocity_passive.view.data(), double_dot_passive.view.data(), nullptr),
                  "SphereVertexSums");
    Engine::check(lpmx_swe_sphere_sums(h, nullptr, LPMX_LAYOUT_RIGHT, 0, nf, mesh.faces.lag_crds.view.data(), LPMX_LAYOUT_RIGHT, 0,
                                       rel_vort_active.view.data(), div_active.view.data(), mesh.faces.area.data(),
                                       mesh.faces.mask.data(), nf, eps, 1, do_velocity, velocity_active.view.data(),
                                       double_dot_active.view.data(), nullptr),
                  "SphereFaceSums");
  }

  template <typename SolverType>
  void advance_timestep(SolverType& solver) {
    solver.advance_timestep_impl();
    t = solver.t_idx * solver.dt;
  }

  std::string info_string(const int tab_level = 0, const bool = false) const {
    std::ostringstream ss;
    ss << "SWE<" << SeedType::id_string() << ">: t = " << t << ", g = " << g << ", eps = " << eps
       << ", Omega = " << coriolis.Omega << "\n" << mesh.info_string("", tab_level + 1);
    return ss.str();
  }
};

/// Host-side surface-Laplacian provider: fills vlaps[nv], flaps[nf] from particle positions (LayoutRight n x 3)
/// and surface heights.  Stands where the reference calls Compadre GMLS.
typedef std::function<void(int stage, Index nv, const Real* vx, const Real* vsurf, Real* vlaps, Index nf, const Real* fx,
                           const Real* fsurf, const unsigned char* fmask, Real* flaps)>
    SurfaceLaplacian;

namespace gmls {
/// gmls::Params (src/lpm_compadre.hpp:23-60): same members, same defaults; Compadre::GMLS::getNP(order, 2) spelled out
struct Params {
  Real eps_multiplier;
  Int samples_order;
  Int manifold_order;
  Real samples_weight_pwr;
  Real manifold_weight_pwr;
  Int ambient_dim;
  Int topo_dim;
  Int min_neighbors;
  Int amr_max;
  Params() : Params(3) {}
  Params(const Int order, const Int dim = 3)
      : eps_multiplier(2.0), samples_order(order), manifold_order(order), samples_weight_pwr(2.0), manifold_weight_pwr(2.0),
        ambient_dim(dim), topo_dim(2), min_neighbors((order + 1) * (order + 2) / 2), amr_max(0) {}
  lpmx_gmls_params_t c_params() const {
    return lpmx_gmls_params_t{eps_multiplier, samples_order, manifold_order, samples_weight_pwr, manifold_weight_pwr,
                              ambient_dim,    topo_dim,      min_neighbors};
  }
  std::string info_string(const int tab_lev = 0) const {
    std::ostringstream ss;
    const std::string tab(tab_lev + 1, '\t');
    ss << std::string(tab_lev, '\t') << "gmls::Params info:\n"
       << tab << "eps_multiplier = " << eps_multiplier << "\n" << tab << "samples_order = " << samples_order << "\n"
       << tab << "manifold_order = " << manifold_order << "\n" << tab << "samples_weight_pwr = " << samples_weight_pwr << "\n"
       << tab << "manifold_weight_pwr = " << manifold_weight_pwr << "\n" << tab << "ambient_dim = " << ambient_dim << "\n"
       << tab << "topo_dim = " << topo_dim << "\n" << tab << "min_neighbors = " << min_neighbors << "\n";
    return ss.str();
  }
};
}  // namespace gmls

template <typename SeedType, typename TopoType = ZeroFunctor>
class SWERK2 {
 public:
  Real dt;
  SWE<SeedType>& swe;
  Int t_idx;
  Real eps;
  SurfaceLaplacian laplacian;
  gmls::Params gmls_params;
  bool use_gmls = false;

  /// The reference's constructor (src/lpm_swe_rk2.hpp:31-32, rk2_impl.hpp:14-78): the surface Laplacian comes from GMLS
  /// with these parameters -- here evaluated on the device (lpmx_gmls_swe_laplacian), not by Compadre on the host.
  SWERK2(const Real dt, SWE<SeedType>& swe, const TopoType&, const gmls::Params& gmls_params)
      : dt(dt), swe(swe), t_idx(0), eps(swe.eps), gmls_params(gmls_params), use_gmls(true) {
    static_assert(std::is_same<TopoType, ZeroFunctor>::value, "the engine implements the flat-bottom (ZeroFunctor) topography");
    provider_.handle = Engine::get();
    provider_.params = gmls_params.c_params();
    // Laplacian of the initial surface: gather -> GMLS -> scatter (rk2_impl.hpp:56-77)
    auto& m = swe.mesh;
    lpmx_handle_t h = Engine::get();
    const Index nv = m.n_vertices_host(), nf = m.n_faces_host();
    int n = 0;
    Engine::check(lpmx_gather_mesh_data(h, 3, LPMX_LAYOUT_RIGHT, nv, m.vertices.phys_crds.view.data(), 0, nf,
                                        m.faces.phys_crds.view.data(), 0, m.faces.mask.data(), nullptr, 0, &n),
                  "GatherMeshData");
    std::vector<Real> gx(3 * (size_t)n), gs(n), gl(n);
    Engine::check(lpmx_gather_mesh_data(h, 3, LPMX_LAYOUT_RIGHT, nv, m.vertices.phys_crds.view.data(), 0, nf,
                                        m.faces.phys_crds.view.data(), 0, m.faces.mask.data(), gx.data(), 0, &n),
                  "GatherMeshData::gather_coordinates");
    Engine::check(lpmx_gather_mesh_data(h, 1, LPMX_LAYOUT_RIGHT, nv, swe.surf_passive.view.data(), 0, nf,
                                        swe.surf_active.view.data(), 0, m.faces.mask.data(), gs.data(), 0, &n),
                  "GatherMeshData::gather_scalar_fields");
    Engine::check(lpmx_gmls_sphere_laplacian(h, &provider_.params, n, gx.data(), LPMX_LAYOUT_RIGHT, 0, gs.data(), gl.data(),
                                             nullptr, nullptr),
                  "sphere_scalar_gmls");
    Engine::check(lpmx_scatter_mesh_data(h, 1, LPMX_LAYOUT_RIGHT, gl.data(), 0, nv, swe.surf_lap_passive.view.data(), 0, nf,
                                         swe.surf_lap_active.view.data(), 0, m.faces.mask.data()),
                  "ScatterMeshData::scatter_fields");
  }

  SWERK2(const Real dt, SWE<SeedType>& swe, const TopoType&, const SurfaceLaplacian& lap)
      : dt(dt), swe(swe), t_idx(0), eps(swe.eps), laplacian(lap) {
    static_assert(std::is_same<TopoType, ZeroFunctor>::value, "the engine implements the flat-bottom (ZeroFunctor) topography");
    // the reference's constructor evaluates the Laplacian of the initial surface (rk2_impl.hpp:56-77)
    if (laplacian) {
      auto& m = swe.mesh;
      laplacian(0, m.n_vertices_host(), m.vertices.phys_crds.view.data(), swe.surf_passive.view.data(),
                swe.surf_lap_passive.view.data(), m.n_faces_host(), m.faces.phys_crds.view.data(), swe.surf_active.view.data(),
                m.faces.mask.data(), swe.surf_lap_active.view.data());
    }
  }

  /// (src/lpm_swe_rk2_impl.hpp:80-258) = lpmx_swe_rk2_step, in place on swe's views
  void advance_timestep_impl() {
    auto& m = swe.mesh;
    lpmx_swe_passive_t P{m.vertices.phys_crds.view.data(), swe.rel_vort_passive.view.data(), swe.div_passive.view.data(),
                         swe.depth_passive.view.data(), swe.surf_passive.view.data(), swe.bottom_passive.view.data(),
                         swe.velocity_passive.view.data(), swe.double_dot_passive.view.data(), swe.surf_lap_passive.view.data()};
    lpmx_swe_active_t A{m.faces.phys_crds.view.data(), swe.rel_vort_active.view.data(), swe.div_active.view.data(),
                        m.faces.area.data(), swe.mass_active.view.data(), swe.depth_active.view.data(),
                        swe.surf_active.view.data(), swe.bottom_active.view.data(), swe.velocity_active.view.data(),
                        swe.double_dot_active.view.data(), swe.surf_lap_active.view.data(), m.faces.mask.data()};
    Engine::check(lpmx_swe_rk2_step(Engine::get(), dt, swe.coriolis.Omega, swe.g, eps, m.n_vertices_host(), &P, m.n_faces_host(),
                                    &A, LPMX_LAYOUT_RIGHT, 0, 0,
                                    use_gmls ? &lpmx_gmls_swe_laplacian : (laplacian ? &SWERK2::trampoline : nullptr),
                                    use_gmls ? (void*)&provider_ : (void*)this, 1),
                  "SWERK2::advance_timestep_impl");
    ++t_idx;
  }

  std::string info_string(const int tab_level = 0) const {
    std::ostringstream ss;
    ss << std::string(tab_level, '\t') << "SWERK2: dt = " << dt << ", t_idx = " << t_idx << ", eps = " << eps << "\n";
    return ss.str();
  }

 private:
  std::vector<Real> hx_, hs_, hl_;
  std::vector<unsigned char> hm_;
  lpmx_gmls_provider_t provider_{};

  // lpmx_swe_laplacian_fn: device SoA -> host LayoutRight -> provider -> device
  static int trampoline(void* user, int stage, void* /*stream*/, int nv, const double* vx, const double* vsurf, double* vlaps,
                        int nf, const double* fx, const double* fsurf, const unsigned char* fmask, double* flaps, long ld) {
    SWERK2* self = static_cast<SWERK2*>(user);
    try {
      lpmx_handle_t h = Engine::get();
      const long nt = (long)nv + nf;
      std::vector<Real> soa(3 * (size_t)nt);
      self->hx_.resize(3 * (size_t)nt), self->hs_.resize(nt), self->hl_.resize(nt), self->hm_.resize(nf);
      // vertices and faces are contiguous in the engine's arrays (faces follow vertices): one copy per row
      for (int k = 0; k < 3; ++k) Engine::check(lpmx_copy(h, soa.data() + (size_t)k * nt, vx + (size_t)k * ld, 8 * nt), "lpmx_copy");
      Engine::check(lpmx_copy(h, self->hs_.data(), vsurf, 8 * nt), "lpmx_copy");
      Engine::check(lpmx_copy(h, self->hm_.data(), fmask, nf), "lpmx_copy");
      if (fx != vx + nv || fsurf != vsurf + nv || flaps != vlaps + nv) return 2;
      for (long i = 0; i < nt; ++i)
        for (int k = 0; k < 3; ++k) self->hx_[3 * i + k] = soa[(size_t)k * nt + i];
      self->laplacian(stage, nv, self->hx_.data(), self->hs_.data(), self->hl_.data(), nf, self->hx_.data() + 3 * (size_t)nv,
                      self->hs_.data() + nv, self->hm_.data(), self->hl_.data() + nv);
      Engine::check(lpmx_copy(h, vlaps, self->hl_.data(), 8 * nt), "lpmx_copy");
      return 0;
    } catch (...) {
      return 1;
    }
  }
};

}  // namespace Lpm
#endif
