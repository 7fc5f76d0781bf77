// lpm/lpm_plane.hpp -- the planar problems: PlaneGeometry, the two planar seeds, CoriolisBetaPlane, the planar gallery,
// Incompressible2D / Incompressible2DRK2 and SWE / SWERK4 for planar seeds, over the planar entry points of the C ABI.
//   PlaneGeometry                         src/lpm_geometry.hpp:24-166
//   QuadRectSeed, TriHexSeed              src/mesh/lpm_mesh_seed.hpp:46-82
//   CoriolisBetaPlane                     src/lpm_coriolis.hpp:93-148
//   PlanarGaussianMountain, PlanarGaussianSurfacePerturbation   src/lpm_surface_gallery.hpp:41-90
//   CollidingDipolePairPlane, lamb_dipole_vorticity             src/lpm_vorticity_gallery.hpp:192-259
//   pse::PSEKernel<Geo>::get_epsilon, pse::BivariateOrder8      src/lpm_pse.hpp:15-74
//   Incompressible2D<Seed> (PlaneGeometry branches)             src/lpm_incompressible2d_impl.hpp
//   SWE<Seed> (PlaneGeometry branches)                          src/lpm_swe.hpp:29-150, src/lpm_swe_impl.hpp:99-107,301-336,401-427
//   SWERK4<Seed, Topo>                    src/lpm_swe_rk4.hpp:15-100, src/lpm_swe_rk4_impl.hpp:193-445
// Drivers: examples/plane_gravity_wave.cpp, examples/plane_colliding_dipoles.cpp.
// SWE<QuadRectSeed> etc. are explicit specialisations of the class templates declared in lpm_swe.hpp /
// lpm_incompressible2d.hpp, so `SWE<seed_type>` reads the same for either geometry, as in the reference.
#ifndef LPM_SHIM_PLANE_HPP
#define LPM_SHIM_PLANE_HPP

#include <cmath>

#include "lpm_incompressible2d.hpp"
#include "lpm_swe.hpp"

namespace Lpm {

typedef View2<Real, 2> vec2_view_type;

struct PlaneGeometry {
  static constexpr Int ndim = 2;
  typedef vec2_view_type crd_view_type;
  typedef vec2_view_type vec_view_type;
  static std::string id_string() { return "PlaneGeometry"; }
  template <typename A, typename B>
  static Real dot(const A& a, const B& b) { return a[0] * b[0] + a[1] * b[1]; }
  template <typename A>
  static Real norm2(const A& a) { return a[0] * a[0] + a[1] * a[1]; }
  template <typename A>
  static Real mag(const A& a) { return std::sqrt(norm2(a)); }
  template <typename A, typename B>
  static Real distance(const A& a, const B& b) {
    const Real d[2] = {b[0] - a[0], b[1] - a[1]};
    return mag(d);
  }
};

struct QuadRectSeed {
  static constexpr Int nverts = 9, nfaces = 4, nedges = 12, nfaceverts = 4, vertex_degree = 4;
  static constexpr int lpmx_id = LPMX_SEED_QUAD_RECT;
  typedef PlaneGeometry geo;
  typedef QuadFace faceKind;
  static std::string filename() { return "quadRectSeed.dat"; }
  static std::string id_string() { return "quad_rect"; }
};

struct TriHexSeed {
  static constexpr Int nverts = 7, nfaces = 6, nedges = 12, nfaceverts = 3, vertex_degree = 6;
  static constexpr int lpmx_id = LPMX_SEED_TRI_HEX;
  typedef PlaneGeometry geo;
  typedef TriFace faceKind;
  static std::string filename() { return "triHexSeed.dat"; }
  static std::string id_string() { return "tri_hex"; }
};

/// f = f0 + beta y
struct CoriolisBetaPlane {
  Real f0, beta;
  static constexpr Real Omega = 2 * constants::PI;
  CoriolisBetaPlane() : f0(0), beta(0) {}
  explicit CoriolisBetaPlane(const Real phi0) : f0(2 * Omega * std::sin(phi0)), beta(2 * Omega * std::cos(phi0)) {}
  CoriolisBetaPlane(const Real f0, const Real beta) : f0(f0), beta(beta) {}
  template <typename PtType>
  Real f(const PtType& xy) const { return f0 + beta * xy[1]; }
  template <typename PtType>
  Real dfdt(const PtType& uv) const { return beta * uv[1]; }
  template <typename XType, typename UType>
  Real grad_f_cross_u(const XType&, const UType& u) const { return -beta * u[1]; }
};

/// bottom topography 0.8 exp(-5 |x|^2): the engine's LPMX_TOPO_PLANAR_GAUSSIAN_MOUNTAIN
struct PlanarGaussianMountain {
  typedef PlaneGeometry geo;
  static constexpr Int ndim = 2;
  static constexpr Real mtn_height = 0.8;
  static constexpr Real b = 5.0;
  static constexpr int lpmx_topo = LPMX_TOPO_PLANAR_GAUSSIAN_MOUNTAIN;
  template <typename CV>
  Real operator()(const CV& xy) const { return mtn_height * std::exp(-b * PlaneGeometry::norm2(xy)); }
  template <typename CV>
  Real laplacian(const CV& xy) const {
    return 4 * b * mtn_height * (b * PlaneGeometry::norm2(xy) - 1) * std::exp(-b * PlaneGeometry::norm2(xy));
  }
  std::string name() const { return "PlanarGaussianMountain"; }
};

struct PlanarGaussianSurfacePerturbation {
  typedef PlaneGeometry geo;
  static constexpr Int ndim = 2;
  static constexpr Real H0 = 1.0, ptb_height = 0.1, ptb_bx = 20, ptb_by = 5, ptb_x0 = -1.125, ptb_y0 = 0;
  template <typename CV>
  Real operator()(const CV& xy) const {
    return H0 + ptb_height * std::exp(-(ptb_bx * square(xy[0] - ptb_x0) + ptb_by * square(xy[1] - ptb_y0)));
  }
  std::string name() const { return "PlanarGaussianSurfacePerturbation"; }
};

/// Compactly supported Lamb dipole.  As coded in the reference, sin(theta) is y / r with y NOT measured from the dipole
/// centre (src/lpm_vorticity_gallery.hpp:201); kept.  Deviation, flagged: J0 / J1 are std::cyl_bessel_j here, the
/// reference evaluates its own rational approximations (src/util/lpm_math.hpp:310-420; agreement ~1e-8).
inline Real lamb_dipole_vorticity(const Real x, const Real y, const Real xctr, const Real yctr, const Real dipole_radius,
                                  const Real dipole_strength) {
  static constexpr Real LAMB_K0 = 3.8317;
  const Real r = std::sqrt(square(x - xctr) + square(y - yctr));
  Real result = 0;
  if ((r < dipole_radius) && !(std::abs(r) < constants::ZERO_TOL)) {
    const Real k = LAMB_K0 / dipole_radius;
    const Real sintheta = y / r;
    const Real denom = std::cyl_bessel_j(0.0, LAMB_K0);
    result = -2 * dipole_strength * k * std::cyl_bessel_j(1.0, k * r) * sintheta / denom;
  }
  return result;
}

struct CollidingDipolePairPlane {
  typedef PlaneGeometry geo;
  static constexpr bool IsVorticity = true;
  Real dipole_strengthA = 1, dipole_radiusA = 1;
  Real xyz_ctrA[2] = {-1.5, 0};
  Real dipole_strengthB = -1, dipole_radiusB = 1;
  Real xyz_ctrB[2] = {1.5, 0};
  CollidingDipolePairPlane() = default;
  Real operator()(const Real& x, const Real& y) const {
    return lamb_dipole_vorticity(x, y, xyz_ctrA[0], xyz_ctrA[1], dipole_radiusA, dipole_strengthA) +
           lamb_dipole_vorticity(x, y, xyz_ctrB[0], xyz_ctrB[1], dipole_radiusB, dipole_strengthB);
  }
  template <typename CV>
  Real operator()(const CV& xy) const { return (*this)(xy[0], xy[1]); }
  std::string name() const { return "PlanarCollidingDipoles"; }
};

namespace pse {
template <typename Geo>
struct PSEKernel {
  static constexpr Int ndim = Geo::ndim;
  /// kernel width eps = dx^p, p < 1
  static Real get_epsilon(const Real dx, const Real p = 11.0 / 20) {
    LPM_REQUIRE(p < 1);
    return std::pow(dx, p);
  }
};
struct BivariateOrder8 {
  using geo = PlaneGeometry;
  static constexpr Int ndim = 2;
  static Real laplacian(const Real r) {
    const Real rsq = square(r);
    return (40 * (1 - rsq) + 10 * square(rsq) - 2 * rsq * square(rsq) / 3) * std::exp(-rsq) / constants::PI;
  }
};
}  // namespace pse

// --------------------------------------------------------------------------------------------------------------------
// Incompressible2D on the plane
// --------------------------------------------------------------------------------------------------------------------
template <typename SeedType>
class PlaneIncompressible2D {
 public:
  using geo = PlaneGeometry;
  using Coriolis = CoriolisBetaPlane;

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
  Real t, t_ref, eps;

  PlaneIncompressible2D(const PolyMeshParameters<SeedType>& mp, const Coriolis& coriolis, const Real velocity_eps)
      : ref_crds_passive(mp.nmaxverts), ref_crds_active(mp.nmaxfaces), rel_vort_passive("relative_vorticity", mp.nmaxverts),
        rel_vort_active("relative_vorticity", mp.nmaxfaces), abs_vort_passive("absolute_vorticity", mp.nmaxverts),
        abs_vort_active("absolute_vorticity", mp.nmaxfaces), stream_fn_passive("stream_function", mp.nmaxverts),
        stream_fn_active("stream_function", mp.nmaxfaces), velocity_passive("velocity", mp.nmaxverts),
        velocity_active("velocity", mp.nmaxfaces), ftle("ftle", mp.nmaxfaces), mesh(mp), coriolis(coriolis), t(0), t_ref(0),
        eps(velocity_eps) {
    ko::deep_copy(ref_crds_passive.view, mesh.vertices.phys_crds.view);
    ko::deep_copy(ref_crds_active.view, mesh.faces.phys_crds.view);
  }
  void update_host() {}
  void update_device() {}

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

  /// Incompressible2D{Passive,Active}Sums<PlaneGeometry> (src/lpm_incompressible2d_impl.hpp:235-254)
  void init_direct_sums() {
    lpmx_handle_t h = Engine::get();
    const Index nv = mesh.n_vertices_host(), nf = mesh.n_faces_host();
    Engine::check(lpmx_ic2d_plane_sums(h, mesh.vertices.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, nv,
                                       mesh.faces.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, rel_vort_active.view.data(),
                                       mesh.faces.area.data(), mesh.faces.mask.data(), nf, eps, 0, velocity_passive.view.data(),
                                       stream_fn_passive.view.data()),
                  "Incompressible2DPassiveSums<PlaneGeometry>");
    Engine::check(lpmx_ic2d_plane_sums(h, nullptr, LPMX_LAYOUT_RIGHT, 0, nf, mesh.faces.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0,
                                       rel_vort_active.view.data(), mesh.faces.area.data(), mesh.faces.mask.data(), nf, eps, 1,
                                       velocity_active.view.data(), stream_fn_active.view.data()),
                  "Incompressible2DActiveSums<PlaneGeometry>");
  }

  template <typename SolverType>
  void advance_timestep(SolverType& solver) {
    solver.advance_timestep_impl();
    t = solver.t_idx * solver.dt;
  }

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
    ss << "Incompressible2D<" << SeedType::id_string() << ">: t = " << t << ", eps = " << eps << ", f0 = " << coriolis.f0
       << ", beta = " << coriolis.beta << "\n" << mesh.info_string("", tab_level + 1);
    return ss.str();
  }
};

template <>
class Incompressible2D<QuadRectSeed> : public PlaneIncompressible2D<QuadRectSeed> {
 public:
  using PlaneIncompressible2D<QuadRectSeed>::PlaneIncompressible2D;
};
template <>
class Incompressible2D<TriHexSeed> : public PlaneIncompressible2D<TriHexSeed> {
 public:
  using PlaneIncompressible2D<TriHexSeed>::PlaneIncompressible2D;
};

/// Heun step (src/lpm_incompressible2d_rk2_impl.hpp:75-172) for PlaneGeometry = lpmx_ic2d_plane_rk2_step
template <typename SeedType>
class PlaneIncompressible2DRK2 {
 public:
  Real dt;
  Incompressible2D<SeedType>& ic2d;
  Int t_idx;
  Index n_passive, n_active;
  Real eps;
  PlaneIncompressible2DRK2(const Real dt, Incompressible2D<SeedType>& ic2d, const Index t_idx = 0)
      : dt(dt), ic2d(ic2d), t_idx(t_idx), n_passive(ic2d.mesh.n_vertices_host()), n_active(ic2d.mesh.n_faces_host()),
        eps(ic2d.eps) {}
  void advance_timestep_impl() {
    auto& m = ic2d.mesh;
    Engine::check(lpmx_ic2d_plane_rk2_step(Engine::get(), dt, ic2d.coriolis.f0, ic2d.coriolis.beta, eps, n_passive,
                                           m.vertices.phys_crds.view.data(), ic2d.rel_vort_passive.view.data(),
                                           ic2d.velocity_passive.view.data(), ic2d.stream_fn_passive.view.data(), n_active,
                                           m.faces.phys_crds.view.data(), ic2d.rel_vort_active.view.data(),
                                           ic2d.velocity_active.view.data(), ic2d.stream_fn_active.view.data(),
                                           m.faces.area.data(), m.faces.mask.data(), LPMX_LAYOUT_RIGHT, 0, 0, 1),
                  "Incompressible2DRK2<PlaneGeometry>::advance_timestep_impl");
    ++t_idx;
  }
  std::string info_string(const int tab_level = 0) const {
    std::ostringstream ss;
    ss << std::string(tab_level, '\t') << "Incompressible2DRK2 (plane): dt = " << dt << ", t_idx = " << t_idx << ", eps = " << eps
       << "\n";
    return ss.str();
  }
};
template <>
class Incompressible2DRK2<QuadRectSeed> : public PlaneIncompressible2DRK2<QuadRectSeed> {
 public:
  using PlaneIncompressible2DRK2<QuadRectSeed>::PlaneIncompressible2DRK2;
};
template <>
class Incompressible2DRK2<TriHexSeed> : public PlaneIncompressible2DRK2<TriHexSeed> {
 public:
  using PlaneIncompressible2DRK2<TriHexSeed>::PlaneIncompressible2DRK2;
};

// --------------------------------------------------------------------------------------------------------------------
// SWE on the plane
// --------------------------------------------------------------------------------------------------------------------
template <typename SeedType>
class PlaneSWE {
 public:
  using geo = PlaneGeometry;
  using Coriolis = CoriolisBetaPlane;

  ScalarField<VertexField> rel_vort_passive, pot_vort_passive, div_passive, surf_passive, bottom_passive, surf_lap_passive,
      depth_passive, double_dot_passive, du1dx1_passive, du1dx2_passive, du2dx1_passive, du2dx2_passive, stream_fn_passive,
      potential_passive;
  ScalarField<FaceField> rel_vort_active, pot_vort_active, div_active, surf_active, bottom_active, surf_lap_active, depth_active,
      double_dot_active, du1dx1_active, du1dx2_active, du2dx1_active, du2dx2_active, stream_fn_active, potential_active,
      mass_active;
  VectorField<geo, VertexField> velocity_passive;
  VectorField<geo, FaceField> velocity_active;
  PolyMesh2d<SeedType> mesh;
  Coriolis coriolis;
  std::map<std::string, ScalarField<VertexField>> tracer_passive;
  std::map<std::string, ScalarField<FaceField>> tracer_active;
  Real g, t, eps, pse_eps;

  PlaneSWE(const PolyMeshParameters<SeedType>& mp, const Coriolis& coriolis)
      : rel_vort_passive("relative_vorticity", mp.nmaxverts), pot_vort_passive("potential_vorticity", mp.nmaxverts),
        div_passive("divergence", mp.nmaxverts), surf_passive("surface_height", mp.nmaxverts),
        bottom_passive("bottom_height", mp.nmaxverts), surf_lap_passive("surface_laplacian", mp.nmaxverts),
        depth_passive("depth", mp.nmaxverts), double_dot_passive("double_dot", mp.nmaxverts),
        du1dx1_passive("du1dx1", mp.nmaxverts), du1dx2_passive("du1dx2", mp.nmaxverts), du2dx1_passive("du2dx1", mp.nmaxverts),
        du2dx2_passive("du2dx2", mp.nmaxverts), stream_fn_passive("stream_function", mp.nmaxverts),
        potential_passive("potential", mp.nmaxverts), rel_vort_active("relative_vorticity", mp.nmaxfaces),
        pot_vort_active("potential_vorticity", mp.nmaxfaces), div_active("divergence", mp.nmaxfaces),
        surf_active("surface_height", mp.nmaxfaces), bottom_active("bottom_height", mp.nmaxfaces),
        surf_lap_active("surface_laplacian", mp.nmaxfaces), depth_active("depth", mp.nmaxfaces),
        double_dot_active("double_dot", mp.nmaxfaces), du1dx1_active("du1dx1", mp.nmaxfaces), du1dx2_active("du1dx2", mp.nmaxfaces),
        du2dx1_active("du2dx1", mp.nmaxfaces), du2dx2_active("du2dx2", mp.nmaxfaces),
        stream_fn_active("stream_function", mp.nmaxfaces), potential_active("potential", mp.nmaxfaces),
        mass_active("mass", mp.nmaxfaces), velocity_passive("velocity", mp.nmaxverts), velocity_active("velocity", mp.nmaxfaces),
        mesh(mp), coriolis(coriolis), g(1), t(0), eps(0), pse_eps(0) {}

  /// (src/lpm_swe_impl.hpp:99-107)
  void set_kernel_parameters(const Real vel_eps, const Real pse) {
    LPM_REQUIRE(vel_eps >= 0);
    LPM_REQUIRE(pse > 0);
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
      bottom_passive.view(i) = b, surf_passive.view(i) = s, depth_passive.view(i) = s - b;
    }
    const auto fx = mesh.faces.phys_crds.view;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) {
      const Real b = topo(fx.row(i)), s = sfc(fx.row(i));
      bottom_active.view(i) = b, surf_active.view(i) = s, depth_active.view(i) = s - b;
      mass_active.view(i) = (s - b) * mesh.faces.area(i);
    }
  }
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

  void allocate_scalar_tracer(const std::string& name) {
    tracer_passive.emplace(name, ScalarField<VertexField>(name, mesh.params.nmaxverts));
    tracer_active.emplace(name, ScalarField<FaceField>(name, mesh.params.nmaxfaces));
  }

  /// PlanarSWEVertexSums / PlanarSWEFaceSums at the current state (src/lpm_swe_impl.hpp:406-426)
  void init_direct_sums(const bool do_velocity = true) {
    lpmx_handle_t h = Engine::get();
    const Index nv = mesh.n_vertices_host(), nf = mesh.n_faces_host();
    const lpmx_plane_swe_sums_t po{velocity_passive.view.data(), double_dot_passive.view.data(), du1dx1_passive.view.data(),
                                   du1dx2_passive.view.data(),   du2dx1_passive.view.data(),     du2dx2_passive.view.data(),
                                   surf_lap_passive.view.data(), stream_fn_passive.view.data(),  potential_passive.view.data()};
    Engine::check(lpmx_swe_plane_sums(h, mesh.vertices.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, surf_passive.view.data(), nv,
                                      mesh.faces.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, rel_vort_active.view.data(),
                                      div_active.view.data(), mesh.faces.area.data(), mesh.faces.mask.data(),
                                      surf_active.view.data(), nf, eps, pse_eps, 0, do_velocity, &po),
                  "PlanarSWEVertexSums");
    const lpmx_plane_swe_sums_t ao{velocity_active.view.data(), double_dot_active.view.data(), du1dx1_active.view.data(),
                                   du1dx2_active.view.data(),   du2dx1_active.view.data(),     du2dx2_active.view.data(),
                                   surf_lap_active.view.data(), stream_fn_active.view.data(),  potential_active.view.data()};
    Engine::check(lpmx_swe_plane_sums(h, nullptr, LPMX_LAYOUT_RIGHT, 0, surf_active.view.data(), nf,
                                      mesh.faces.phys_crds.view.data(), LPMX_LAYOUT_RIGHT, 0, rel_vort_active.view.data(),
                                      div_active.view.data(), mesh.faces.area.data(), mesh.faces.mask.data(),
                                      surf_active.view.data(), nf, eps, pse_eps, 1, do_velocity, &ao),
                  "PlanarSWEFaceSums");
  }

  template <typename SolverType>
  void advance_timestep(SolverType& solver) {
    solver.advance_timestep_impl();
    t = solver.t_idx * solver.dt;
  }

  /// sum of mass over the leaves (conserved exactly: mass is a Lagrangian invariant of the scheme)
  Real total_mass() const {
    Real s = 0;
    for (Index i = 0; i < mesh.n_faces_host(); ++i) s += (mesh.faces.mask(i) ? 0 : mass_active.view(i));
    return s;
  }

  std::string info_string(const int tab_level = 0, const bool = false) const {
    std::ostringstream ss;
    ss << "SWE<" << SeedType::id_string() << ">: t = " << t << ", g = " << g << ", eps = " << eps << ", pse_eps = " << pse_eps
       << ", f0 = " << coriolis.f0 << ", beta = " << coriolis.beta << "\n" << mesh.info_string("", tab_level + 1);
    return ss.str();
  }
};

template <>
class SWE<QuadRectSeed> : public PlaneSWE<QuadRectSeed> {
 public:
  using PlaneSWE<QuadRectSeed>::PlaneSWE;
};
template <>
class SWE<TriHexSeed> : public PlaneSWE<TriHexSeed> {
 public:
  using PlaneSWE<TriHexSeed>::PlaneSWE;
};

/// SWERK4<Seed, Topo>::advance_timestep_impl (src/lpm_swe_rk4_impl.hpp:203-445) = lpmx_swe_plane_rk4_step, in place on the
/// SWE object's views: 4 pair-sum launches + 5 fused stage kernels per step.  Topo: ZeroFunctor or PlanarGaussianMountain.
template <typename SeedType, typename TopoType = ZeroFunctor>
class SWERK4 {
  static_assert(std::is_same<typename SeedType::geo, PlaneGeometry>::value, "SWERK4 is implemented for planar problems");

 public:
  Real dt;
  Int t_idx;
  SWE<SeedType>& swe;
  TopoType topo;
  Real eps, pse_eps;

  SWERK4(const Real timestep, SWE<SeedType>& swe_mesh, TopoType& topo)
      : dt(timestep), t_idx(0), swe(swe_mesh), topo(topo), eps(swe_mesh.eps), pse_eps(swe_mesh.pse_eps) {}

  void advance_timestep_impl() {
    auto& m = swe.mesh;
    const lpmx_plane_swe_passive_t P{m.vertices.phys_crds.view.data(), swe.rel_vort_passive.view.data(),
                                     swe.div_passive.view.data(),      swe.depth_passive.view.data(),
                                     swe.surf_passive.view.data(),     swe.bottom_passive.view.data(),
                                     swe.velocity_passive.view.data(), swe.double_dot_passive.view.data(),
                                     swe.du1dx1_passive.view.data(),   swe.du1dx2_passive.view.data(),
                                     swe.du2dx1_passive.view.data(),   swe.du2dx2_passive.view.data(),
                                     swe.surf_lap_passive.view.data(), swe.stream_fn_passive.view.data(),
                                     swe.potential_passive.view.data()};
    const lpmx_plane_swe_active_t A{m.faces.phys_crds.view.data(),   swe.rel_vort_active.view.data(), swe.div_active.view.data(),
                                    m.faces.area.data(),             swe.mass_active.view.data(),     swe.depth_active.view.data(),
                                    swe.surf_active.view.data(),     swe.bottom_active.view.data(),   swe.velocity_active.view.data(),
                                    swe.double_dot_active.view.data(), swe.du1dx1_active.view.data(), swe.du1dx2_active.view.data(),
                                    swe.du2dx1_active.view.data(),   swe.du2dx2_active.view.data(),   swe.surf_lap_active.view.data(),
                                    swe.stream_fn_active.view.data(), swe.potential_active.view.data(), m.faces.mask.data()};
    Engine::check(lpmx_swe_plane_rk4_step(Engine::get(), dt, swe.coriolis.f0, swe.coriolis.beta, swe.g, eps, pse_eps, topo_id(),
                                          m.n_vertices_host(), &P, m.n_faces_host(), &A, LPMX_LAYOUT_RIGHT, 0, 0, 1),
                  "SWERK4::advance_timestep_impl");
    ++t_idx;
  }

  std::string info_string(const int tab_level = 0) const {
    std::ostringstream ss;
    ss << std::string(tab_level, '\t') << "SWERK4: dt = " << dt << ", t_idx = " << t_idx << ", eps = " << eps
       << ", pse_eps = " << pse_eps << ", topography " << topo.name() << "\n";
    return ss.str();
  }

 private:
  static constexpr int topo_id() {
    static_assert(std::is_same<TopoType, ZeroFunctor>::value || std::is_same<TopoType, PlanarGaussianMountain>::value,
                  "the engine carries ZeroFunctor and PlanarGaussianMountain bottom topographies");
    return std::is_same<TopoType, PlanarGaussianMountain>::value ? LPMX_TOPO_PLANAR_GAUSSIAN_MOUNTAIN : LPMX_TOPO_ZERO;
  }
};

}  // namespace Lpm
#endif
