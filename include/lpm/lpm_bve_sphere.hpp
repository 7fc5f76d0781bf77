// lpm/lpm_bve_sphere.hpp -- BVESphere<Seed> and BVERK4 with the reference's public members; every O(N^2)
// evaluation and the whole RK4 step run in the sm_100a engine behind the C ABI.
//   BVESphere<Seed>                    src/lpm_bve_sphere.hpp:24-106, src/lpm_bve_sphere_impl.hpp:10-231
//   BVERK4                             src/lpm_bve_rk4.hpp:15-92, src/lpm_bve_rk4_impl.hpp:55-167
//   BVE{Vertex,Face}{Velocity,StreamFn} (free functions here)   src/lpm_bve_sphere_kernels.hpp:141-211,329-394
#ifndef LPM_SHIM_BVE_SPHERE_HPP
#define LPM_SHIM_BVE_SPHERE_HPP

#include "lpm_polymesh2d.hpp"

namespace Lpm {

/// BVEVertexVelocity: u(x_i) at passive targets from the active sources (src/lpm_bve_sphere_kernels.hpp:179-211)
inline void bve_vertex_velocity(const vec3_view_type& u, const vec3_view_type& vx, const Index nverts, const vec3_view_type& fx,
                                const scalar_view_type& zeta, const scalar_view_type& area, const mask_view_type& mask,
                                const Index nsrc) {
  Engine::check(lpmx_bve_velocity(Engine::get(), vx.data(), LPMX_LAYOUT_RIGHT, 0, nverts, fx.data(), LPMX_LAYOUT_RIGHT, 0,
                                  zeta.data(), area.data(), mask.data(), nsrc, 0, u.data()),
                "BVEVertexVelocity");
}
/// BVEFaceVelocity (collocated, skips j == i; :365-394)
inline void bve_face_velocity(const vec3_view_type& u, const vec3_view_type& fx, const scalar_view_type& zeta,
                              const scalar_view_type& area, const mask_view_type& mask, const Index nsrc) {
  Engine::check(lpmx_bve_velocity(Engine::get(), nullptr, LPMX_LAYOUT_RIGHT, 0, nsrc, fx.data(), LPMX_LAYOUT_RIGHT, 0,
                                  zeta.data(), area.data(), mask.data(), nsrc, 1, u.data()),
                "BVEFaceVelocity");
}
/// BVEVertexStreamFn (:141-170)
inline void bve_vertex_stream_fn(const scalar_view_type& psi, const vec3_view_type& vx, const Index nverts,
                                 const vec3_view_type& fx, const scalar_view_type& zeta, const scalar_view_type& area,
                                 const mask_view_type& mask, const Index nsrc) {
  Engine::check(lpmx_bve_streamfn(Engine::get(), vx.data(), LPMX_LAYOUT_RIGHT, 0, nverts, fx.data(), LPMX_LAYOUT_RIGHT, 0,
                                  zeta.data(), area.data(), mask.data(), nsrc, 0, psi.data()),
                "BVEVertexStreamFn");
}
/// BVEFaceStreamFn (:329-356)
inline void bve_face_stream_fn(const scalar_view_type& psi, const vec3_view_type& fx, const scalar_view_type& zeta,
                               const scalar_view_type& area, const mask_view_type& mask, const Index nsrc) {
  Engine::check(lpmx_bve_streamfn(Engine::get(), nullptr, LPMX_LAYOUT_RIGHT, 0, nsrc, fx.data(), LPMX_LAYOUT_RIGHT, 0,
                                  zeta.data(), area.data(), mask.data(), nsrc, 1, psi.data()),
                "BVEFaceStreamFn");
}

/// SphereTangentFunctor (src/lpm_sphere_functions.hpp:62-78): out(i) = x_i . u_i
inline void sphere_tangent(const scalar_view_type& out, const vec3_view_type& x, const vec3_view_type& u, const Index n) {
  for (Index i = 0; i < n; ++i) out(i) = SphereGeometry::dot(x.row(i), u.row(i));
}

template <typename SeedType>
class BVESphere : public PolyMesh2d<SeedType> {
 public:
  typedef scalar_view_type scalar_field;
  typedef vec3_view_type vector_field;
  typedef SeedType seed_type;
  typedef typename SeedType::geo Geo;
  typedef typename SeedType::faceKind FaceType;

  ScalarField<VertexField> rel_vort_verts, abs_vort_verts, stream_fn_verts;
  VectorField<SphereGeometry, VertexField> velocity_verts;
  ScalarField<FaceField> rel_vort_faces, abs_vort_faces, stream_fn_faces;
  VectorField<SphereGeometry, FaceField> velocity_faces;
  Real Omega;  ///< background rotation rate about the positive z-axis
  Real t;      ///< time
  std::vector<ScalarField<VertexField>> tracer_verts;
  std::vector<ScalarField<FaceField>> tracer_faces;

  BVESphere(const Index nmaxverts, const Index nmaxedges, const Index nmaxfaces, const Int nq = 0)
      : PolyMesh2d<SeedType>(nmaxverts, nmaxedges, nmaxfaces), Omega(2 * constants::PI), t(0), omg_set(false) {
    alloc(nmaxverts, nmaxfaces);
    for (int k = 0; k < nq; ++k) {
      tracer_verts.emplace_back("tracer" + std::to_string(k), nmaxverts);
      tracer_faces.emplace_back("tracer" + std::to_string(k), nmaxfaces);
    }
  }
  BVESphere(const Index nmaxverts, const Index nmaxedges, const Index nmaxfaces, const std::vector<std::string>& tracers)
      : PolyMesh2d<SeedType>(nmaxverts, nmaxedges, nmaxfaces), Omega(2 * constants::PI), t(0), omg_set(false) {
    alloc(nmaxverts, nmaxfaces);
    for (const auto& name : tracers) {
      tracer_verts.emplace_back(name, nmaxverts);
      tracer_faces.emplace_back(name, nmaxfaces);
    }
  }

  /// zeta = fn(x, y, z), abs = zeta + 2 Omega z on every vertex and face (src/lpm_bve_sphere_impl.hpp:148-179)
  template <typename VorticityInitialCondition>
  void init_vorticity(const VorticityInitialCondition& vorticity_fn) {
    const auto vx = this->vertices.phys_crds.view;
    for (Index i = 0; i < this->n_vertices_host(); ++i) {
      const Real zeta = vorticity_fn(vx(i, 0), vx(i, 1), vx(i, 2));
      rel_vort_verts.view(i) = zeta;
      abs_vort_verts.view(i) = zeta + 2 * Omega * vx(i, 2);
    }
    const auto fx = this->faces.phys_crds.view;
    for (Index i = 0; i < this->n_faces_host(); ++i) {
      const Real zeta = vorticity_fn(fx(i, 0), fx(i, 1), fx(i, 2));
      rel_vort_faces.view(i) = zeta;
      abs_vort_faces.view(i) = zeta + 2 * Omega * fx(i, 2);
    }
  }

  /// (:181-207)
  void init_velocity() {
    bve_vertex_velocity(velocity_verts.view, this->vertices.phys_crds.view, this->n_vertices_host(), this->faces.phys_crds.view,
                        rel_vort_faces.view, this->faces.area, this->faces.mask, this->n_faces_host());
    bve_face_velocity(velocity_faces.view, this->faces.phys_crds.view, rel_vort_faces.view, this->faces.area, this->faces.mask,
                      this->n_faces_host());
  }
  /// (:209-231)
  void init_stream_fn() {
    bve_vertex_stream_fn(stream_fn_verts.view, this->vertices.phys_crds.view, this->n_vertices_host(), this->faces.phys_crds.view,
                         rel_vort_faces.view, this->faces.area, this->faces.mask, this->n_faces_host());
    bve_face_stream_fn(stream_fn_faces.view, this->faces.phys_crds.view, rel_vort_faces.view, this->faces.area, this->faces.mask,
                       this->n_faces_host());
  }
  void update_device() const override {}
  void update_host() const override {}
  void set_omega(const Real& omg) {
    if (!omg_set) {
      Omega = omg;
      omg_set = true;
    }
  }
  Real avg_mesh_size_radians() const { return this->faces.appx_mesh_size(); }
  Real avg_mesh_size_degrees() const { return 180.0 / constants::PI * avg_mesh_size_radians(); }
  std::string info_string(const std::string& label = "", const int tab_level = 0, const bool dump = false) const override {
    std::ostringstream ss;
    ss << "BVESphere: Omega = " << Omega << ", t = " << t << ", " << tracer_verts.size() << " tracers\n"
       << PolyMesh2d<SeedType>::info_string(label, tab_level + 1, dump);
    return ss.str();
  }

 protected:
  bool omg_set;

 private:
  void alloc(const Index nv, const Index nf) {
    rel_vort_verts = ScalarField<VertexField>("rel_vort_verts", nv);
    abs_vort_verts = ScalarField<VertexField>("abs_vort_verts", nv);
    stream_fn_verts = ScalarField<VertexField>("stream_fn_verts", nv);
    velocity_verts = VectorField<SphereGeometry, VertexField>("velocity_verts", nv);
    rel_vort_faces = ScalarField<FaceField>("rel_vort_faces", nf);
    abs_vort_faces = ScalarField<FaceField>("abs_vort_faces", nf);
    stream_fn_faces = ScalarField<FaceField>("stream_fn_faces", nf);
    velocity_faces = VectorField<SphereGeometry, FaceField>("velocity_faces", nf);
  }
};

/// Fourth-order Runge-Kutta for BVESphere (src/lpm_bve_rk4.hpp, src/lpm_bve_rk4_impl.hpp:55-167).  One call =
/// lpmx_bve_rk4_step: 4 pair-sum launches + 4 fused stage kernels, in place on the sphere's views.
class BVERK4 {
 public:
  Real dt;
  Real Omega;
  Index nverts;
  Index nfaces;

  template <typename SeedType>
  BVERK4(const Real timestep, BVESphere<SeedType>& sph)
      : dt(timestep), Omega(sph.Omega), nverts(sph.n_vertices_host()), nfaces(sph.n_faces_host()) {}

  template <typename SeedType>
  void advance_timestep(BVESphere<SeedType>& sph, const int n_steps = 1) {
    Engine::check(lpmx_bve_rk4_step(Engine::get(), dt, Omega, nverts, sph.vertices.phys_crds.view.data(),
                                    sph.rel_vort_verts.view.data(), sph.velocity_verts.view.data(), nfaces,
                                    sph.faces.phys_crds.view.data(), sph.rel_vort_faces.view.data(),
                                    sph.velocity_faces.view.data(), sph.faces.area.data(), sph.faces.mask.data(),
                                    LPMX_LAYOUT_RIGHT, 0, 0, n_steps),
                  "BVERK4::advance_timestep");
  }
};

}  // namespace Lpm
#endif
