// lpm/lpm_compadre_remesh.hpp -- CompadreRemesh<SeedType> for uniform meshes (src/mesh/lpm_compadre_remesh.hpp:18-112,
// _impl.hpp:20-210,316-351) and compadre_remesh(new_ic2d, old_ic2d, gmls_params) (src/lpm_incompressible2d_impl.hpp:389-456):
// the hand-off that rebuilds the particle set every remesh_interval steps (examples/sphere_rh54.cpp:257-300).
//   interpolate_lag_crds     lag x, y, z as three scalar point evaluations, then normalised            _impl.hpp:136-162
//   uniform_direct_remesh    + every scalar field and every vector field interpolated                  _impl.hpp:182-210
//   uniform_indirect_remesh  + omega = vorticity(a) + f(a), zeta = omega - f(x), tracers tracer(a); vectors interpolated  _impl.hpp:316-351
// All interpolation is ONE device call (lpmx_gmls_sphere_interpolate) on the gathered old particles (vertices + leaves);
// gather/scatter are lpmx_gather_mesh_data / lpmx_scatter_mesh_data.  Deviation, flagged: the reference reconstructs
// vector fields with Compadre's ManifoldVectorPointSample basis; here the Cartesian components are interpolated as
// scalars and the result is projected onto the tangent plane of the target (same order of accuracy, not the same
// numbers -- and Compadre is unpinned anyway).
//   adaptive_direct_remesh / adaptive_indirect_remesh   the uniform remesh, then amr_limit passes of {flag the faces the
//                            previous pass added, divide_flagged_faces, interpolate everything again}   _impl.hpp:212-314,354-661
#ifndef LPM_SHIM_COMPADRE_REMESH_HPP
#define LPM_SHIM_COMPADRE_REMESH_HPP

#include <map>
#include <string>
#include <vector>

#include "lpm_coriolis.hpp"
#include "lpm_polymesh2d.hpp"
#include "lpm_refinement.hpp"
#include "lpm_swe.hpp"  // gmls::Params

namespace Lpm {

template <typename SeedType>
struct CompadreRemesh {
  static_assert(std::is_same<typename SeedType::geo, SphereGeometry>::value, "the engine's GMLS interpolation is spherical");
  using vert_scalar_field_map = std::map<std::string, ScalarField<VertexField>>;
  using vert_vector_field_map = std::map<std::string, VectorField<typename SeedType::geo, VertexField>>;
  using face_scalar_field_map = std::map<std::string, ScalarField<FaceField>>;
  using face_vector_field_map = std::map<std::string, VectorField<typename SeedType::geo, FaceField>>;
  using CoriolisType = CoriolisSphere;

  gmls::Params gmls_params;
  Logger logger{"CompadreRemesh", Log::warn};
  PolyMesh2d<SeedType>& new_mesh;
  vert_scalar_field_map new_vert_scalars;
  face_scalar_field_map new_face_scalars;
  vert_vector_field_map new_vert_vectors;
  face_vector_field_map new_face_vectors;
  const PolyMesh2d<SeedType>& old_mesh;
  vert_scalar_field_map old_vert_scalars;
  face_scalar_field_map old_face_scalars;
  vert_vector_field_map old_vert_vectors;
  face_vector_field_map old_face_vectors;

  CompadreRemesh(PolyMesh2d<SeedType>& new_mesh, vert_scalar_field_map& new_vert_scalars, face_scalar_field_map& new_face_scalars,
                 vert_vector_field_map& new_vert_vectors, face_vector_field_map& new_face_vectors,
                 const PolyMesh2d<SeedType>& old_mesh, const vert_scalar_field_map& old_vert_scalars,
                 const face_scalar_field_map& old_face_scalars, const vert_vector_field_map& old_vert_vectors,
                 const face_vector_field_map& old_face_vectors, const gmls::Params& params)
      : gmls_params(params), new_mesh(new_mesh), new_vert_scalars(new_vert_scalars), new_face_scalars(new_face_scalars),
        new_vert_vectors(new_vert_vectors), new_face_vectors(new_face_vectors), old_mesh(old_mesh),
        old_vert_scalars(old_vert_scalars), old_face_scalars(old_face_scalars), old_vert_vectors(old_vert_vectors),
        old_face_vectors(old_face_vectors) {}

  void uniform_direct_remesh() { remesh(true, nullptr); }

  template <typename VorticityFunctor>
  void uniform_indirect_remesh(const VorticityFunctor& vorticity, const CoriolisType& coriolis) {
    remesh(false, [&](const Real* a, const Real* x, std::map<std::string, Real>& out) {
      const Real omega = vorticity(a) + coriolis.f(a);
      out["absolute_vorticity"] = omega;
      out["relative_vorticity"] = omega - coriolis.f(x);
    });
  }
  template <typename VorticityFunctor, typename Tracer1>
  void uniform_indirect_remesh(const VorticityFunctor& vorticity, const CoriolisType& coriolis, const Tracer1& tracer1) {
    remesh(false, [&](const Real* a, const Real* x, std::map<std::string, Real>& out) {
      const Real omega = vorticity(a) + coriolis.f(a);
      out["absolute_vorticity"] = omega;
      out["relative_vorticity"] = omega - coriolis.f(x);
      out[tracer1.name()] = tracer1(a);
    });
  }
  template <typename VorticityFunctor, typename Tracer1, typename Tracer2>
  void uniform_indirect_remesh(const VorticityFunctor& vorticity, const CoriolisType& coriolis, const Tracer1& tracer1,
                               const Tracer2& tracer2) {
    remesh(false, [&](const Real* a, const Real* x, std::map<std::string, Real>& out) {
      const Real omega = vorticity(a) + coriolis.f(a);
      out["absolute_vorticity"] = omega;
      out["relative_vorticity"] = omega - coriolis.f(x);
      out[tracer1.name()] = tracer1(a);
      out[tracer2.name()] = tracer2(a);
    });
  }

  /// uniform_direct_remesh, then new_mesh.params.amr_limit refinement passes, each followed by a fresh interpolation of
  /// the Lagrangian coordinates and every field onto the refined particle set (_impl.hpp:212-314)
  template <typename FlagType>
  void adaptive_direct_remesh(Refinement<SeedType>& refiner, const FlagType& flag) {
    uniform_direct_remesh();
    adaptive_passes(refiner, flag, true, nullptr);
  }
  template <typename FlagType, typename VorticityFunctor>
  void adaptive_indirect_remesh(Refinement<SeedType>& refiner, const FlagType& flag, const VorticityFunctor& vorticity,
                                const CoriolisType& coriolis) {
    uniform_indirect_remesh(vorticity, coriolis);
    adaptive_passes(refiner, flag, false, last_point_fn_);
  }
  template <typename FlagType, typename VorticityFunctor, typename Tracer1>
  void adaptive_indirect_remesh(Refinement<SeedType>& refiner, const FlagType& flag, const VorticityFunctor& vorticity,
                                const CoriolisType& coriolis, const Tracer1& tracer1) {
    uniform_indirect_remesh(vorticity, coriolis, tracer1);
    adaptive_passes(refiner, flag, false, last_point_fn_);
  }
  template <typename FlagType, typename VorticityFunctor, typename Tracer1, typename Tracer2>
  void adaptive_indirect_remesh(Refinement<SeedType>& refiner, const FlagType& flag, const VorticityFunctor& vorticity,
                                const CoriolisType& coriolis, const Tracer1& tracer1, const Tracer2& tracer2) {
    uniform_indirect_remesh(vorticity, coriolis, tracer1, tracer2);
    adaptive_passes(refiner, flag, false, last_point_fn_);
  }

 private:
  typedef std::function<void(const Real* lag, const Real* phys, std::map<std::string, Real>&)> PointFn;
  PointFn last_point_fn_;  // the indirect definition of the fields, kept for the adaptive passes

  template <typename FlagType>
  void adaptive_passes(Refinement<SeedType>& refiner, const FlagType& flag, const bool direct, const PointFn& point_fn) {
    Index face_start_idx = 0;
    for (int i = 0; i < new_mesh.params.amr_limit; ++i) {
      const Index face_end_idx = new_mesh.n_faces_host();
      refiner.iterate(face_start_idx, face_end_idx, flag);
      new_mesh.divide_flagged_faces(refiner.flags, logger);
      do_remesh(direct, point_fn);
      face_start_idx = face_end_idx;
    }
  }

  void remesh(const bool direct, const PointFn& point_fn) {
    last_point_fn_ = point_fn;
    do_remesh(direct, point_fn);
  }

  // gathered copy (vertices then leaves) of an n x ncomp vertex/face pair of host arrays
  static std::vector<Real> gather(const PolyMesh2d<SeedType>& m, int ncomp, const Real* vdata, const Real* fdata) {
    int n = 0;
    lpmx_handle_t h = Engine::get();
    const Index nv = m.n_vertices_host(), nf = m.n_faces_host();
    Engine::check(lpmx_gather_mesh_data(h, ncomp, LPMX_LAYOUT_RIGHT, nv, vdata, 0, nf, fdata, 0, m.faces.mask.data(), nullptr, 0, &n),
                  "GatherMeshData");
    std::vector<Real> g((size_t)ncomp * n);
    Engine::check(lpmx_gather_mesh_data(h, ncomp, LPMX_LAYOUT_RIGHT, nv, vdata, 0, nf, fdata, 0, m.faces.mask.data(), g.data(), 0, &n),
                  "GatherMeshData");
    return g;
  }
  static void scatter(const PolyMesh2d<SeedType>& m, int ncomp, const std::vector<Real>& g, Real* vdata, Real* fdata) {
    Engine::check(lpmx_scatter_mesh_data(Engine::get(), ncomp, LPMX_LAYOUT_RIGHT, g.data(), 0, m.n_vertices_host(), vdata, 0,
                                         m.n_faces_host(), fdata, 0, m.faces.mask.data()),
                  "ScatterMeshData");
  }

  void do_remesh(const bool direct, const PointFn& point_fn) {
    // sources: the old particles where they are now; targets: the new mesh's particles
    const std::vector<Real> src = gather(old_mesh, 3, old_mesh.vertices.phys_crds.view.data(), old_mesh.faces.phys_crds.view.data());
    const std::vector<Real> tgt = gather(new_mesh, 3, new_mesh.vertices.phys_crds.view.data(), new_mesh.faces.phys_crds.view.data());
    const int ns = (int)(src.size() / 3), nt = (int)(tgt.size() / 3);
    // field list: lag x, y, z; (direct) every scalar; every vector component
    std::vector<std::vector<Real>> in;
    std::vector<std::string> scalar_names, vector_names;
    {
      const std::vector<Real> lag = gather(old_mesh, 3, old_mesh.vertices.lag_crds.view.data(), old_mesh.faces.lag_crds.view.data());
      for (int k = 0; k < 3; ++k) {
        in.emplace_back(ns);
        for (int i = 0; i < ns; ++i) in.back()[i] = lag[3 * (size_t)i + k];
      }
    }
    if (direct)
      for (const auto& sf : old_vert_scalars) {
        scalar_names.push_back(sf.first);
        in.push_back(gather(old_mesh, 1, sf.second.view.data(), old_face_scalars.at(sf.first).view.data()));
      }
    for (const auto& vf : old_vert_vectors) {
      vector_names.push_back(vf.first);
      const std::vector<Real> v = gather(old_mesh, 3, vf.second.view.data(), old_face_vectors.at(vf.first).view.data());
      for (int k = 0; k < 3; ++k) {
        in.emplace_back(ns);
        for (int i = 0; i < ns; ++i) in.back()[i] = v[3 * (size_t)i + k];
      }
    }
    std::vector<std::vector<Real>> out(in.size(), std::vector<Real>(nt));
    std::vector<const Real*> pin(in.size());
    std::vector<Real*> pout(in.size());
    for (size_t f = 0; f < in.size(); ++f) pin[f] = in[f].data(), pout[f] = out[f].data();
    const lpmx_gmls_params_t cp = gmls_params.c_params();
    Engine::check(lpmx_gmls_sphere_interpolate(Engine::get(), &cp, ns, src.data(), LPMX_LAYOUT_RIGHT, 0, (int)in.size(), pin.data(), nt,
                                               tgt.data(), LPMX_LAYOUT_RIGHT, 0, pout.data()),
                  "CompadreRemesh: scalar point evaluation");
    // interpolate_lag_crds: assemble and normalise (_impl.hpp:153-161)
    std::vector<Real> lag_new(3 * (size_t)nt);
    for (int i = 0; i < nt; ++i) {
      Real a[3] = {out[0][i], out[1][i], out[2][i]};
      const Real s = 1.0 / std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      for (int k = 0; k < 3; ++k) lag_new[3 * (size_t)i + k] = a[k] * s;
    }
    size_t f = 3;
    if (direct) {
      for (const auto& name : scalar_names) {
        scatter(new_mesh, 1, out[f], new_vert_scalars.at(name).view.data(), new_face_scalars.at(name).view.data());
        ++f;
      }
    } else {
      // indirect: functions of the interpolated Lagrangian coordinate (_impl.hpp:322-333)
      std::map<std::string, std::vector<Real>> vals;
      std::map<std::string, Real> pt;
      for (int i = 0; i < nt; ++i) {
        pt.clear();
        point_fn(&lag_new[3 * (size_t)i], &tgt[3 * (size_t)i], pt);
        for (const auto& kv : pt) {
          auto& v = vals[kv.first];
          if (v.empty()) v.resize(nt);
          v[i] = kv.second;
        }
      }
      for (const auto& kv : vals)
        scatter(new_mesh, 1, kv.second, new_vert_scalars.at(kv.first).view.data(), new_face_scalars.at(kv.first).view.data());
    }
    for (const auto& name : vector_names) {
      std::vector<Real> v(3 * (size_t)nt);
      for (int i = 0; i < nt; ++i) {
        const Real* x = &tgt[3 * (size_t)i];
        const Real u[3] = {out[f][i], out[f + 1][i], out[f + 2][i]};
        const Real r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
        const Real un = (u[0] * x[0] + u[1] * x[1] + u[2] * x[2]) / r2;
        for (int k = 0; k < 3; ++k) v[3 * (size_t)i + k] = u[k] - un * x[k];  // tangent projection at the target
      }
      scatter(new_mesh, 3, v, new_vert_vectors.at(name).view.data(), new_face_vectors.at(name).view.data());
      f += 3;
    }
    // new_scatter->scatter_lag_crds()
    scatter(new_mesh, 3, lag_new, new_mesh.vertices.lag_crds.view.data(), new_mesh.faces.lag_crds.view.data());
  }
};

}  // namespace Lpm
#endif
