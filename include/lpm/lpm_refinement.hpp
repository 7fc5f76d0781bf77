// lpm/lpm_refinement.hpp -- Refinement<Seed> and the refinement-flag functors, evaluated on the device through
// lpmx_refine_flag / lpmx_refine_flag_max.
//   Refinement<Seed>::iterate (1, 2 or 3 flags)     src/mesh/lpm_refinement.hpp:20-92
//   FlowMapVariationFlag<Seed>                      src/mesh/lpm_refinement_flags.hpp:55-130
//   ScalarMaxFlag / ScalarIntegralFlag / ScalarVariationFlag    :132-183, :185-229, :231-310
// Same constructor arguments, public members (relative_tol, tol, nfaces, flags), set_tol_from_relative_value(),
// description() and info_string().  Where the reference passes the functor to Kokkos::parallel_for over [start, end), here
// Refinement::iterate calls its apply(start, end): one kernel launch that switches flags on in that range and returns the
// number of flags set there.  NeighborsFlag (:30-53; unused by the drivers) runs on the host, where the edge tree lives.
#ifndef LPM_SHIM_REFINEMENT_HPP
#define LPM_SHIM_REFINEMENT_HPP

#include <algorithm>
#include <sstream>
#include <vector>

#include "lpm_polymesh2d.hpp"

namespace Lpm {

typedef mask_view_type flag_view;  // Kokkos::View<bool*>: one byte per face

namespace impl {
inline Real flag_max(const lpmx_flag_desc_t& d) {
  Real mx = 0;
  Engine::check(lpmx_refine_flag_max(Engine::get(), &d, &mx), "set_tol_from_relative_value");
  return mx;
}
inline Index flag_apply(lpmx_flag_desc_t d, const flag_view& flags, const Index start, const Index end) {
  // the functor was built for the faces that existed then; iterate() may run it over faces added since
  d.n_faces = std::max(d.n_faces, (int)end);
  LPM_REQUIRE((size_t)d.n_faces <= flags.extent(0));
  int count = 0;
  Engine::check(lpmx_refine_flag(Engine::get(), &d, start, end, 0, flags.data(), &count), "refinement flag");
  return count;
}
template <typename FlagType>
std::string flag_info_string(const FlagType& f) {
  std::ostringstream ss;
  ss << f.description() << ": relative_tol = " << f.relative_tol << ", tol = " << f.tol << "\n";
  return ss.str();
}
}  // namespace impl

struct ScalarMaxFlag {
  flag_view flags;
  scalar_view_type face_vals;
  mask_view_type facemask;
  Index nfaces;
  Real relative_tol;
  Real tol;
  ScalarMaxFlag(flag_view f, const scalar_view_type fv, const mask_view_type m, const Index n, const Real rtol)
      : flags(f), face_vals(fv), facemask(m), nfaces(n), relative_tol(rtol), tol(rtol) {
    LPM_REQUIRE(rtol > 0);
  }
  /// tol = relative_tol * max over ALL faces of |f| (divided faces included, as in the reference)
  void set_tol_from_relative_value() { tol = relative_tol * impl::flag_max(desc()); }
  Index apply(const Index start, const Index end) const { return impl::flag_apply(desc(), flags, start, end); }
  std::string description() const { return "ScalarMaxFlag"; }
  std::string info_string() const { return impl::flag_info_string(*this); }

 private:
  lpmx_flag_desc_t desc() const {
    lpmx_flag_desc_t d{};
    d.kind = LPMX_FLAG_SCALAR_MAX, d.n_faces = nfaces, d.face_vals = face_vals.data(), d.mask = facemask.data(), d.tol = tol;
    return d;
  }
};

struct ScalarIntegralFlag {
  flag_view flags;
  scalar_view_type face_vals;
  scalar_view_type area;
  mask_view_type facemask;
  Index nfaces;
  Real relative_tol;
  Real tol;
  ScalarIntegralFlag(flag_view f, const scalar_view_type fv, const scalar_view_type a, const mask_view_type m, const Index n,
                     const Real rtol)
      : flags(f), face_vals(fv), area(a), facemask(m), nfaces(n), relative_tol(rtol), tol(rtol) {
    LPM_REQUIRE(rtol > 0);
  }
  void set_tol_from_relative_value() { tol = relative_tol * impl::flag_max(desc()); }
  Index apply(const Index start, const Index end) const { return impl::flag_apply(desc(), flags, start, end); }
  std::string description() const { return "ScalarIntegralFlag"; }
  std::string info_string() const { return impl::flag_info_string(*this); }

 private:
  lpmx_flag_desc_t desc() const {
    lpmx_flag_desc_t d{};
    d.kind = LPMX_FLAG_SCALAR_INTEGRAL, d.n_faces = nfaces, d.face_vals = face_vals.data(), d.area = area.data();
    d.mask = facemask.data(), d.tol = tol;
    return d;
  }
};

/// NV = vertices per face; deduced from the face-vertex view, so `ScalarVariationFlag f(flags, fv, vv, mesh.faces.verts, ...)`
/// reads as in the reference
template <int NV>
struct ScalarVariationFlag {
  flag_view flags;
  scalar_view_type face_vals;
  scalar_view_type vert_vals;
  View2<Index, NV> face_vertex_view;
  mask_view_type facemask;
  Index nfaces;
  Real relative_tol;
  Real tol;
  ScalarVariationFlag(flag_view f, const scalar_view_type fv, const scalar_view_type vv, const View2<Index, NV> verts,
                      const mask_view_type m, const Index n, const Real rtol)
      : flags(f), face_vals(fv), vert_vals(vv), face_vertex_view(verts), facemask(m), nfaces(n), relative_tol(rtol), tol(rtol) {
    LPM_REQUIRE(rtol > 0);
  }
  void set_tol_from_relative_value() { tol = relative_tol * impl::flag_max(desc()); }
  Index apply(const Index start, const Index end) const { return impl::flag_apply(desc(), flags, start, end); }
  std::string description() const { return "ScalarVariationFlag"; }
  std::string info_string() const { return impl::flag_info_string(*this); }

 private:
  lpmx_flag_desc_t desc() const {
    lpmx_flag_desc_t d{};
    d.kind = LPMX_FLAG_SCALAR_VARIATION, d.n_faces = nfaces, d.n_verts = (int)vert_vals.extent(0), d.n_face_verts = NV;
    d.face_vals = face_vals.data(), d.vert_vals = vert_vals.data(), d.face_verts = face_vertex_view.data();
    d.mask = facemask.data(), d.tol = tol;
    return d;
  }
};

template <typename MeshSeedType>
struct FlowMapVariationFlag {
  typedef typename MeshSeedType::geo::crd_view_type crd_view_type;
  static constexpr Int nverts = MeshSeedType::faceKind::nverts;
  flag_view flags;
  crd_view_type vertex_lag_crds;
  View2<Index, MeshSeedType::faceKind::nverts> face_vertex_view;
  mask_view_type facemask;
  Index nfaces;
  Real relative_tol;
  Real tol;
  FlowMapVariationFlag(flag_view f, const PolyMesh2d<MeshSeedType>& mesh, const Real rtol)
      : flags(f), vertex_lag_crds(mesh.vertices.lag_crds.view), face_vertex_view(mesh.faces.verts), facemask(mesh.faces.mask),
        nfaces(mesh.n_faces_host()), relative_tol(rtol), tol(rtol) {
    LPM_REQUIRE(rtol > 0);
  }
  void set_tol_from_relative_value() { tol = relative_tol * impl::flag_max(desc()); }
  Index apply(const Index start, const Index end) const { return impl::flag_apply(desc(), flags, start, end); }
  std::string description() const { return "FlowMapVariationFlag"; }
  std::string info_string() const { return impl::flag_info_string(*this); }

 private:
  lpmx_flag_desc_t desc() const {
    lpmx_flag_desc_t d{};
    d.kind = LPMX_FLAG_FLOW_MAP_VARIATION, d.n_faces = nfaces, d.n_verts = (int)vertex_lag_crds.extent(0), d.n_face_verts = nverts;
    d.face_verts = face_vertex_view.data(), d.vert_lag = vertex_lag_crds.data(), d.ndim = MeshSeedType::geo::ndim;
    d.layout = LPMX_LAYOUT_RIGHT, d.ld = d.n_verts, d.mask = facemask.data(), d.tol = tol;
    return d;
  }
};

/// NeighborsFlag (src/mesh/lpm_refinement_flags.hpp:30-52): a face whose neighbour is more than one level finer is flagged
/// too (2:1 balance).  Host code: it walks the edge tree, which lives in the host generator.
template <typename MeshSeedType>
struct NeighborsFlag {
  flag_view flags;
  const PolyMesh2d<MeshSeedType>& mesh;

  NeighborsFlag(flag_view f, const PolyMesh2d<MeshSeedType>& mesh) : flags(f), mesh(mesh) {}

  std::string description() const { return "NeighborsFlag"; }

  /// flags on in [start, end) after this functor (the convention of the other flags' apply)
  Index apply(const Index start, const Index end) {
    mesh.neighbors_flag(flags, start, end);
    Index n = 0;
    for (Index i = start; i < end; ++i) n += flags(i) ? 1 : 0;
    return n;
  }
};

/// A refinement iteration: clear the flags, run every flag functor over faces [start_idx, end_idx), count what each one
/// added.  Dividing the flagged faces (mesh.divide_flagged_faces(flags, logger)) and setting data on the new particles is
/// the caller's job, as in the reference.
template <typename SeedType>
struct Refinement {
  flag_view flags;
  std::vector<Index> count;
  PolyMesh2d<SeedType>& mesh;

  explicit Refinement(PolyMesh2d<SeedType>& mesh) : flags("refinement_flags", mesh.faces.area.extent(0)), mesh(mesh) {}

  template <typename FlagType>
  void iterate(const Index start_idx, const Index end_idx, FlagType& flagger) {
    clear();
    count = std::vector<Index>(1, 0);
    count[0] = flagger.apply(start_idx, end_idx);
  }
  template <typename FlagType1, typename FlagType2>
  void iterate(const Index start_idx, const Index end_idx, FlagType1& flag1, FlagType2& flag2) {
    clear();
    count = std::vector<Index>(2, 0);
    count[0] = flag1.apply(start_idx, end_idx);
    count[1] = flag2.apply(start_idx, end_idx) - count[0];
  }
  template <typename FlagType1, typename FlagType2, typename FlagType3>
  void iterate(const Index start_idx, const Index end_idx, FlagType1& flag1, FlagType2& flag2, FlagType3& flag3) {
    clear();
    count = std::vector<Index>(3, 0);
    count[0] = flag1.apply(start_idx, end_idx);
    count[1] = flag2.apply(start_idx, end_idx) - count[0];
    count[2] = flag3.apply(start_idx, end_idx) - (count[0] + count[1]);
  }

 private:
  void clear() { std::fill(flags.data(), flags.data() + flags.extent(0), (unsigned char)0); }
};

}  // namespace Lpm
#endif
