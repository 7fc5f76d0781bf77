// lpm/lpm_ftle.hpp -- ComputeFTLE<SeedType> and get_max_ftle over lpmx_ftle.
//   ComputeFTLE<SeedType>   src/mesh/lpm_ftle.hpp:15-319 (same constructor arguments; quadrilateral faces only)
//   get_max_ftle            src/mesh/lpm_ftle.hpp:327-338
// The reference launches the functor with Kokkos::parallel_for(n_faces, ComputeFTLE<seed>(...))
// (examples/sphere_rh54.cpp:308-316); here the same object is run with .apply(n_faces), one kernel launch.
#ifndef LPM_SHIM_FTLE_HPP
#define LPM_SHIM_FTLE_HPP

#include <type_traits>

#include "lpm_polymesh2d.hpp"

namespace Lpm {

template <typename SeedType>
struct ComputeFTLE {
  using face_kind = typename SeedType::faceKind;
  using geo = typename SeedType::geo;
  using crd_view = typename geo::crd_view_type;
  using face_vertex_view = View2<Index, face_kind::nverts>;
  static_assert(std::is_same<face_kind, QuadFace>::value, "FTLE for non-quadrilateral faces not implemented yet.");

  scalar_view_type ftle;
  crd_view phys_crds_verts, ref_crds_verts, phys_crds_faces, ref_crds_faces;
  face_vertex_view face_verts;
  mask_view_type face_mask;
  Real t;
  Real max_ftle = 0;  ///< filled by apply(): get_max_ftle of the same launch

  ComputeFTLE(scalar_view_type ftle, const crd_view phys_crds_verts, const crd_view ref_crds_verts,
              const crd_view phys_crds_faces, const crd_view ref_crds_faces, const face_vertex_view face_verts,
              const mask_view_type face_mask, const Real& time_since_ref)
      : ftle(ftle), phys_crds_verts(phys_crds_verts), ref_crds_verts(ref_crds_verts), phys_crds_faces(phys_crds_faces),
        ref_crds_faces(ref_crds_faces), face_verts(face_verts), face_mask(face_mask), t(time_since_ref) {}

  /// all faces [0, n_faces): writes ftle at the leaves and, on the sphere, normalises phys_crds_faces of the leaves
  void apply(const Index n_faces) {
    const Index n_verts = (Index)phys_crds_verts.extent(0);
    const int g = std::is_same<geo, SphereGeometry>::value ? LPMX_GEOM_SPHERE : LPMX_GEOM_PLANE;
    Engine::check(lpmx_ftle(Engine::get(), g, n_verts, phys_crds_verts.data(), ref_crds_verts.data(), LPMX_LAYOUT_RIGHT, 0,
                            n_faces, phys_crds_faces.data(), ref_crds_faces.data(), 0, face_verts.data(), LPMX_LAYOUT_RIGHT,
                            face_mask.data(), ftle.data(), &max_ftle),
                  "ComputeFTLE");
  }
};

inline Real get_max_ftle(const scalar_view_type ftle, const mask_view_type mask, const Index& nfaces) {
  Real result = std::numeric_limits<Real>::lowest();
  for (Index i = 0; i < nfaces; ++i)
    if (!mask(i)) result = (result > ftle(i) ? result : ftle(i));
  return result;
}

}  // namespace Lpm
#endif
