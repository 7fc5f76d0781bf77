// ref_flags_driver.cpp -- C entry points around the REFERENCE's own refinement-flag functors
// (/root/reference/src/mesh/lpm_refinement_flags.hpp, compiled in place, never copied) against oracle/kokkos_shim.
// Linked into oracle/_ref/liblpm_ref.so.  TEST INFRASTRUCTURE: pins oracle/refinement_oracle.py and generates
// tests/golden/ref_flags.npz (tests/golden/make_ref_flags_golden.py).
//
// The header only needs PolyMesh2d<Seed> as the constructor argument of FlowMapVariationFlag (four members are read:
// vertices.lag_crds.view, faces.verts, faces.mask, n_faces_host()); the reference's own PolyMesh2d cannot be compiled
// here (COMPOSE absent), so a four-member stand-in of that name is declared before the header is included.  The functor
// bodies -- set_tol_from_relative_value() and operator() -- are the reference's, as shipped.
//
// Refinement<Seed>::iterate (mesh/lpm_refinement.hpp:28-41) includes the mesh class as well; its three statements
// (clear the flags, run the functor over [start, end), count) are replayed by ref_flag_iterate below.
#include <cstdint>
#include <memory>

#include "LpmConfig.h"
#include "lpm_geometry.hpp"
#include "lpm_kokkos_defs.hpp"
#include "mesh/lpm_mesh_seed.hpp"

namespace Lpm {
template <typename SeedType>
struct PolyMesh2d {
  struct V {
    struct C {
      typename SeedType::geo::crd_view_type view;
    } lag_crds;
  } vertices;
  struct F {
    Kokkos::View<Index**> verts;
    mask_view_type mask;
  } faces;
  Index nf;
  Index n_faces_host() const { return nf; }
};
}  // namespace Lpm

#include "mesh/lpm_refinement_flags.hpp"

using namespace Lpm;

namespace {
struct Bools {
  std::unique_ptr<bool[]> b;
  Kokkos::View<bool*> v;
  Bools(const uint8_t* m, int n) : b(new bool[n > 0 ? n : 1]) {
    for (int i = 0; i < n; ++i) b[i] = m[i] != 0;
    v = Kokkos::View<bool*>(b.get(), n);
  }
  void store(uint8_t* m, int n) const {
    for (int i = 0; i < n; ++i) m[i] = b[i] ? 1 : 0;
  }
};
inline scalar_view_type wrap1(const double* p, int n) { return scalar_view_type(const_cast<double*>(p), n); }

// Refinement::iterate with one flag functor; flags cleared first, count returned
template <typename Flag>
int iterate(Flag& flag, Bools& flags, int n_flags, int start, int end) {
  for (int i = 0; i < n_flags; ++i) flags.b[i] = false;
  Kokkos::parallel_for(Kokkos::RangePolicy<>(start, end), flag);
  int ct = 0;
  for (int i = start; i < end; ++i) ct += flags.b[i] ? 1 : 0;
  return ct;
}

template <typename Seed>
int flow_map(int n_verts, const double* vert_lag, int n_faces, const int* face_verts, const uint8_t* mask, double rtol,
             int relative, int start, int end, uint8_t* flags_out, double* tol_out) {
  Bools fl(flags_out, n_faces), fm(mask, n_faces);
  PolyMesh2d<Seed> mesh;
  mesh.vertices.lag_crds.view = typename Seed::geo::crd_view_type(const_cast<double*>(vert_lag), n_verts);
  mesh.faces.verts = Kokkos::View<Index**>(const_cast<int*>(face_verts), n_faces, Seed::faceKind::nverts);
  mesh.faces.mask = fm.v;
  mesh.nf = n_faces;
  FlowMapVariationFlag<Seed> flag(fl.v, mesh, rtol);
  if (relative) flag.set_tol_from_relative_value();
  *tol_out = flag.tol;
  const int ct = iterate(flag, fl, n_faces, start, end);
  fl.store(flags_out, n_faces);
  return ct;
}
}  // namespace

extern "C" {

// kind 0: ScalarMaxFlag, 1: ScalarIntegralFlag, 2: ScalarVariationFlag (nfv vertices per face)
int ref_flag_scalar(int kind, int n_faces, const double* face_vals, const double* area, int n_verts,
                    const double* vert_vals, const int* face_verts, int nfv, const uint8_t* mask, double rtol,
                    int relative, int start, int end, uint8_t* flags_out, double* tol_out) {
  Bools fl(flags_out, n_faces), fm(mask, n_faces);
  int ct = -1;
  if (kind == 0) {
    ScalarMaxFlag flag(fl.v, wrap1(face_vals, n_faces), fm.v, n_faces, rtol);
    if (relative) flag.set_tol_from_relative_value();
    *tol_out = flag.tol;
    ct = iterate(flag, fl, n_faces, start, end);
  } else if (kind == 1) {
    ScalarIntegralFlag flag(fl.v, wrap1(face_vals, n_faces), wrap1(area, n_faces), fm.v, n_faces, rtol);
    if (relative) flag.set_tol_from_relative_value();
    *tol_out = flag.tol;
    ct = iterate(flag, fl, n_faces, start, end);
  } else if (kind == 2) {
    Kokkos::View<Index**> fv(const_cast<int*>(face_verts), n_faces, nfv);
    ScalarVariationFlag flag(fl.v, wrap1(face_vals, n_faces), wrap1(vert_vals, n_verts), fv, fm.v, n_faces, rtol);
    if (relative) flag.set_tol_from_relative_value();
    *tol_out = flag.tol;
    ct = iterate(flag, fl, n_faces, start, end);
  }
  fl.store(flags_out, n_faces);
  return ct;
}

// FlowMapVariationFlag<IcosTriSphereSeed> (nfv == 3) / <CubedSphereSeed> (nfv == 4)
int ref_flag_flow_map(int nfv, int n_verts, const double* vert_lag, int n_faces, const int* face_verts,
                      const uint8_t* mask, double rtol, int relative, int start, int end, uint8_t* flags_out,
                      double* tol_out) {
  if (nfv == 3)
    return flow_map<IcosTriSphereSeed>(n_verts, vert_lag, n_faces, face_verts, mask, rtol, relative, start, end, flags_out,
                                       tol_out);
  return flow_map<CubedSphereSeed>(n_verts, vert_lag, n_faces, face_verts, mask, rtol, relative, start, end, flags_out, tol_out);
}

}  // extern "C"
