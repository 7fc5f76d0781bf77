// ref_mesh_driver.cpp -- C entry points around the REFERENCE's own mesh classes and BVE stepper, compiled in place from
// /root/reference/src (never copied) against oracle/kokkos_shim.  Output: oracle/_ref/liblpm_ref_mesh.so.
// TEST INFRASTRUCTURE: generates tests/golden/mesh_*.npz / mesh_amr_*.npz / ref_bve_rk4.npz (tests/golden/make_mesh_golden.py,
// make_amr_golden.py, make_ref_stepper_golden.py) and backs the live comparisons in tests/test_mesh.py /
// tests/test_oracle_golden.py when the library is present.  Nothing under lpm_b200/ or include/ loads it.
//
// Reference code exercised (all as shipped; the translation units are compiled from where they lie, see oracle/Makefile):
//   mesh/lpm_polymesh2d{.hpp,_impl.hpp,.cpp}   PolyMesh2d<Seed>::tree_init, divide_flagged_faces
//   mesh/lpm_faces{.hpp,_impl.hpp,.cpp}        Faces, FaceDivider<Geo, TriFace|QuadFace>::divide, scan_leaves
//   mesh/lpm_edges.{hpp,cpp}                   Edges::divide, insert_host
//   mesh/lpm_vertices{.hpp,_impl.hpp,.cpp}, lpm_coords{.hpp,_impl.hpp,.cpp}, mesh/lpm_mesh_seed.{hpp,cpp}
//   lpm_bve_sphere{.hpp,_impl.hpp}             BVESphere<Seed>::init_velocity, init_stream_fn
//   lpm_bve_rk4{.hpp,_impl.hpp,.cpp}           BVERK4::advance_timestep
// The only stand-ins are oracle/kokkos_shim/{Kokkos_Core.hpp, KokkosBlas.hpp, mpi.h, LpmConfig.h, compose/siqk_sqr.hpp}:
// siqk::sqr::calc_sphere_to_ref (COMPOSE, absent) is referenced by PolyMesh2d::quad_ref only, which nothing here calls.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "lpm_bve_rk4.hpp"
#include "lpm_bve_rk4_impl.hpp"
#include "lpm_bve_sphere.hpp"
#include "lpm_bve_sphere_impl.hpp"
#include "lpm_logger.hpp"
#include "mesh/lpm_mesh_seed.hpp"
#include "mesh/lpm_polymesh2d.hpp"
#include "mesh/lpm_polymesh2d_impl.hpp"

using namespace Lpm;

namespace {

struct MeshBase {
  virtual ~MeshBase() {}
  virtual void counts(int* c) = 0;
  virtual void get(double* vx, double* vlx, int* eo, int* ed, int* el, int* er, int* ep, int* ek, double* fx, double* flx,
                   double* fa, uint8_t* fm, int* fv, int* fe, int* fp, int* fk, int* flev, int* fleaf, int* fcrd) = 0;
  virtual int divide_flagged(const uint8_t* flags, int n, int* refine_count) = 0;
};

// a logger that records what divide_flagged_faces reported, so the three outcomes (all divided / not enough memory /
// level limit reached) can be told apart without parsing text
struct OutcomeLogger {
  int outcome = 0;  // 0 all divided, 1 not enough memory, 2 level limit reached
  template <class... A>
  void debug(const A&...) {}
  template <class... A>
  void info(const A&...) {}
  template <class... A>
  void warn(const std::string& msg, const A&...) {
    if (msg.find("not enough memory") != std::string::npos) outcome = 1;
    if (msg.find("limit reached") != std::string::npos) outcome = 2;
  }
  template <class... A>
  void error(const A&...) {}
};

template <class Seed>
struct MeshT : MeshBase {
  static constexpr int NV = Seed::faceKind::nverts;
  static constexpr int ND = Seed::geo::ndim;
  PolyMeshParameters<Seed> params;
  std::unique_ptr<PolyMesh2d<Seed>> mesh;
  MeshT(int depth, double radius, int amr_buffer, int amr_limit) : params(depth, radius, amr_buffer, amr_limit) {
    mesh = std::make_unique<PolyMesh2d<Seed>>(params);  // tree_init(depth, seed) + update_device()
  }
  void counts(int* c) override {
    c[0] = mesh->n_vertices_host();
    c[1] = mesh->edges.nh();
    c[2] = mesh->n_faces_host();
    c[3] = NV;
    c[4] = ND;
    c[5] = mesh->faces.n_leaves_host();
    c[6] = params.nmaxverts;
    c[7] = params.nmaxedges;
    c[8] = params.nmaxfaces;
  }
  void get(double* vx, double* vlx, int* eo, int* ed, int* el, int* er, int* ep, int* ek, double* fx, double* flx, double* fa,
           uint8_t* fm, int* fv, int* fe, int* fp, int* fk, int* flev, int* fleaf, int* fcrd) override {
    mesh->update_host();
    const int nv = mesh->n_vertices_host(), ne = mesh->edges.nh(), nf = mesh->n_faces_host();
    const auto hv = mesh->vertices.phys_crds.get_const_host_crd_view();
    const auto hlv = mesh->vertices.lag_crds.get_const_host_crd_view();
    for (int i = 0; i < nv; ++i)
      for (int k = 0; k < ND; ++k) {
        vx[ND * i + k] = hv(i, k);
        vlx[ND * i + k] = hlv(i, k);
      }
    for (int i = 0; i < ne; ++i) {
      eo[i] = mesh->edges.orig_host(i);
      ed[i] = mesh->edges.dest_host(i);
      el[i] = mesh->edges.left_host(i);
      er[i] = mesh->edges.right_host(i);
      ep[i] = mesh->edges.parent_host(i);
      ek[2 * i] = mesh->edges.kid_host(i, 0);
      ek[2 * i + 1] = mesh->edges.kid_host(i, 1);
    }
    const auto hf = mesh->faces.phys_crds.get_const_host_crd_view();
    const auto hlf = mesh->faces.lag_crds.get_const_host_crd_view();
    const auto hmask = mesh->faces.leaf_mask_host();
    const auto hlevel = mesh->faces.levels_host();
    auto hleaf = Kokkos::create_mirror_view(mesh->faces.leaf_idx);
    Kokkos::deep_copy(hleaf, mesh->faces.leaf_idx);
    for (int i = 0; i < nf; ++i) {
      for (int k = 0; k < ND; ++k) {
        fx[ND * i + k] = hf(i, k);
        flx[ND * i + k] = hlf(i, k);
      }
      fa[i] = mesh->faces.area_host(i);
      fm[i] = hmask(i) ? 1 : 0;
      for (int k = 0; k < NV; ++k) {
        fv[NV * i + k] = mesh->faces.vert_host(i, k);
        fe[NV * i + k] = mesh->faces.edge_host(i, k);
      }
      fp[i] = mesh->faces.parent_host(i);
      for (int k = 0; k < 4; ++k) fk[4 * i + k] = mesh->faces.kid_host(i, k);
      flev[i] = hlevel(i);
      fleaf[i] = hleaf(i);
      fcrd[i] = mesh->faces.crd_idx_host(i);
    }
  }
  int divide_flagged(const uint8_t* flags, int n, int* refine_count) override {
    const int nf0 = mesh->n_faces_host();
    Kokkos::View<bool*> f("flags", params.nmaxfaces);
    for (int i = 0; i < n && i < (int)params.nmaxfaces; ++i) f(i) = flags[i] != 0;
    OutcomeLogger lg;
    mesh->divide_flagged_faces(f, lg);
    mesh->update_device();
    *refine_count = (mesh->n_faces_host() - nf0) / 4;
    return lg.outcome;
  }
};

}  // namespace

extern "C" {

// seed: 0 icos (IcosTriSphereSeed), 1 cubed (CubedSphereSeed), 2 quad_rect (QuadRectSeed), 3 tri_hex (TriHexSeed)
void* ref_mesh_create(int seed, int depth, double radius, int amr_buffer, int amr_limit) {
  switch (seed) {
    case 0: return new MeshT<IcosTriSphereSeed>(depth, radius, amr_buffer, amr_limit);
    case 1: return new MeshT<CubedSphereSeed>(depth, radius, amr_buffer, amr_limit);
    case 2: return new MeshT<QuadRectSeed>(depth, radius, amr_buffer, amr_limit);
    case 3: return new MeshT<TriHexSeed>(depth, radius, amr_buffer, amr_limit);
  }
  return nullptr;
}
void ref_mesh_destroy(void* m) { delete static_cast<MeshBase*>(m); }
// c[9] = {n_verts, n_edges, n_faces, verts per face, ndim, n_leaves, nmaxverts, nmaxedges, nmaxfaces}
void ref_mesh_counts(void* m, int* c) { static_cast<MeshBase*>(m)->counts(c); }
void ref_mesh_get(void* m, double* vx, double* vlx, int* eo, int* ed, int* el, int* er, int* ep, int* ek, double* fx, double* flx,
                  double* fa, uint8_t* fm, int* fv, int* fe, int* fp, int* fk, int* flev, int* fleaf, int* fcrd) {
  static_cast<MeshBase*>(m)->get(vx, vlx, eo, ed, el, er, ep, ek, fx, flx, fa, fm, fv, fe, fp, fk, flev, fleaf, fcrd);
}
// PolyMesh2d::divide_flagged_faces; returns the outcome (0 all divided, 1 not enough memory, 2 level limit reached)
int ref_mesh_divide_flagged(void* m, const uint8_t* flags, int n, int* refine_count) {
  return static_cast<MeshBase*>(m)->divide_flagged(flags, n, refine_count);
}

int ref_mesh_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void ref_mesh_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

}  // extern "C"

// ---- BVESphere + BVERK4 as shipped ---------------------------------------------------------------------------------
namespace {
template <class Seed>
int bve_rk4_run(int depth, double dt, double omega, int n_steps, const double* vert_zeta, const double* face_zeta, double* vx,
                double* vz, double* vu, double* vpsi, double* fx, double* fz, double* fu, double* fpsi, int do_psi) {
  MeshSeed<Seed> seed;
  Index nmaxverts, nmaxedges, nmaxfaces;
  seed.set_max_allocations(nmaxverts, nmaxedges, nmaxfaces, depth);
  auto sphere = std::make_unique<BVESphere<Seed>>(nmaxverts, nmaxedges, nmaxfaces, 0);
  sphere->tree_init(depth, seed);
  sphere->update_device();
  sphere->set_omega(omega);
  const int nv = sphere->n_vertices_host(), nf = sphere->n_faces_host();
  // the caller's relative vorticity (the same arrays the engine under test gets); absolute vorticity as init_vorticity does
  for (int i = 0; i < nv; ++i) {
    sphere->rel_vort_verts.view(i) = vert_zeta[i];
    sphere->abs_vort_verts.view(i) = vert_zeta[i] + 2 * omega * sphere->vertices.phys_crds.view(i, 2);
  }
  for (int i = 0; i < nf; ++i) {
    sphere->rel_vort_faces.view(i) = face_zeta[i];
    sphere->abs_vort_faces.view(i) = face_zeta[i] + 2 * omega * sphere->faces.phys_crds.view(i, 2);
  }
  sphere->init_velocity();
  if (n_steps > 0) {
    BVERK4 solver(dt, *sphere);
    for (int s = 0; s < n_steps; ++s) solver.advance_timestep(*sphere);
  }
  if (do_psi) sphere->init_stream_fn();
  for (int i = 0; i < nv; ++i) {
    for (int k = 0; k < 3; ++k) {
      vx[3 * i + k] = sphere->vertices.phys_crds.view(i, k);
      vu[3 * i + k] = sphere->velocity_verts.view(i, k);
    }
    vz[i] = sphere->rel_vort_verts.view(i);
    if (do_psi) vpsi[i] = sphere->stream_fn_verts.view(i);
  }
  for (int i = 0; i < nf; ++i) {
    for (int k = 0; k < 3; ++k) {
      fx[3 * i + k] = sphere->faces.phys_crds.view(i, k);
      fu[3 * i + k] = sphere->velocity_faces.view(i, k);
    }
    fz[i] = sphere->rel_vort_faces.view(i);
    if (do_psi) fpsi[i] = sphere->stream_fn_faces.view(i);
  }
  return 0;
}
}  // namespace

extern "C" {
// BVESphere<Seed>(depth) with the given relative vorticity -> init_velocity -> n_steps x BVERK4::advance_timestep
// [-> init_stream_fn].  Outputs sized for the mesh (ref_mesh_counts of the same seed/depth).  seed: 0 icos, 1 cubed.
int ref_bve_rk4_run(int seed, int depth, double dt, double omega, int n_steps, const double* vert_zeta, const double* face_zeta,
                    double* vx, double* vz, double* vu, double* vpsi, double* fx, double* fz, double* fu, double* fpsi,
                    int do_psi) {
  if (seed == 0)
    return bve_rk4_run<IcosTriSphereSeed>(depth, dt, omega, n_steps, vert_zeta, face_zeta, vx, vz, vu, vpsi, fx, fz, fu, fpsi,
                                          do_psi);
  if (seed == 1)
    return bve_rk4_run<CubedSphereSeed>(depth, dt, omega, n_steps, vert_zeta, face_zeta, vx, vz, vu, vpsi, fx, fz, fu, fpsi,
                                        do_psi);
  return -1;
}
}
