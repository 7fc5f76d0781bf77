// ref_driver.cpp -- C entry points around the REFERENCE's own direct-sum functors, compiled in place from
// /root/reference/src (never copied) against oracle/kokkos_shim.  Output: oracle/_ref/liblpm_ref.so.
// TEST INFRASTRUCTURE: used by tests/test_ref_build.py to pin oracle/lpm_oracle.c (same C signatures, so
// either library can sit behind oracle/oracle.py), and optionally as the timed CPU baseline
// (bench.py --impl reference, cpu_baseline.kind == "reference").
//
// Reference code exercised (all as shipped):
//   lpm_sphere_functions.hpp   greens_fn, biot_savart
//   lpm_bve_sphere_kernels.hpp BVEVertexVelocity, BVEFaceVelocity, BVEVertexStreamFn, BVEFaceStreamFn,
//                              BVEVorticityTendency
//   lpm_incompressible2d_kernels.hpp  Incompressible2DPassiveSums / ActiveSums <SphereGeometry>
//   lpm_swe_kernels.hpp        kzeta_sphere, ksigma_sphere, grad_kzeta, grad_ksigma, SphereVertexSums,
//                              SphereFaceSums
//   util/lpm_matlab_io.hpp     write_vector_matlab, write_array_matlab
//   mesh/lpm_ftle.hpp          ComputeFTLE<CubedSphereSeed>, ComputeFTLE<QuadRectSeed>, get_max_ftle
#include <cstdint>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "lpm_coriolis.hpp"  // lpm_incompressible2d_kernels.hpp relies on its includer for this
#include "lpm_bve_sphere_kernels.hpp"
#include "lpm_incompressible2d_kernels.hpp"
#include "lpm_swe_kernels.hpp"
#include "lpm_surface_gallery.hpp"
#include "mesh/lpm_ftle.hpp"
#include "util/lpm_matlab_io.hpp"
#include <sstream>

using namespace Lpm;
using crd = SphereGeometry::crd_view_type;
using vec = SphereGeometry::vec_view_type;

namespace {
struct Mask {
  std::unique_ptr<bool[]> b;
  mask_view_type v;
  Mask(const uint8_t* m, int n) : b(new bool[n > 0 ? n : 1]) {
    for (int i = 0; i < n; ++i) b[i] = m[i] != 0;
    v = mask_view_type(b.get(), n);
  }
};
inline crd wrap3(const double* p, int n) { return crd(const_cast<double*>(p), n); }
inline scalar_view_type wrap1(const double* p, int n) { return scalar_view_type(const_cast<double*>(p), n); }
}  // namespace

extern "C" {

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the timed CPU legs set the team size explicitly */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_bve_velocity(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                         const double* area, const uint8_t* mask, int collocated, double* vel) {
  Mask fm(mask, n_src);
  crd fx = wrap3(sx, n_src);
  scalar_view_type fz = wrap1(zeta, n_src), fa = wrap1(area, n_src);
  vec u = wrap3(vel, n_tgt);
  if (collocated) {
    Kokkos::parallel_for("ref face velocity", Kokkos::TeamPolicy<>(n_src, Kokkos::AUTO()),
                         BVEFaceVelocity(u, fx, fz, fa, fm.v, n_src));
  } else {
    crd vx = wrap3(tx, n_tgt);
    Kokkos::parallel_for("ref vertex velocity", Kokkos::TeamPolicy<>(n_tgt, Kokkos::AUTO()),
                         BVEVertexVelocity(u, vx, fx, fz, fa, fm.v, n_src));
  }
}

// BVEFaceVelocity (the collocated functor, lpm_bve_sphere_kernels.hpp:365-394) run for a subset of its league: the functor is
// the reference's, only the set of league ranks it is called with is chosen here.  Row k of vel = velocity at particle idx[k].
void oracle_bve_velocity_subset(int n_idx, const int* idx, int n_src, const double* sx, const double* zeta, const double* area,
                                const uint8_t* mask, double* vel) {
  Mask fm(mask, n_src);
  crd fx = wrap3(sx, n_src);
  scalar_view_type fz = wrap1(zeta, n_src), fa = wrap1(area, n_src);
  vec u("subset velocity", n_src);
  BVEFaceVelocity f(u, fx, fz, fa, fm.v, n_src);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < n_idx; ++k) f(Kokkos::TeamMember{idx[k]});
  for (int k = 0; k < n_idx; ++k)
    for (int c = 0; c < 3; ++c) vel[3 * k + c] = u(idx[k], c);
}

void oracle_bve_streamfn(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                         const double* area, const uint8_t* mask, int collocated, double* psi) {
  Mask fm(mask, n_src);
  crd fx = wrap3(sx, n_src);
  scalar_view_type fz = wrap1(zeta, n_src), fa = wrap1(area, n_src);
  scalar_view_type p = wrap1(psi, n_tgt);
  if (collocated) {
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_src, Kokkos::AUTO()), BVEFaceStreamFn(p, fx, fz, fa, fm.v, n_src));
  } else {
    crd vx = wrap3(tx, n_tgt);
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_tgt, Kokkos::AUTO()), BVEVertexStreamFn(p, vx, fx, fz, fa, fm.v, n_src));
  }
}

void oracle_ic2d_sums(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                      const double* area, const uint8_t* mask, double eps, int targets_are_sources, double* vel,
                      double* psi) {
  Mask am(mask, n_src);
  crd ay = wrap3(sx, n_src);
  scalar_view_type az = wrap1(zeta, n_src), aa = wrap1(area, n_src);
  vec u = wrap3(vel, n_tgt);
  std::vector<double> scratch;
  if (!psi) {
    scratch.resize(n_tgt > 0 ? n_tgt : 1);
    psi = scratch.data();
  }
  scalar_view_type p = wrap1(psi, n_tgt);
  if (targets_are_sources) {
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_src, Kokkos::AUTO()),
                         Incompressible2DActiveSums<SphereGeometry>(u, p, ay, az, aa, am.v, eps, n_src));
  } else {
    crd px = wrap3(tx, n_tgt);
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_tgt, Kokkos::AUTO()),
                         Incompressible2DPassiveSums<SphereGeometry>(u, p, px, ay, az, aa, am.v, eps, n_src));
  }
}

void oracle_kzeta_sphere(double* u, const double* x, const double* y, double vort, double area, double eps) {
  kzeta_sphere(u, x, y, vort, area, eps);
}
void oracle_ksigma_sphere(double* u, const double* x, const double* y, double div, double area, double eps) {
  ksigma_sphere(u, x, y, div, area, eps);
}
void oracle_grad_kzeta(double* g, const double* x, const double* y, double eps) { grad_kzeta(g, x, y, eps); }
void oracle_grad_ksigma(double* g, const double* x, const double* y, double eps) { grad_ksigma(g, x, y, eps); }

// SphereVertexSums / SphereFaceSums.  The reference leaves velz/vels uninitialised inside
// sphere_swe_velocity_sums (undefined behaviour, SURVEY.md quirk C-i), so the VELOCITY part of this entry
// point is whatever the compiler made of that; ddot / gradient sums are well defined.  grad9 is not
// exposed by the reference functors and is left untouched here.
void oracle_swe_sphere_sums(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                            const double* sigma, const double* area, const uint8_t* mask, double eps,
                            int targets_are_sources, int do_velocity, double* vel, double* ddot, double* grad9) {
  (void)grad9;
  Mask fm(mask, n_src);
  crd fy = wrap3(sx, n_src);
  scalar_view_type fz = wrap1(zeta, n_src), fs = wrap1(sigma, n_src), fa = wrap1(area, n_src);
  vec u = wrap3(vel, n_tgt);
  scalar_view_type dd = wrap1(ddot, n_tgt);
  if (targets_are_sources) {
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_src, Kokkos::AUTO()),
                         SphereFaceSums(u, dd, fy, fz, fs, fa, fm.v, eps, n_src, do_velocity != 0));
  } else {
    crd vx = wrap3(tx, n_tgt);
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_tgt, Kokkos::AUTO()),
                         SphereVertexSums(u, dd, vx, fy, fz, fs, fa, fm.v, eps, n_src, do_velocity != 0));
  }
}

// SWEVorticityDivergenceHeightTendencies / ...AreaTendencies <SphereGeometry>, CoriolisSphere(Omega)
void oracle_swe_tendencies(int n, int is_area, double* dzeta, double* dsigma, double* dthird, const double* x,
                           const double* u, const double* zeta, const double* sigma, const double* third,
                           const double* ddot, const double* laps, double Omega, double g, double dt) {
  scalar_view_type dz = wrap1(dzeta, n), ds = wrap1(dsigma, n), d3 = wrap1(dthird, n);
  crd xv = wrap3(x, n);
  vec uv = wrap3(u, n);
  CoriolisSphere cor(Omega);
  if (is_area) {
    Kokkos::parallel_for(n, SWEVorticityDivergenceAreaTendencies<SphereGeometry>(
                                dz, ds, d3, xv, uv, wrap1(zeta, n), wrap1(sigma, n), wrap1(third, n), wrap1(ddot, n),
                                wrap1(laps, n), cor, g, dt));
  } else {
    Kokkos::parallel_for(n, SWEVorticityDivergenceHeightTendencies<SphereGeometry>(
                                dz, ds, d3, xv, uv, wrap1(zeta, n), wrap1(sigma, n), wrap1(third, n), wrap1(ddot, n),
                                wrap1(laps, n), cor, g, dt));
  }
}

// SetSurfaceFromDepth<SphereGeometry, ZeroFunctor>; x is not needed by ZeroFunctor but the functor reads it
void ref_swe_set_surface_from_depth(int n, double* s, double* b, const double* x, const double* h) {
  scalar_view_type sv = wrap1(s, n), bv = wrap1(b, n);
  Kokkos::parallel_for(n, SetSurfaceFromDepth<SphereGeometry, ZeroFunctor>(sv, bv, wrap3(x, n), wrap1(h, n),
                                                                            ZeroFunctor()));
}

// SetDepthAndSurfaceFromMassAndArea<SphereGeometry, ZeroFunctor>
void ref_swe_set_depth_surface_from_mass_area(int n, double* h, double* s, double* b, const double* x,
                                              const double* m, const double* area, const uint8_t* mask) {
  Mask fm(mask, n);
  scalar_view_type hv = wrap1(h, n), sv = wrap1(s, n), bv = wrap1(b, n);
  Kokkos::parallel_for(n, SetDepthAndSurfaceFromMassAndArea<SphereGeometry, ZeroFunctor>(
                              hv, sv, bv, wrap3(x, n), wrap1(m, n), wrap1(area, n), fm.v, ZeroFunctor()));
}

// BVEVorticityTendency over n particles (O(N); used to pin the oracle's stage algebra)
void ref_bve_vorticity_tendency(int n, double* dzeta, const double* vel, double dt, double Omega) {
  scalar_view_type dz = wrap1(dzeta, n);
  vec u = wrap3(vel, n);
  Kokkos::parallel_for(n, BVEVorticityTendency(dz, u, dt, Omega));
}

// ---------------------------------------------------------------------------------------------------------
// Planar paths: the reference's PlaneGeometry functors, as shipped.
//   lpm_incompressible2d_kernels.hpp  Incompressible2DPassiveSums / ActiveSums / Tendencies <PlaneGeometry>
//   lpm_swe_kernels.hpp               planar_swe_sums_rhs_pse, PlanarSWEVertexSums, PlanarSWEFaceSums,
//                                     SWEVorticityDivergence{Height,Area}Tendencies<PlaneGeometry>,
//                                     SetSurfaceFromDepth / SetDepthAndSurfaceFromMassAndArea <PlaneGeometry, Topo>
//   lpm_surface_gallery.hpp           PlanarGaussianMountain, ZeroFunctor
// ---------------------------------------------------------------------------------------------------------
using crd2 = PlaneGeometry::crd_view_type;
using vec2 = PlaneGeometry::vec_view_type;
namespace {
inline crd2 wrap2(const double* p, int n) { return crd2(const_cast<double*>(p), n); }
}

void oracle_ic2d_plane_sums(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                            const double* area, const uint8_t* mask, double eps, int targets_are_sources, double* vel,
                            double* psi) {
  Mask am(mask, n_src);
  crd2 ay = wrap2(sx, n_src);
  scalar_view_type az = wrap1(zeta, n_src), aa = wrap1(area, n_src);
  vec2 u = wrap2(vel, n_tgt);
  std::vector<double> scratch;
  if (!psi) {
    scratch.resize(n_tgt > 0 ? n_tgt : 1);
    psi = scratch.data();
  }
  scalar_view_type p = wrap1(psi, n_tgt);
  if (targets_are_sources) {
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_src, Kokkos::AUTO()),
                         Incompressible2DActiveSums<PlaneGeometry>(u, p, ay, az, aa, am.v, eps, n_src));
  } else {
    crd2 px = wrap2(tx, n_tgt);
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_tgt, Kokkos::AUTO()),
                         Incompressible2DPassiveSums<PlaneGeometry>(u, p, px, ay, az, aa, am.v, eps, n_src));
  }
}

// Incompressible2DTendencies<PlaneGeometry> with CoriolisBetaPlane(f0, beta)
void ref_ic2d_plane_tendency(int n, double* dzeta, const double* vel, double f0, double beta) {
  CoriolisBetaPlane cor(f0, beta);
  Kokkos::parallel_for(n, Incompressible2DTendencies<PlaneGeometry>(wrap1(dzeta, n), wrap2(vel, n), cor));
}

void oracle_planar_swe_sums_rhs_pse(double* result, const double* tgt_x, const double* src_y, double src_zeta,
                                    double src_sigma, double src_area, double src_s, double tgt_s, double eps,
                                    double pse_eps) {
  const auto r = planar_swe_sums_rhs_pse(tgt_x, src_y, src_zeta, src_sigma, src_area, src_s, tgt_s, eps, pse_eps);
  for (int k = 0; k < 9; ++k) result[k] = r[k];
}

void oracle_swe_plane_sums(int n_tgt, const double* tx, const double* tsurf, int n_src, const double* sx,
                           const double* zeta, const double* sigma, const double* area, const uint8_t* mask,
                           const double* ssurf, double eps, double pse_eps, int targets_are_sources, int do_velocity,
                           double* vel, double* ddot, double* du1dx1, double* du1dx2, double* du2dx1, double* du2dx2,
                           double* laps, double* psi, double* phi) {
  Mask fm(mask, n_src);
  std::vector<double> scratch((size_t)8 * (n_tgt > 0 ? n_tgt : 1));
  double* outs[8] = {ddot, du1dx1, du1dx2, du2dx1, du2dx2, laps, psi, phi};
  scalar_view_type ov[8];
  for (int k = 0; k < 8; ++k) ov[k] = wrap1(outs[k] ? outs[k] : scratch.data() + (size_t)k * n_tgt, n_tgt);
  vec2 u = wrap2(vel, n_tgt);
  crd2 fy = wrap2(sx, n_src);
  scalar_view_type fz = wrap1(zeta, n_src), fs = wrap1(sigma, n_src), fa = wrap1(area, n_src), fsf = wrap1(ssurf, n_src);
  if (targets_are_sources) {
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_src, Kokkos::AUTO()),
                         PlanarSWEFaceSums(u, ov[0], ov[1], ov[2], ov[3], ov[4], ov[5], ov[6], ov[7], fy, fz, fs, fa,
                                           fm.v, fsf, eps, pse_eps, n_src, do_velocity != 0));
  } else {
    Kokkos::parallel_for(Kokkos::TeamPolicy<>(n_tgt, Kokkos::AUTO()),
                         PlanarSWEVertexSums(u, ov[0], ov[1], ov[2], ov[3], ov[4], ov[5], ov[6], ov[7], wrap2(tx, n_tgt),
                                             wrap1(tsurf, n_tgt), fy, fz, fs, fa, fm.v, fsf, eps, pse_eps, n_src,
                                             do_velocity != 0));
  }
}

void oracle_swe_plane_tendencies(int n, int is_area, double* dzeta, double* dsigma, double* dthird, const double* x,
                                 const double* u, const double* zeta, const double* sigma, const double* third,
                                 const double* ddot, const double* laps, double f0, double beta, double g, double dt) {
  scalar_view_type dz = wrap1(dzeta, n), ds = wrap1(dsigma, n), d3 = wrap1(dthird, n);
  crd2 xv = wrap2(x, n);
  vec2 uv = wrap2(u, n);
  CoriolisBetaPlane cor(f0, beta);
  if (is_area) {
    Kokkos::parallel_for(n, SWEVorticityDivergenceAreaTendencies<PlaneGeometry>(
                                dz, ds, d3, xv, uv, wrap1(zeta, n), wrap1(sigma, n), wrap1(third, n), wrap1(ddot, n),
                                wrap1(laps, n), cor, g, dt));
  } else {
    Kokkos::parallel_for(n, SWEVorticityDivergenceHeightTendencies<PlaneGeometry>(
                                dz, ds, d3, xv, uv, wrap1(zeta, n), wrap1(sigma, n), wrap1(third, n), wrap1(ddot, n),
                                wrap1(laps, n), cor, g, dt));
  }
}

double oracle_plane_topography(int topo, const double* xy) {
  crd2 x = wrap2(xy, 1);
  const auto x0 = Kokkos::subview(x, 0, Kokkos::ALL);
  return topo == 1 ? PlanarGaussianMountain()(x0) : ZeroFunctor()(x0);
}

void oracle_swe_plane_set_surface_from_depth(int n, double* s, double* b, const double* x, const double* h, int topo) {
  scalar_view_type sv = wrap1(s, n), bv = wrap1(b, n);
  if (topo == 1)
    Kokkos::parallel_for(n, SetSurfaceFromDepth<PlaneGeometry, PlanarGaussianMountain>(sv, bv, wrap2(x, n), wrap1(h, n),
                                                                                       PlanarGaussianMountain()));
  else
    Kokkos::parallel_for(n, SetSurfaceFromDepth<PlaneGeometry, ZeroFunctor>(sv, bv, wrap2(x, n), wrap1(h, n),
                                                                            ZeroFunctor()));
}

void oracle_swe_plane_set_depth_surface_from_mass_area(int n, double* h, double* s, double* b, const double* x,
                                                       const double* m, const double* area, const uint8_t* mask,
                                                       int topo) {
  Mask fm(mask, n);
  scalar_view_type hv = wrap1(h, n), sv = wrap1(s, n), bv = wrap1(b, n);
  if (topo == 1)
    Kokkos::parallel_for(n, SetDepthAndSurfaceFromMassAndArea<PlaneGeometry, PlanarGaussianMountain>(
                                hv, sv, bv, wrap2(x, n), wrap1(m, n), wrap1(area, n), fm.v, PlanarGaussianMountain()));
  else
    Kokkos::parallel_for(n, SetDepthAndSurfaceFromMassAndArea<PlaneGeometry, ZeroFunctor>(
                                hv, sv, bv, wrap2(x, n), wrap1(m, n), wrap1(area, n), fm.v, ZeroFunctor()));
}

// ComputeFTLE (mesh/lpm_ftle.hpp) as launched by examples/sphere_rh54.cpp:308-316 / plane_colliding_dipoles.cpp:269-277
void oracle_ftle(int geom, int n_verts, const double* vert_phys, const double* vert_ref, int n_faces, double* face_phys,
                 const double* face_ref, const int* face_verts, const uint8_t* mask, double* ftle) {
  Mask fm(mask, n_faces);
  Kokkos::View<Index* [4]> fv(const_cast<int*>(face_verts), n_faces);
  if (geom == 0) {
    Kokkos::parallel_for(n_faces, ComputeFTLE<CubedSphereSeed>(wrap1(ftle, n_faces), wrap3(vert_phys, n_verts),
                                                               wrap3(vert_ref, n_verts), wrap3(face_phys, n_faces),
                                                               wrap3(face_ref, n_faces), fv, fm.v, 0.0));
  } else {
    using crd2d = PlaneGeometry::crd_view_type;
    auto w2 = [](const double* p, int n) { return crd2d(const_cast<double*>(p), n); };
    Kokkos::parallel_for(n_faces, ComputeFTLE<QuadRectSeed>(wrap1(ftle, n_faces), w2(vert_phys, n_verts),
                                                            w2(vert_ref, n_verts), w2(face_phys, n_faces),
                                                            w2(face_ref, n_faces), fv, fm.v, 0.0));
  }
}

double oracle_max_ftle(int n_faces, const double* ftle, const uint8_t* mask) {
  Mask fm(mask, n_faces);
  return get_max_ftle(wrap1(ftle, n_faces), fm.v, n_faces);
}

// write_vector_matlab / write_array_matlab (util/lpm_matlab_io.hpp) into a caller buffer; returns the length
static int copy_out(const std::string& s, char* buf, int cap) {
  if ((int)s.size() + 1 <= cap) std::memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}
int oracle_write_vector_matlab(const char* name, int n, const double* v, char* buf, int cap) {
  std::ostringstream os;
  write_vector_matlab(os, name, wrap1(v, n));
  return copy_out(os.str(), buf, cap);
}
int oracle_write_array_matlab(const char* name, int nrow, int ncol, const double* a, char* buf, int cap) {
  std::ostringstream os;
  Kokkos::View<Real**> av(const_cast<double*>(a), nrow, ncol);
  write_array_matlab(os, name, av);
  return copy_out(os.str(), buf, cap);
}

}  // extern "C"
