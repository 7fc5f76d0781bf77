// ref_ic2d_driver.cpp -- C entry point around the REFERENCE's own Incompressible2D<Seed> + Incompressible2DRK2 (the model and stepper
// examples/sphere_rh54.cpp and sphere_gaussian_vortex.cpp run), compiled in place from /root/reference/src (never copied) against
// oracle/kokkos_shim and the declarations-only Compadre stand-in (oracle/kokkos_shim/compadre_stub: the remesh headers these classes
// include mention Compadre; the stepper path calls none of it).  Part of oracle/_ref/liblpm_ref_mesh.so.  TEST INFRASTRUCTURE.
//
// Reference code exercised (as shipped): lpm_incompressible2d{.hpp,_impl.hpp} (constructor, init_direct_sums),
// lpm_incompressible2d_rk2{.hpp,_impl.hpp} (advance_timestep_impl, :75-172), lpm_incompressible2d_kernels.hpp
// (Incompressible2DPassiveSums / ActiveSums / Tendencies), lpm_coriolis.hpp (CoriolisSphere), mesh/* (PolyMesh2d).
#include <memory>

#include "lpm_coriolis.hpp"
#include "lpm_incompressible2d.hpp"
#include "lpm_incompressible2d_impl.hpp"
#include "lpm_incompressible2d_rk2.hpp"
#include "lpm_incompressible2d_rk2_impl.hpp"

using namespace Lpm;

namespace {
template <class Seed>
int ic2d_rk2_run(int depth, double dt, double omega, double eps, int n_steps, const double* vert_zeta, const double* face_zeta,
                 double* vx, double* vz, double* vu, double* vpsi, double* fx, double* fz, double* fu, double* fpsi) {
  PolyMeshParameters<Seed> params(depth, 1.0, 0, 0);
  CoriolisSphere coriolis(omega);
  auto ic2d = std::make_unique<Incompressible2D<Seed>>(params, coriolis, eps);
  const int nv = ic2d->mesh.n_vertices_host(), nf = ic2d->mesh.n_faces_host();
  // the caller's relative vorticity (the arrays the engine under test gets); absolute vorticity as init_vorticity forms it
  for (int i = 0; i < nv; ++i) {
    ic2d->rel_vort_passive.view(i) = vert_zeta[i];
    ic2d->abs_vort_passive.view(i) = vert_zeta[i] + 2 * omega * ic2d->mesh.vertices.phys_crds.view(i, 2);
  }
  for (int i = 0; i < nf; ++i) {
    ic2d->rel_vort_active.view(i) = face_zeta[i];
    ic2d->abs_vort_active.view(i) = face_zeta[i] + 2 * omega * ic2d->mesh.faces.phys_crds.view(i, 2);
  }
  ic2d->init_direct_sums();
  if (n_steps > 0) {
    Incompressible2DRK2<Seed> solver(dt, *ic2d);
    for (int s = 0; s < n_steps; ++s) ic2d->advance_timestep(solver);
  }
  for (int i = 0; i < nv; ++i) {
    for (int k = 0; k < 3; ++k) {
      vx[3 * i + k] = ic2d->mesh.vertices.phys_crds.view(i, k);
      vu[3 * i + k] = ic2d->velocity_passive.view(i, k);
    }
    vz[i] = ic2d->rel_vort_passive.view(i);
    vpsi[i] = ic2d->stream_fn_passive.view(i);
  }
  for (int i = 0; i < nf; ++i) {
    for (int k = 0; k < 3; ++k) {
      fx[3 * i + k] = ic2d->mesh.faces.phys_crds.view(i, k);
      fu[3 * i + k] = ic2d->velocity_active.view(i, k);
    }
    fz[i] = ic2d->rel_vort_active.view(i);
    fpsi[i] = ic2d->stream_fn_active.view(i);
  }
  return 0;
}
}  // namespace

extern "C" {
// Incompressible2D<Seed>(depth, CoriolisSphere(omega), eps) with the given relative vorticity -> init_direct_sums ->
// n_steps x advance_timestep(Incompressible2DRK2).  seed: 0 icos, 1 cubed.  Outputs sized for the mesh.
int ref_ic2d_rk2_run(int seed, int depth, double dt, double omega, double eps, int n_steps, const double* vert_zeta,
                     const double* face_zeta, double* vx, double* vz, double* vu, double* vpsi, double* fx, double* fz, double* fu,
                     double* fpsi) {
  if (seed == 0)
    return ic2d_rk2_run<IcosTriSphereSeed>(depth, dt, omega, eps, n_steps, vert_zeta, face_zeta, vx, vz, vu, vpsi, fx, fz, fu, fpsi);
  if (seed == 1)
    return ic2d_rk2_run<CubedSphereSeed>(depth, dt, omega, eps, n_steps, vert_zeta, face_zeta, vx, vz, vu, vpsi, fx, fz, fu, fpsi);
  return -1;
}
}
