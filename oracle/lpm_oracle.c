/* lpm_oracle.c -- CPU restatement of pbosler/lpm's spherical direct-sum hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (lpm_b200/, include/) may call, link or load
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do, and there only as the checker or the timed CPU baseline.
 *
 * Every function restates, AS CODED (quirks included), a piece of the reference under
 * /root/reference/src; the citation is on each function.  Loop nest = what Kokkos-OpenMP executes
 * for TeamPolicy(n, AUTO) + TeamThreadRange/TeamVectorRange with team size 1: one OpenMP thread
 * per target, sequential j = 0..n_src-1, per-pair divide and log.  Views are LayoutRight
 * (x[i*3+k]) as on the reference's host execution space.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against the known-answer vectors
 * of the reference's tests/lpm_swe_kernels_tests.cpp:20-32,47-50,90-99,122-127, against analytic
 * solutions (solid-body rotation, examples/bve_rotation.cpp:147-167) and -- in the build
 * container -- against the reference's own headers compiled in place (oracle/_ref, see
 * oracle/Makefile and oracle/ref_driver.cpp).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp; GCC's default -ffp-contract=fast applies exactly
 * as it would for the reference's RelWithDebInfo build).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_PI 3.1415926535897932384626433832795027975 /* lpm_constants.hpp:11 */
#define ORACLE_ZERO_TOL 2.220446049250313e-16             /* lpm_floating_point.hpp:22 */

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

/* SphereGeometry::dot / cross (lpm_geometry.hpp:369-372, :381-386) */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* c, const double* a, const double* b) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* ---------------------------------------------------------------------------------------------
 * Family A: BVE on the sphere
 * ------------------------------------------------------------------------------------------- */

/* biot_savart (lpm_sphere_functions.hpp:44-57) */
static inline void biot_savart(double* u, const double* tgt, const double* src, double vort, double area) {
  cross3(u, tgt, src);
  const double strength = -vort * area / (4 * ORACLE_PI * (1 - dot3(src, tgt)));
  for (int k = 0; k < 3; ++k) u[k] *= strength;
}

/* greens_fn (lpm_sphere_functions.hpp:21-29) */
static inline double greens_fn(const double* tgt, const double* src, double vort, double area) {
  const double circ = -vort * area;
  return log(1 - dot3(tgt, src)) * circ / (4 * ORACLE_PI);
}

/* BVEVertexVelocity (collocated=0; lpm_bve_sphere_kernels.hpp:179-211 with VelocityReduceDistinct
 * :53-82) and BVEFaceVelocity (collocated=1; :365-394 with VelocityReduceCollocated :249-274).
 * In the collocated case the targets are the sources (tx is ignored). */
void oracle_bve_velocity(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                         const double* area, const uint8_t* mask, int collocated, double* vel) {
  if (collocated) tx = sx;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_tgt; ++i) {
    double acc[3] = {0, 0, 0};
    for (int j = 0; j < n_src; ++j) {
      double u[3] = {0, 0, 0};
      if (!mask[j] && !(collocated && i == j)) biot_savart(u, tx + 3 * i, sx + 3 * j, zeta[j], area[j]);
      for (int k = 0; k < 3; ++k) acc[k] += u[k];
    }
    for (int k = 0; k < 3; ++k) vel[3 * i + k] = acc[k];
  }
}

/* BVEVertexStreamFn (:141-170, StreamReduceDistinct :20-48), BVEFaceStreamFn (:329-356,
 * StreamReduceCollocated :218-242) */
void oracle_bve_streamfn(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                         const double* area, const uint8_t* mask, int collocated, double* psi) {
  if (collocated) tx = sx;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_tgt; ++i) {
    double acc = 0;
    for (int j = 0; j < n_src; ++j) {
      double p = 0;
      if (!mask[j] && !(collocated && i == j)) p = greens_fn(tx + 3 * i, sx + 3 * j, zeta[j], area[j]);
      acc += p;
    }
    psi[i] = acc;
  }
}

/* Higher-precision adjudicator for round-off questions: same sums with long double (x87 64-bit
 * mantissa) pair arithmetic and accumulation.  Not a restatement of anything in the reference. */
void oracle_bve_velocity_ld(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                            const double* area, const uint8_t* mask, int collocated, double* vel) {
  if (collocated) tx = sx;
  const long double four_pi = 4 * 3.14159265358979323846264338327950288L;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_tgt; ++i) {
    long double acc[3] = {0, 0, 0};
    const long double x0 = tx[3 * i], x1 = tx[3 * i + 1], x2 = tx[3 * i + 2];
    for (int j = 0; j < n_src; ++j) {
      if (mask[j] || (collocated && i == j)) continue;
      const long double y0 = sx[3 * j], y1 = sx[3 * j + 1], y2 = sx[3 * j + 2];
      const long double s = -(long double)zeta[j] * area[j] / (four_pi * (1 - (x0 * y0 + x1 * y1 + x2 * y2)));
      acc[0] += (x1 * y2 - x2 * y1) * s;
      acc[1] += (x2 * y0 - x0 * y2) * s;
      acc[2] += (x0 * y1 - x1 * y0) * s;
    }
    for (int k = 0; k < 3; ++k) vel[3 * i + k] = (double)acc[k];
  }
}

/* BVEFaceVelocity (collocated) for a SUBSET of the targets: row k of vel is the velocity at particle idx[k], summed over all
 * unmasked sources j != idx[k].  Lets the synthetic N = 1e6 configuration meet the reference arithmetic on a sample. */
void oracle_bve_velocity_subset(int n_idx, const int* idx, int n_src, const double* sx, const double* zeta, const double* area,
                                const uint8_t* mask, double* vel) {
#pragma omp parallel for schedule(static)
  for (int k = 0; k < n_idx; ++k) {
    const int i = idx[k];
    double acc[3] = {0, 0, 0};
    for (int j = 0; j < n_src; ++j) {
      double u[3] = {0, 0, 0};
      if (!mask[j] && i != j) biot_savart(u, sx + 3 * i, sx + 3 * j, zeta[j], area[j]);
      for (int c = 0; c < 3; ++c) acc[c] += u[c];
    }
    for (int c = 0; c < 3; ++c) vel[3 * k + c] = acc[c];
  }
}
void oracle_bve_velocity_subset_ld(int n_idx, const int* idx, int n_src, const double* sx, const double* zeta,
                                   const double* area, const uint8_t* mask, double* vel) {
  const long double four_pi = 4 * 3.14159265358979323846264338327950288L;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < n_idx; ++k) {
    const int i = idx[k];
    long double acc[3] = {0, 0, 0};
    const long double x0 = sx[3 * i], x1 = sx[3 * i + 1], x2 = sx[3 * i + 2];
    for (int j = 0; j < n_src; ++j) {
      if (mask[j] || i == j) continue;
      const long double y0 = sx[3 * j], y1 = sx[3 * j + 1], y2 = sx[3 * j + 2];
      const long double s = -(long double)zeta[j] * area[j] / (four_pi * (1 - (x0 * y0 + x1 * y1 + x2 * y2)));
      acc[0] += (x1 * y2 - x2 * y1) * s;
      acc[1] += (x2 * y0 - x0 * y2) * s;
      acc[2] += (x0 * y1 - x1 * y0) * s;
    }
    for (int c = 0; c < 3; ++c) vel[3 * k + c] = (double)acc[c];
  }
}

/* KokkosBlas::scal(r, a, x): r = a*x ; KokkosBlas::update(alpha, x, beta, y, gamma, z):
 * z = gamma*z + alpha*x + beta*y  (KokkosKernels 4.7 semantics; call sites cited below). */
static void blas_scal(long n, double* r, double a, const double* x) {
  for (long i = 0; i < n; ++i) r[i] = a * x[i];
}
static void blas_update(long n, double alpha, const double* x, double beta, const double* y, double gamma, double* z) {
  if (gamma == 0.0) {
    for (long i = 0; i < n; ++i) z[i] = alpha * x[i] + beta * y[i];
  } else {
    for (long i = 0; i < n; ++i) z[i] = gamma * z[i] + alpha * x[i] + beta * y[i];
  }
}

/* BVEVorticityTendency (lpm_bve_sphere_kernels.hpp:399-413) */
static void bve_vort_tendency(int n, double* dzeta, const double* vel, double dt, double Omega) {
  for (int i = 0; i < n; ++i) dzeta[i] = -2.0 * Omega * vel[3 * i + 2] * dt;
}

/* BVERK4Update (lpm_bve_rk4_impl.hpp:12-53) */
static void bve_rk4_update(int n, double* x, const double* x1, const double* x2, const double* x3, const double* x4,
                           double* vort, const double* v1, const double* v2, const double* v3, const double* v4) {
  const double sixth = 1.0 / 6.0, third = 1.0 / 3.0;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < 3; ++j)
      x[3 * i + j] += sixth * (x1[3 * i + j] + x4[3 * i + j]) + third * (x2[3 * i + j] + x3[3 * i + j]);
    vort[i] += sixth * (v1[i] + v4[i]) + third * (v2[i] + v3[i]);
  }
}

/* BVERK4::advance_timestep (lpm_bve_rk4_impl.hpp:63-167), n_steps times, in place.
 * Work views are zero-initialised as Kokkos Views are (lpm_bve_rk4.cpp:5-39).
 * Quirk A-i: the face update passes facevort4 in the facevort3 slot (:155-157). */
void oracle_bve_rk4_step(double dt, double Omega, int nv, double* vx, double* vz, double* vu, int nf, double* fx,
                         double* fz, double* fu, const double* fa, const uint8_t* fm, int n_steps) {
  double* w = (double*)calloc((size_t)20 * (nv + nf), sizeof(double));
  double *vx1 = w, *vx2 = vx1 + 3L * nv, *vx3 = vx2 + 3L * nv, *vx4 = vx3 + 3L * nv, *vxw = vx4 + 3L * nv;
  double *vz1 = vxw + 3L * nv, *vz2 = vz1 + nv, *vz3 = vz2 + nv, *vz4 = vz3 + nv, *vzw = vz4 + nv;
  double *fx1 = vzw + nv, *fx2 = fx1 + 3L * nf, *fx3 = fx2 + 3L * nf, *fx4 = fx3 + 3L * nf, *fxw = fx4 + 3L * nf;
  double *fz1 = fxw + 3L * nf, *fz2 = fz1 + nf, *fz3 = fz2 + nf, *fz4 = fz3 + nf, *fzw = fz4 + nf;
  for (int s = 0; s < n_steps; ++s) {
    /* stage 1 (:85-91) */
    blas_scal(3L * nv, vx1, dt, vu);
    bve_vort_tendency(nv, vz1, vu, dt, Omega);
    blas_scal(3L * nf, fx1, dt, fu);
    bve_vort_tendency(nf, fz1, fu, dt, Omega);
    /* stage 2 (:94-111) */
    blas_update(3L * nv, 1.0, vx, 0.5, vx1, 0.0, vxw);
    blas_update(nv, 1.0, vz, 0.5, vz1, 0.0, vzw);
    blas_update(3L * nf, 1.0, fx, 0.5, fx1, 0.0, fxw);
    blas_update(nf, 1.0, fz, 0.5, fz1, 0.0, fzw);
    oracle_bve_velocity(nv, vxw, nf, fxw, fzw, fa, fm, 0, vu);
    oracle_bve_velocity(nf, NULL, nf, fxw, fzw, fa, fm, 1, fu);
    blas_scal(3L * nv, vx2, dt, vu);
    blas_scal(3L * nf, fx2, dt, fu);
    bve_vort_tendency(nv, vz2, vu, dt, Omega);
    bve_vort_tendency(nf, fz2, fu, dt, Omega);
    /* stage 3 (:114-130) */
    blas_update(3L * nv, 1.0, vx, 0.5, vx2, 0.0, vxw);
    blas_update(nv, 1.0, vz, 0.5, vz2, 0.0, vzw);
    blas_update(3L * nf, 1.0, fx, 0.5, fx2, 0.0, fxw);
    blas_update(nf, 1.0, fz, 0.5, fz2, 0.0, fzw);
    oracle_bve_velocity(nv, vxw, nf, fxw, fzw, fa, fm, 0, vu);
    oracle_bve_velocity(nf, NULL, nf, fxw, fzw, fa, fm, 1, fu);
    blas_scal(3L * nv, vx3, dt, vu);
    blas_scal(3L * nf, fx3, dt, fu);
    bve_vort_tendency(nv, vz3, vu, dt, Omega);
    bve_vort_tendency(nf, fz3, fu, dt, Omega);
    /* stage 4 (:133-150) */
    blas_update(3L * nv, 1.0, vx, 1.0, vx3, 0.0, vxw);
    blas_update(nv, 1.0, vz, 1.0, vz3, 0.0, vzw);
    blas_update(3L * nf, 1.0, fx, 1.0, fx3, 0.0, fxw);
    blas_update(nf, 1.0, fz, 1.0, fz3, 0.0, fzw);
    oracle_bve_velocity(nv, vxw, nf, fxw, fzw, fa, fm, 0, vu);
    oracle_bve_velocity(nf, NULL, nf, fxw, fzw, fa, fm, 1, fu);
    blas_scal(3L * nv, vx4, dt, vu);
    blas_scal(3L * nf, fx4, dt, fu);
    bve_vort_tendency(nv, vz4, vu, dt, Omega);
    bve_vort_tendency(nf, fz4, fu, dt, Omega);
    /* update (:152-157); faces: vort3 slot receives facevort4 */
    bve_rk4_update(nv, vx, vx1, vx2, vx3, vx4, vz, vz1, vz2, vz3, vz4);
    bve_rk4_update(nf, fx, fx1, fx2, fx3, fx4, fz, fz1, fz2, fz4, fz4);
    /* velocity at the new state (:159-164) */
    oracle_bve_velocity(nv, vx, nf, fx, fz, fa, fm, 0, vu);
    oracle_bve_velocity(nf, NULL, nf, fx, fz, fa, fm, 1, fu);
  }
  free(w);
}

/* ---------------------------------------------------------------------------------------------
 * Family B: Incompressible2D on the sphere
 * ------------------------------------------------------------------------------------------- */

/* Incompressible2DKernels<SphereGeometry>::kernel_vals (lpm_incompressible2d_kernels.hpp:35-47) */
static inline void ic2d_kernel_vals(double* r, const double* x, const double* y, double eps) {
  const double arg = 1 - dot3(x, y) + eps * eps;
  cross3(r, x, y);
  for (int k = 0; k < 3; ++k) r[k] /= (-4 * ORACLE_PI * arg);
  r[3] = -log(arg) / (4 * ORACLE_PI);
}

/* Incompressible2DPassiveSums (targets_are_sources=0, :144-193) / ActiveSums (=1, :201-246) with
 * Incompressible2DReducer (:92-136).  collocated (skip j==i) = targets_are_sources && |eps| <
 * zero_tol (:235). */
void oracle_ic2d_sums(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                      const double* area, const uint8_t* mask, double eps, int targets_are_sources, double* vel,
                      double* psi) {
  if (targets_are_sources) tx = sx;
  const int collocated = targets_are_sources && (fabs(eps) < ORACLE_ZERO_TOL);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_tgt; ++i) {
    double s[4] = {0, 0, 0, 0};
    for (int j = 0; j < n_src; ++j) {
      if (!collocated || i != j) {
        if (!mask[j]) {
          const double gamma = zeta[j] * area[j];
          double r[4];
          ic2d_kernel_vals(r, tx + 3 * i, sx + 3 * j, eps);
          for (int k = 0; k < 4; ++k) s[k] += r[k] * gamma;
        }
      }
    }
    for (int k = 0; k < 3; ++k) vel[3 * i + k] = s[k];
    if (psi) psi[i] = s[3];
  }
}

/* Incompressible2DTendencies (:257-278) with CoriolisSphere::dfdt (lpm_coriolis.hpp:181-184):
 * dzeta = -2*Omega*u_z  (no dt) */
static void ic2d_tendency(int n, double* dzeta, const double* vel, double Omega) {
  for (int i = 0; i < n; ++i) dzeta[i] = -(2 * Omega * vel[3 * i + 2]);
}

/* Incompressible2DRK2::advance_timestep_impl (lpm_incompressible2d_rk2_impl.hpp:75-172), n_steps
 * times, in place.  passive = vertices, active = faces. */
void oracle_ic2d_rk2_step(double dt, double Omega, double eps, int np, double* px, double* pz, double* pu,
                          double* ppsi, int na, double* ax, double* az, double* au, double* apsi, const double* aa,
                          const uint8_t* am, int n_steps) {
  double* w = (double*)calloc((size_t)12 * (np + na), sizeof(double));
  double *px1 = w, *px2 = px1 + 3L * np, *pxw = px2 + 3L * np;
  double *pz1 = pxw + 3L * np, *pz2 = pz1 + np, *pzw = pz2 + np;
  double *ax1 = pzw + np, *ax2 = ax1 + 3L * na, *axw = ax2 + 3L * na;
  double *az1 = axw + 3L * na, *az2 = az1 + na, *azw = az2 + na;
  for (int s = 0; s < n_steps; ++s) {
    blas_scal(3L * np, px1, dt, pu); /* :77-82 */
    blas_scal(3L * na, ax1, dt, au);
    ic2d_tendency(np, pz1, pu, Omega); /* :85-92 */
    ic2d_tendency(na, az1, au, Omega);
    blas_update(np, 1, pz, dt, pz1, 0, pzw); /* :95-98 */
    blas_update(na, 1, az, dt, az1, 0, azw);
    blas_update(3L * np, 1, px, dt, pu, 0, pxw); /* :100-110 */
    blas_update(3L * na, 1, ax, dt, au, 0, axw);
    oracle_ic2d_sums(np, pxw, na, axw, azw, aa, am, eps, 0, pu, ppsi); /* :113-124 */
    oracle_ic2d_sums(na, NULL, na, axw, azw, aa, am, eps, 1, au, apsi);
    blas_scal(3L * np, px2, dt, pu); /* :127-132 */
    blas_scal(3L * na, ax2, dt, au);
    ic2d_tendency(np, pz2, pu, Omega); /* :135-142 */
    ic2d_tendency(na, az2, au, Omega);
    blas_update(np, 0.5 * dt, pz1, 0.5 * dt, pz2, 1, pz); /* :145-155 */
    blas_update(na, 0.5 * dt, az1, 0.5 * dt, az2, 1, az);
    blas_update(3L * np, 0.5, px1, 0.5, px2, 1, px);
    blas_update(3L * na, 0.5, ax1, 0.5, ax2, 1, ax);
    oracle_ic2d_sums(np, px, na, ax, az, aa, am, eps, 0, pu, ppsi); /* :157-170 */
    oracle_ic2d_sums(na, NULL, na, ax, az, aa, am, eps, 1, au, apsi);
  }
  free(w);
}

/* ---------------------------------------------------------------------------------------------
 * Family C: spherical shallow-water direct sums
 * ------------------------------------------------------------------------------------------- */

/* kzeta_sphere (lpm_swe_kernels.hpp:62-75): u += (x cross y) * (-vort*area / (4pi(1 - x.y + eps^2))) */
void oracle_kzeta_sphere(double* u, const double* x, const double* y, double vort, double area, double eps) {
  const double denom = 4 * ORACLE_PI * (1 - dot3(x, y) + eps * eps);
  const double strength = -vort * area / denom;
  double c[3];
  cross3(c, x, y);
  for (int k = 0; k < 3; ++k) u[k] += c[k] * strength;
}

/* ksigma_sphere (:86-104): u += (P_x y) * (-div*area / denom), P_x rows from proj_row
 * (lpm_geometry.hpp:486-493) */
void oracle_ksigma_sphere(double* u, const double* x, const double* y, double div, double area, double eps) {
  const double denom = 4 * ORACLE_PI * (1 - dot3(x, y) + eps * eps);
  const double strength = -div * area / denom;
  for (int j = 0; j < 3; ++j) {
    double row[3], uloc = 0;
    for (int k = 0; k < 3; ++k) row[k] = -x[j] * x[k];
    row[j] += 1;
    for (int k = 0; k < 3; ++k) uloc += row[k] * y[k];
    u[j] += uloc * strength;
  }
}

/* grad_kzeta (:113-148) and grad_ksigma (:178-307).
 * The reference spells both out as fully expanded polynomials in the components of x and y (about
 * 135 and 890 flops).  They are restated here in closed form -- the identity was checked against
 * the reference's own functions compiled in place (oracle/_ref, tests/test_ref_build.py) on and
 * off the unit sphere, to ~1e-14 relative:
 *   with kappa = 1 + eps^2, d = kappa - x.y, c = x cross y, q = kappa*x - y, p = y - (x.y) x,
 *   P = I - x x^T, [y]x the cross-product matrix ([y]x v = y cross v):
 *     grad_kzeta (x,y,eps)[3a+b] = ( d*[y]x[a][b] + c[a]*q[b] ) / (4 pi d^2)
 *     grad_ksigma(x,y,eps)[3a+b] = -( d*(x.y)*P[a][b] + q[a]*p[b] ) / (4 pi d^2)
 * NOTE the identities above hold for the polynomials as coded, which were derived under |x| = 1;
 * see oracle_swe_pair_coded_form for the literal (x-dependent) statement used off the sphere. */
void oracle_grad_kzeta(double* g, const double* x, const double* y, double eps) {
  const double kappa = 1 + eps * eps;
  const double xy = dot3(x, y);
  const double d = kappa - xy;
  const double denom = 1.0 / (4 * ORACLE_PI * d * d);
  double c[3];
  cross3(c, x, y);
  const double q[3] = {kappa * x[0] - y[0], kappa * x[1] - y[1], kappa * x[2] - y[2]};
  /* [y]x = [[0,-y2,y1],[y2,0,-y0],[-y1,y0,0]] */
  const double yx[9] = {0, -y[2], y[1], y[2], 0, -y[0], -y[1], y[0], 0};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) g[3 * a + b] = (d * yx[3 * a + b] + c[a] * q[b]) * denom;
}

void oracle_grad_ksigma(double* g, const double* x, const double* y, double eps) {
  const double kappa = 1 + eps * eps;
  const double xy = dot3(x, y);
  const double d = kappa - xy;
  const double denom = 1.0 / (4 * ORACLE_PI * d * d);
  const double q[3] = {kappa * x[0] - y[0], kappa * x[1] - y[1], kappa * x[2] - y[2]};
  const double p[3] = {y[0] - xy * x[0], y[1] - xy * x[1], y[2] - xy * x[2]};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      const double P = (a == b ? 1.0 : 0.0) - x[a] * x[b];
      g[3 * a + b] = -(d * xy * P + q[a] * p[b]) * denom;
    }
}

/* sphere_swe_velocity_sums (:334-362).  velz/vels are uninitialised in the reference (quirk C-i);
 * defined as 0 here, which is the evident intent and what the reference's unit test does
 * (tests/lpm_swe_kernels_tests.cpp:51-56). */
void oracle_swe_velocity_sums(double* r12, const double* x, const double* y, double zeta, double sigma,
                              double area, double eps) {
  double velz[3] = {0, 0, 0}, vels[3] = {0, 0, 0}, gkz[9], gks[9];
  oracle_kzeta_sphere(velz, x, y, zeta, area, eps);
  oracle_ksigma_sphere(vels, x, y, sigma, area, eps);
  for (int k = 0; k < 3; ++k) r12[k] = velz[k] + vels[k];
  oracle_grad_kzeta(gkz, x, y, eps);
  oracle_grad_ksigma(gks, x, y, eps);
  const double rot_str = -zeta * area;
  const double pot_str = -sigma * area;
  for (int k = 0; k < 9; ++k) r12[3 + k] = gkz[k] * rot_str + gks[k] * pot_str;
}

/* SphereVertexSums (targets_are_sources=0, :723-780) / SphereFaceSums (=1, :877-930) with
 * SphereSweDirectSumReducer (:578-619).  grad9 (optional) receives the 9 accumulated gradient
 * sums per target. */
void oracle_swe_sphere_sums(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                            const double* sigma, const double* area, const uint8_t* mask, double eps,
                            int targets_are_sources, int do_velocity, double* vel, double* ddot, double* grad9) {
  if (targets_are_sources) tx = sx;
  const int collocated = targets_are_sources && (fabs(eps) < ORACLE_ZERO_TOL);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_tgt; ++i) {
    double s[12];
    for (int k = 0; k < 12; ++k) s[k] = 0;
    for (int j = 0; j < n_src; ++j) {
      if (!mask[j]) {
        if (!collocated || i != j) {
          double r[12];
          oracle_swe_velocity_sums(r, tx + 3 * i, sx + 3 * j, zeta[j], sigma[j], area[j], eps);
          for (int k = 0; k < 12; ++k) s[k] += r[k];
        }
      }
    }
    if (do_velocity)
      for (int k = 0; k < 3; ++k) vel[3 * i + k] = s[k];
    double dd = 0;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) dd += s[3 + 3 * a + b] * s[3 + 3 * b + a];
    ddot[i] = dd;
    if (grad9)
      for (int k = 0; k < 9; ++k) grad9[9L * i + k] = s[3 + k];
  }
}

/* Higher-precision adjudicator of the family C sums (long double pair arithmetic and accumulation, the same formulas as
 * oracle_swe_velocity_sums): the 1/d^2 gradient terms are O(N) each and cancel to O(1), so two correct double-precision
 * summations differ by ~N * 2^-53 and ddot by more; this settles which side of a comparison carries the error.  Not a
 * restatement of anything in the reference.  Sequential j, as the reference. */
void oracle_swe_sphere_sums_ld(int n_tgt, const double* tx, int n_src, const double* sx, const double* zeta,
                               const double* sigma, const double* area, const uint8_t* mask, double eps,
                               int targets_are_sources, double* vel, double* ddot, double* grad9) {
  if (targets_are_sources) tx = sx;
  const int collocated = targets_are_sources && (fabs(eps) < ORACLE_ZERO_TOL);
  const long double four_pi = 4 * 3.14159265358979323846264338327950288L;
  const long double kappa = 1 + (long double)eps * eps;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_tgt; ++i) {
    long double s[12];
    for (int k = 0; k < 12; ++k) s[k] = 0;
    const long double x[3] = {tx[3 * i], tx[3 * i + 1], tx[3 * i + 2]};
    for (int j = 0; j < n_src; ++j) {
      if (mask[j] || (collocated && i == j)) continue;
      const long double y[3] = {sx[3 * j], sx[3 * j + 1], sx[3 * j + 2]};
      const long double xy = x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
      const long double d = kappa - xy;
      const long double rot = -(long double)zeta[j] * area[j], pot = -(long double)sigma[j] * area[j];
      const long double c[3] = {x[1] * y[2] - x[2] * y[1], x[2] * y[0] - x[0] * y[2], x[0] * y[1] - x[1] * y[0]};
      const long double q[3] = {kappa * x[0] - y[0], kappa * x[1] - y[1], kappa * x[2] - y[2]};
      const long double p[3] = {y[0] - xy * x[0], y[1] - xy * x[1], y[2] - xy * x[2]};
      const long double yx[9] = {0, -y[2], y[1], y[2], 0, -y[0], -y[1], y[0], 0};
      const long double inv1 = 1 / (four_pi * d), inv2 = 1 / (four_pi * d * d);
      for (int a = 0; a < 3; ++a) {
        s[a] += c[a] * rot * inv1 + p[a] * pot * inv1; /* kzeta + ksigma (P_x y = y - (x.y) x) */
        for (int b = 0; b < 3; ++b) {
          const long double P = (a == b ? 1.0L : 0.0L) - x[a] * x[b];
          s[3 + 3 * a + b] += (d * yx[3 * a + b] + c[a] * q[b]) * inv2 * rot - (d * xy * P + q[a] * p[b]) * inv2 * pot;
        }
      }
    }
    for (int k = 0; k < 3; ++k) vel[3 * i + k] = (double)s[k];
    long double dd = 0;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) dd += s[3 + 3 * a + b] * s[3 + 3 * b + a];
    ddot[i] = (double)dd;
    if (grad9)
      for (int k = 0; k < 9; ++k) grad9[9L * i + k] = (double)s[3 + k];
  }
}

/* ---------------------------------------------------------------------------------------------
 * Family C stepper: SWERK2 on the sphere (direct sums + O(N) algebra; the GMLS surface Laplacian
 * is an external input, supplied through `laps_fn`)
 * ------------------------------------------------------------------------------------------- */

/* CoriolisSphere (lpm_coriolis.hpp:154-195) */
static inline double cor_f(double Omega, const double* x) { return 2 * Omega * x[2]; }
static inline double cor_dfdt(double Omega, const double* u) { return 2 * Omega * u[2]; }
static inline double cor_grad_f_cross_u(double Omega, const double* x, const double* u) {
  return -2 * Omega * (-u[0] * x[1] + u[1] * x[0]);
}

/* SWEVorticityDivergenceHeightTendencies<SphereGeometry> (lpm_swe_kernels.hpp:941-999) when
 * is_area == 0 (third output dh = -sigma*h*dt, third input h = depth) and
 * SWEVorticityDivergenceAreaTendencies<SphereGeometry> (:1010-1069) when is_area != 0 (third
 * output darea = sigma*area*dt, third input = area). */
void oracle_swe_tendencies(int n, int is_area, double* dzeta, double* dsigma, double* dthird, const double* x,
                           const double* u, const double* zeta, const double* sigma, const double* third,
                           const double* ddot, const double* laps, double Omega, double g, double dt) {
  for (int i = 0; i < n; ++i) {
    const double* xi = x + 3 * i;
    const double* ui = u + 3 * i;
    const double f = cor_f(Omega, xi);
    dzeta[i] = (-cor_dfdt(Omega, ui) - (zeta[i] + f) * sigma[i]) * dt;
    dsigma[i] = (f * zeta[i] + cor_grad_f_cross_u(Omega, xi, ui) - ddot[i] - g * laps[i] - dot3(ui, ui)) * dt;
    dthird[i] = is_area ? (sigma[i] * third[i]) * dt : (-sigma[i] * third[i]) * dt;
  }
}

/* SetSurfaceFromDepth<SphereGeometry, ZeroFunctor> (:1079-1100): b = topo(x) = 0, s = h + b.
 * (The constructor never copies `topo`, quirk C-v: the functor is default-constructed.) */
void oracle_swe_set_surface_from_depth(int n, double* s, double* b, const double* h) {
  for (int i = 0; i < n; ++i) {
    b[i] = 0.0;
    s[i] = h[i] + b[i];
  }
}

/* SetDepthAndSurfaceFromMassAndArea<SphereGeometry, ZeroFunctor> (:1110-1140): unmasked faces only */
void oracle_swe_set_depth_surface_from_mass_area(int n, double* h, double* s, double* b, const double* m,
                                                 const double* area, const uint8_t* mask) {
  for (int i = 0; i < n; ++i) {
    if (!mask[i]) {
      h[i] = m[i] / area[i];
      b[i] = 0.0;
      s[i] = b[i] + h[i];
    }
  }
}

/* The surface-Laplacian provider: fills plaps[np] and alaps[na] for the particle positions and surface
 * heights it is given.  stage 1 = predictor state (lpm_swe_rk2_impl.hpp:134-154), stage 2 = new state
 * (:233-252).  NULL = leave the arrays as they are. */
typedef void (*oracle_laps_fn)(void* user, int stage, int np, const double* px, const double* psurf, double* plaps,
                               int na, const double* ax, const double* asurf, const uint8_t* amask, double* alaps);

/* SWERK2<Seed, ZeroFunctor>::advance_timestep_impl (lpm_swe_rk2_impl.hpp:80-258) for SphereGeometry,
 * n_steps times, in place.  p* = passive (vertices), a* = active (faces).  On entry velocity, double dot and
 * surface Laplacian must belong to the current state (SWE::init_direct_sums + the SWERK2 constructor,
 * :56-77). */
void oracle_swe_rk2_step(double dt, double Omega, double g, double eps, int np, double* px, double* pz, double* ps,
                         double* ph, double* psurf, double* pbot, double* pu, double* pdd, double* plaps, int na,
                         double* ax, double* az, double* as, double* aarea, double* amass, double* ah, double* asurf,
                         double* abot, double* au, double* add, double* alaps, const uint8_t* am,
                         oracle_laps_fn laps_fn, void* user, int n_steps) {
  double* w = (double*)calloc((size_t)18 * (np + na), sizeof(double));
  double *px1 = w, *px2 = px1 + 3L * np, *pxw = px2 + 3L * np;
  double *pz1 = pxw + 3L * np, *pz2 = pz1 + np, *pzw = pz2 + np;
  double *ps1 = pzw + np, *ps2 = ps1 + np, *psw = ps2 + np;
  double *ph1 = psw + np, *ph2 = ph1 + np, *phw = ph2 + np;
  double *ax1 = phw + np, *ax2 = ax1 + 3L * na, *axw = ax2 + 3L * na;
  double *az1 = axw + 3L * na, *az2 = az1 + na, *azw = az2 + na;
  double *as1 = azw + na, *as2 = as1 + na, *asw = as2 + na;
  double *aa1 = asw + na, *aa2 = aa1 + na, *aaw = aa2 + na;
  for (int s = 0; s < n_steps; ++s) {
    /* stage 1 (:85-104) */
    blas_scal(3L * np, px1, dt, pu);
    blas_scal(3L * na, ax1, dt, au);
    oracle_swe_tendencies(np, 0, pz1, ps1, ph1, px, pu, pz, ps, ph, pdd, plaps, Omega, g, dt);
    oracle_swe_tendencies(na, 1, az1, as1, aa1, ax, au, az, as, aarea, add, alaps, Omega, g, dt);
    /* predictor state (:108-123) */
    blas_update(3L * np, 1, px, 1, px1, 0, pxw);
    blas_update(np, 1, pz, 1, pz1, 0, pzw);
    blas_update(np, 1, ps, 1, ps1, 0, psw);
    blas_update(3L * na, 1, ax, 1, ax1, 0, axw);
    blas_update(na, 1, az, 1, az1, 0, azw);
    blas_update(na, 1, as, 1, as1, 0, asw);
    blas_update(np, 1, ph, 1, ph1, 0, phw);
    blas_update(na, 1, aarea, 1, aa1, 0, aaw);
    oracle_swe_set_surface_from_depth(np, psurf, pbot, phw);                              /* :124-127 */
    oracle_swe_set_depth_surface_from_mass_area(na, ah, asurf, abot, amass, aaw, am);     /* :128-132 */
    if (laps_fn) laps_fn(user, 1, np, pxw, psurf, plaps, na, axw, asurf, am, alaps);      /* :134-154 */
    oracle_swe_sphere_sums(np, pxw, na, axw, azw, asw, aaw, am, eps, 0, 1, pu, pdd, NULL); /* :157-168 */
    oracle_swe_sphere_sums(na, NULL, na, axw, azw, asw, aaw, am, eps, 1, 1, au, add, NULL);
    /* stage 2 tendencies at the predictor state (:170-185) */
    oracle_swe_tendencies(np, 0, pz2, ps2, ph2, pxw, pu, pzw, psw, phw, pdd, plaps, Omega, g, dt);
    oracle_swe_tendencies(na, 1, az2, as2, aa2, axw, au, azw, asw, aaw, add, alaps, Omega, g, dt);
    blas_scal(3L * np, px2, dt, pu); /* :187-188 */
    blas_scal(3L * na, ax2, dt, au);
    /* Heun combine (:190-205) */
    blas_update(3L * np, 0.5, px1, 0.5, px2, 1, px);
    blas_update(np, 0.5, pz1, 0.5, pz2, 1, pz);
    blas_update(np, 0.5, ps1, 0.5, ps2, 1, ps);
    blas_update(np, 0.5, ph1, 0.5, ph2, 1, ph);
    blas_update(3L * na, 0.5, ax1, 0.5, ax2, 1, ax);
    blas_update(na, 0.5, az1, 0.5, az2, 1, az);
    blas_update(na, 0.5, as1, 0.5, as2, 1, as);
    blas_update(na, 0.5, aa1, 0.5, aa2, 1, aarea);
    oracle_swe_set_surface_from_depth(np, psurf, pbot, ph);                               /* :207-211 */
    oracle_swe_set_depth_surface_from_mass_area(na, ah, asurf, abot, amass, aarea, am);   /* :212-216 */
    oracle_swe_sphere_sums(np, px, na, ax, az, as, aarea, am, eps, 0, 1, pu, pdd, NULL);  /* :218-231 */
    oracle_swe_sphere_sums(na, NULL, na, ax, az, as, aarea, am, eps, 1, 1, au, add, NULL);
    if (laps_fn) laps_fn(user, 2, np, px, psurf, plaps, na, ax, asurf, am, alaps);        /* :233-252 */
  }
  free(w);
}

/* ---------------------------------------------------------------------------------------------
 * FTLE diagnostic: ComputeFTLE<SeedType>::operator() for quadrilateral faces,
 * sphere (mesh/lpm_ftle.hpp:86-220) and plane (:222-319), and get_max_ftle (:327-338).
 * As coded: the "Cauchy-Green tensor" is the ELEMENTWISE product F_ij * F_ji (:70-79), not F^T F; on the
 * sphere the face's physical coordinates are normalised IN PLACE (:98, a write to phys_crds_faces); the
 * result is log(lambda_1) with no division by 2t (:218).
 * ------------------------------------------------------------------------------------------- */

/* north_pole_rotation_matrix (util/lpm_math.hpp:199-217) */
static void north_pole_rotation_matrix(double* r, const double* x) {
  const double cosy = sqrt(x[1] * x[1] + x[2] * x[2]);
  const double siny = x[0];
  const int on_x_axis = fabs(cosy) < ORACLE_ZERO_TOL;
  const double cosx = on_x_axis ? 1 : x[2] / cosy;
  const double sinx = on_x_axis ? 0 : x[1] / cosy;
  r[0] = cosy, r[1] = -sinx * siny, r[2] = -cosx * siny;
  r[3] = 0, r[4] = cosx, r[5] = -sinx;
  r[6] = siny, r[7] = cosy * sinx, r[8] = cosx * cosy;
}

/* apply_3by3 (util/lpm_math.hpp:244-253) */
static void apply_3by3(double* out, const double* m, const double* x) {
  for (int i = 0; i < 3; ++i) {
    out[i] = 0;
    for (int j = 0; j < 3; ++j) out[i] += m[3 * i + j] * x[j];
  }
}

/* set_flow_map_gradient + cauchy_green_tensor + two_by_two_real_eigenvalues (mesh/lpm_ftle.hpp:54-79,
 * util/lpm_math.hpp:119-141); returns log(lambda_1) */
static double ftle_from_edges(const double* e0rev_phys, const double* e1_phys, const double* xdir, const double* ydir,
                              double dx0, double dy0) {
  double F[4], cg[4];
  F[0] = (e1_phys[0] * xdir[0] + e1_phys[1] * xdir[1]) / dx0;
  F[1] = (e0rev_phys[0] * xdir[0] + e0rev_phys[1] * xdir[1]) / dx0;
  F[2] = (e1_phys[0] * ydir[0] + e1_phys[1] * ydir[1]) / dy0;
  F[3] = (e0rev_phys[0] * ydir[0] + e0rev_phys[1] * ydir[1]) / dy0;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) cg[2 * i + j] = F[2 * i + j] * F[2 * j + i];
  const double det = cg[0] * cg[3] - cg[1] * cg[2];
  const double half_trace = 0.5 * (cg[0] + cg[3]);
  double sqrt_arg = half_trace * half_trace - det;
  if (fabs(sqrt_arg) < ORACLE_ZERO_TOL) sqrt_arg = 0;
  return log(half_trace + sqrt(sqrt_arg));
}

/* geom: 0 = sphere (ndim 3), 1 = plane (ndim 2).  face_verts is [n_faces][4] row-major.  ftle(i) is written for
 * leaves only (masked entries are left untouched, as in the reference). */
void oracle_ftle(int geom, int n_verts, const double* vert_phys, const double* vert_ref, int n_faces, double* face_phys,
                 const double* face_ref, const int* face_verts, const uint8_t* mask, double* ftle) {
  (void)n_verts;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < n_faces; ++f) {
    if (mask[f]) continue;
    if (geom == 0) {
      const double* fa = face_ref + 3 * f;
      double* fx = face_phys + 3 * f;
      const double s = 1.0 / sqrt(fx[0] * fx[0] + fx[1] * fx[1] + fx[2] * fx[2]); /* SphereGeometry::normalize */
      for (int k = 0; k < 3; ++k) fx[k] *= s;
      double rr[9], rp[9], vp[4][3], vr[4][3];
      north_pole_rotation_matrix(rr, fa);
      north_pole_rotation_matrix(rp, fx);
      for (int i = 0; i < 4; ++i) {
        const int v = face_verts[4 * f + i];
        apply_3by3(vr[i], rr, vert_ref + 3 * v);
        apply_3by3(vp[i], rp, vert_phys + 3 * v);
      }
      /* "shift so that vertex 1 is the origin" (:148-153) is an in-place loop over i = 0..3: i == 1 zeroes vertex 1,
       * so vertices 2 and 3 are then shifted by 0 (edge 1 below is vertex 2's tangent-plane position).  As coded. */
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j) {
          vp[i][j] -= vp[1][j];
          vr[i][j] -= vr[1][j];
        }
      const double e1r[2] = {vr[2][0] - vr[1][0], vr[2][1] - vr[1][1]};
      const double e1p[2] = {vp[2][0] - vp[1][0], vp[2][1] - vp[1][1]};
      const double dx0 = sqrt(e1r[0] * e1r[0] + e1r[1] * e1r[1]);
      double xdir[2], ydir[2];
      double xr = sqrt(e1r[0] * e1r[0] + e1r[1] * e1r[1]);
      xdir[0] = e1r[0] / xr, xdir[1] = e1r[1] / xr;
      const double e0r[2] = {vr[0][0], vr[0][1]}, e0p[2] = {vp[0][0], vp[0][1]};
      const double dxy = xdir[0] * e0r[0] + xdir[1] * e0r[1];
      ydir[0] = e0r[0] - dxy * xdir[0], ydir[1] = e0r[1] - dxy * xdir[1];
      const double dy0 = sqrt(ydir[0] * ydir[0] + ydir[1] * ydir[1]);
      const double yr = sqrt(ydir[0] * ydir[0] + ydir[1] * ydir[1]);
      ydir[0] /= yr, ydir[1] /= yr;
      ftle[f] = ftle_from_edges(e0p, e1p, xdir, ydir, dx0, dy0);
    } else {
      double vp[4][2], vr[4][2];
      for (int i = 0; i < 4; ++i) {
        const int v = face_verts[4 * f + i];
        for (int j = 0; j < 2; ++j) vp[i][j] = vert_phys[2 * v + j], vr[i][j] = vert_ref[2 * v + j];
      }
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j) {
          vp[i][j] -= vp[1][j];
          vr[i][j] -= vr[1][j];
        }
      const double e1r[2] = {vr[2][0] - vr[1][0], vr[2][1] - vr[1][1]};
      const double e1p[2] = {vp[2][0] - vp[1][0], vp[2][1] - vp[1][1]};
      const double dx0 = sqrt(e1r[0] * e1r[0] + e1r[1] * e1r[1]);
      const double xl = sqrt(e1r[0] * e1r[0] + e1r[1] * e1r[1]);
      const double xdir[2] = {e1r[0] / xl, e1r[1] / xl};
      const double e0p[2] = {vp[0][0], vp[0][1]}, e0r[2] = {vr[0][0], vr[0][1]};
      const double dy0 = sqrt(e0r[0] * e0r[0] + e0r[1] * e0r[1]);
      const double yl = sqrt(e0r[0] * e0r[0] + e0r[1] * e0r[1]);
      const double ydir[2] = {e0r[0] / yl, e0r[1] / yl};
      ftle[f] = ftle_from_edges(e0p, e1p, xdir, ydir, dx0, dy0);
    }
  }
}

/* get_max_ftle (mesh/lpm_ftle.hpp:327-338): Kokkos::Max identity = lowest double */
double oracle_max_ftle(int n_faces, const double* ftle, const uint8_t* mask) {
  double m = -1.7976931348623157e308;
  for (int i = 0; i < n_faces; ++i)
    if (!mask[i]) m = m > ftle[i] ? m : ftle[i];
  return m;
}
