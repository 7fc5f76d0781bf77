// lpmx_gmls_core.h -- per-target arithmetic of the surface Laplacian on the sphere.
//
// Replaces, on the device, the host round trip SWERK2 makes every stage (src/lpm_swe_rk2_impl.hpp:56-77,134-154,
// 233-252): gmls::Neighborhoods (src/lpm_compadre.cpp:27-70: k-nearest search for min_neighbors points, window radius
// = eps_multiplier x distance to the k-th, then every point inside the window) and gmls::sphere_scalar_gmls +
// Compadre::Evaluator with LaplacianOfScalarPointEvaluation (src/lpm_compadre.hpp:163-195: ScalarTaylorPolynomial
// basis of samples_order, MANIFOLD problem with a manifold_order graph reconstruction, Power weights).
//
// Compadre 1.6.2 is a third-party dependency that is ABSENT from /root/reference (tools/README.md:12-19 pins the
// version; no source is vendored), so this is a restatement of its PUBLISHED algorithm (Trask, Kuberry et al.,
// "Compatible meshfree discretization of surface PDEs", and the Compadre toolkit documentation), not of its code:
//   1. chart: orthonormal tangent frame (T1, T2, n) at the target; neighbours y_j get s = (y_j - x).T1,
//      t = (y_j - x).T2, height = (y_j - x).n.  Compadre estimates n by PCA and orients it with the reference
//      normal it is given (the target coordinates, setReferenceOutwardNormalDirection); here n is that reference
//      normal itself, x/|x| -- on a sphere they coincide up to the reconstruction error.
//   2. weighted least squares in the scaled Taylor basis b_k(s/eps, t/eps) = (s/eps)^ax (t/eps)^ay / (ax! ay!),
//      n = 0..order, ay = 0..n, weights W(r) = (1 - r/eps)^p (WeightingFunctionType::Power, p = *_weight_pwr), once
//      for the height (manifold_order) and once for the data (samples_order).  Compadre factors sqrt(W) P with QR;
//      here the normal equations P^T W P are accumulated on the fly as weighted moments (no neighbour list is stored)
//      and solved by Cholesky -- the same minimiser; the scaled basis keeps cond(P^T W P) < 1e7 for order <= 4.
//   3. Laplace-Beltrami at the target in the graph chart, metric g_ij = delta_ij + h_i h_j:
//        lap f = g^ij ( f_ij - h_ij (grad h . grad f) / (1 + |grad h|^2) ).
// Parity is therefore UNPINNED against Compadre; the pins are analytic: spherical harmonics
// (lap Y_l = -l(l+1) Y_l), the TC2 closed form (examples/sphere_swe_tc2.cpp:243-244) and the numpy restatement
// oracle/gmls_oracle.py (independent least-squares path: sqrt-weighted lstsq on explicit neighbour lists).
//
// Everything here is __host__ __device__ so that tests can run the very same arithmetic on the CPU
// (oracle/gmls_core_host.cpp, test infrastructure); the product path is lpmx_gmls.cu only.
#ifndef LPMX_GMLS_CORE_H
#define LPMX_GMLS_CORE_H

#include <math.h>

#ifdef __CUDACC__
#define LPMX_HD __host__ __device__ __forceinline__
#define LPMX_UNROLL _Pragma("unroll")
#else
#define LPMX_HD inline
#define LPMX_UNROLL
#endif

namespace lpmx {
namespace gmls {

constexpr int kMaxOrder = 4;
constexpr int kMaxNP = 15;  // (order + 1)(order + 2) / 2 at order 4 = Compadre::GMLS::getNP(4, 2)
constexpr int kMaxK = 32;   // largest min_neighbors

LPMX_HD int np_of(int order) { return (order + 1) * (order + 2) / 2; }

// Sorted-by-cell point cloud on a uniform grid over [-box, box]^3
struct Cloud {
  const double* x;        // sorted coordinates, SoA: x[k * n + i]
  const double* f;        // sorted data values
  const int* cell_start;  // G^3 + 1 entries: cell q holds the sorted indices [cell_start[q], cell_start[q + 1])
  int n;
  int G;           // cells per dimension
  double box;      // half width of the grid
  double inv_cell; // 1 / cell size
  double cell;
};

struct Params {
  int samples_order, manifold_order, min_neighbors;
  double eps_multiplier, weight_pwr;
};

// Grid sizing shared by the device path and the host harness.  For n quasi-uniform points on a sphere of radius R the
// K-th neighbour sits at about 2 R sqrt(K / n); the cell is the expected window radius (multiplier x that), so the
// window search visits 27 cells.  G is capped to bound the cell table (G^3 + 1 ints); a coarser grid is only slower.
struct GridDims {
  int G;
  double cell, box;
};
inline GridDims grid_dims(int n, int min_neighbors, double eps_multiplier, double radius) {
  GridDims g;
  g.box = 1.05 * radius;
  double cell = eps_multiplier * 2.0 * radius * sqrt((double)min_neighbors / (double)(n > 0 ? n : 1));
  int G = (int)ceil(2.0 * g.box / cell);
  if (G > 320) G = 320;
  if (G < 1) G = 1;
  g.G = G;
  g.cell = 2.0 * g.box / G;
  return g;
}

LPMX_HD int cell_coord(const Cloud& c, double v) {
  int q = (int)floor((v + c.box) * c.inv_cell);
  return q < 0 ? 0 : (q >= c.G ? c.G - 1 : q);
}

// Scaled Taylor monomials, Compadre's 2-D ordering: n = 0..order, ay = 0..n, ax = n - ay
LPMX_HD void taylor_basis(int order, double u, double v, double* b) {
  double pu[kMaxOrder + 1], pv[kMaxOrder + 1];  // u^a / a!
  pu[0] = 1.0, pv[0] = 1.0;
  for (int a = 1; a <= order; ++a) pu[a] = pu[a - 1] * u / a, pv[a] = pv[a - 1] * v / a;
  int k = 0;
  for (int n = 0; n <= order; ++n)
    for (int ay = 0; ay <= n; ++ay) b[k++] = pu[n - ay] * pv[ay];
}

LPMX_HD double power_weight(double r_over_eps, double p) {
  const double a = 1.0 - r_over_eps;
  if (!(a > 0.0)) return 0.0;
  return p == 2.0 ? a * a : pow(a, p);
}

// packed lower-triangular index
LPMX_HD int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// In-place Cholesky of the leading np x np block of the packed matrix M and solution of M a = r for two right-hand
// sides.  Returns false when a pivot is not positive (too few / degenerate neighbours) or when the pivots span more than
// kMaxPivotRatio: the fit is then ill-conditioned (cond(P^T W P) ~ max pivot / min pivot; the scaled basis keeps it below 1e7 on
// the quasi-uniform meshes) and its Taylor coefficients carry no digits -- the callers return NaN instead of a number that
// looks like a Laplacian (strongly non-uniform clouds, e.g. all neighbours of a target along one line).
constexpr double kMaxPivotRatio = 1e12;
LPMX_HD bool cholesky_factor(int np, double* M) {
  double dmin = 0.0, dmax = 0.0;
  for (int j = 0; j < np; ++j) {
    double d = M[tri(j, j)];
    for (int k = 0; k < j; ++k) d -= M[tri(j, k)] * M[tri(j, k)];
    if (!(d > 0.0)) return false;
    dmin = (j == 0 || d < dmin) ? d : dmin;
    dmax = d > dmax ? d : dmax;
    if (!(dmin * kMaxPivotRatio > dmax)) return false;
    d = sqrt(d);
    M[tri(j, j)] = d;
    for (int i = j + 1; i < np; ++i) {
      double s = M[tri(i, j)];
      for (int k = 0; k < j; ++k) s -= M[tri(i, k)] * M[tri(j, k)];
      M[tri(i, j)] = s / d;
    }
  }
  return true;
}
LPMX_HD void cholesky_solve(int np, const double* L, double* r) {
  for (int i = 0; i < np; ++i) {
    double s = r[i];
    for (int k = 0; k < i; ++k) s -= L[tri(i, k)] * r[k];
    r[i] = s / L[tri(i, i)];
  }
  for (int i = np - 1; i >= 0; --i) {
    double s = r[i];
    for (int k = i + 1; k < np; ++k) s -= L[tri(k, i)] * r[k];
    r[i] = s / L[tri(i, i)];
  }
}

// Laplace-Beltrami at the chart origin from the Taylor coefficients of the data (af) and of the height (ah),
// both in the basis scaled by eps (order >= 2 for af; ah may be of order 1, whose second derivatives are 0)
LPMX_HD double laplace_beltrami(const double* af, int order_f, const double* ah, int order_h, double eps) {
  const double ie = 1.0 / eps, ie2 = ie * ie;
  const double fs = af[1] * ie, ft = af[2] * ie;
  const double fss = order_f >= 2 ? af[3] * ie2 : 0.0, fst = order_f >= 2 ? af[4] * ie2 : 0.0,
               ftt = order_f >= 2 ? af[5] * ie2 : 0.0;
  const double hs = ah[1] * ie, ht = ah[2] * ie;
  const double hss = order_h >= 2 ? ah[3] * ie2 : 0.0, hst = order_h >= 2 ? ah[4] * ie2 : 0.0,
               htt = order_h >= 2 ? ah[5] * ie2 : 0.0;
  const double g = 1.0 + hs * hs + ht * ht;
  const double q = (hs * fs + ht * ft) / g;
  const double g11 = (1.0 + ht * ht) / g, g12 = -hs * ht / g, g22 = (1.0 + hs * hs) / g;
  return g11 * (fss - hss * q) + 2.0 * g12 * (fst - hst * q) + g22 * (ftt - htt * q);
}

// Orthonormal tangent frame for the unit normal n (any frame gives the same Laplacian: the polynomial space of
// total degree <= m and the radial weights are rotation invariant)
LPMX_HD void tangent_frame(const double* n, double* t1, double* t2) {
  // cross n with the coordinate axis it is least aligned with
  const double ax = fabs(n[0]), ay = fabs(n[1]), az = fabs(n[2]);
  double e[3] = {0.0, 0.0, 0.0};
  if (ax <= ay && ax <= az) e[0] = 1.0;
  else if (ay <= az) e[1] = 1.0;
  else e[2] = 1.0;
  t1[0] = e[1] * n[2] - e[2] * n[1];
  t1[1] = e[2] * n[0] - e[0] * n[2];
  t1[2] = e[0] * n[1] - e[1] * n[0];
  const double s = 1.0 / sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
  t1[0] *= s, t1[1] *= s, t1[2] *= s;
  t2[0] = n[1] * t1[2] - n[2] * t1[1];
  t2[1] = n[2] * t1[0] - n[0] * t1[2];
  t2[2] = n[0] * t1[1] - n[1] * t1[0];
}

struct TargetResult {
  double lap;      // Laplace-Beltrami of the data at the target (NaN when the fit is rank deficient)
  double eps;      // window radius = eps_multiplier x distance to the min_neighbors-th nearest point
  int n_neighbors; // points inside the window (the target itself included)
};

// Everything for one target: sorted index `it` of the cloud (targets and sources are collocated).
//
// OM = max(samples_order, manifold_order) and KMAX >= min_neighbors are compile-time so that every hot array has
// static indices and lives in registers.  The entries of P^T W P are products of two Taylor monomials, i.e. weighted
// MOMENTS of the neighbour set: (P^T W P)_rq = mu(ax_r + ax_q, ay_r + ay_q) / (ax_r! ay_r! ax_q! ay_q!) with
// mu(a, b) = sum_j w_j u_j^a v_j^b.  There are only (2 OM + 1)(2 OM + 2) / 2 distinct moments (45 at order 4) against
// 120 matrix entries, and the two right-hand sides are the data- and height-weighted moments of degree <= OM, so one
// neighbour costs ~120 register FMAs; the matrix is assembled once per target.
template <int OM, int KMAX>
LPMX_HD TargetResult laplacian_at_target(const Cloud& c, const Params& p, int it) {
  TargetResult out;
  const int n = c.n;
  const double x0 = c.x[it], x1 = c.x[n + it], x2 = c.x[2 * n + it];
  const int ci = cell_coord(c, x0), cj = cell_coord(c, x1), ck = cell_coord(c, x2);
  const int K = p.min_neighbors < KMAX ? p.min_neighbors : KMAX;

  // ---- pass 1: distance to the K-th nearest point (the target counts, at distance 0) ----
  double best[KMAX];  // ascending; static indices only
  double rK2 = 1e300;
  for (int ring = 1;; ++ring) {
LPMX_UNROLL
    for (int q = 0; q < KMAX; ++q) best[q] = 1e300;
    rK2 = 1e300;
    const int i0 = ci - ring < 0 ? 0 : ci - ring, i1 = ci + ring >= c.G ? c.G - 1 : ci + ring;
    const int j0 = cj - ring < 0 ? 0 : cj - ring, j1 = cj + ring >= c.G ? c.G - 1 : cj + ring;
    const int k0 = ck - ring < 0 ? 0 : ck - ring, k1 = ck + ring >= c.G ? c.G - 1 : ck + ring;
    for (int a = i0; a <= i1; ++a)
      for (int b = j0; b <= j1; ++b) {
        // cells (a, b, k0..k1) are consecutive cell indices, so their points are one contiguous sorted range
        const long base = ((long)a * c.G + b) * c.G;
        const int first = c.cell_start[base + k0], hi = c.cell_start[base + k1 + 1];
        for (int j = first; j < hi; ++j) {
          const double d0 = c.x[j] - x0, d1 = c.x[n + j] - x1, d2 = c.x[2 * n + j] - x2;
          double d = d0 * d0 + d1 * d1 + d2 * d2;
          if (d < rK2) {  // sift d into the ascending list (the displaced maximum falls off the end)
LPMX_UNROLL
            for (int q = 0; q < KMAX; ++q) {
              const double lo = d < best[q] ? d : best[q];
              d = d < best[q] ? best[q] : d;
              best[q] = lo;
            }
            rK2 = 1e300;
LPMX_UNROLL
            for (int q = 0; q < KMAX; ++q)
              if (q == K - 1) rK2 = best[q];
          }
        }
      }
    const double reach = ring * c.cell;  // every point closer than this has been seen
    const bool whole_grid = i0 == 0 && j0 == 0 && k0 == 0 && i1 == c.G - 1 && j1 == c.G - 1 && k1 == c.G - 1;
    if ((rK2 < 1e299 && rK2 <= reach * reach) || whole_grid) break;
  }
  if (!(rK2 < 1e299)) {  // fewer than K points in the whole cloud
    out.lap = NAN, out.eps = 0.0, out.n_neighbors = 0;
    return out;
  }
  // Compadre: eps = sqrt(d_K^2) * multiplier, with a floor for coincident points
  const double eps = (rK2 > 0.0 ? sqrt(rK2) : 1e-14) * p.eps_multiplier;
  const double eps2 = eps * eps, ieps = 1.0 / eps;

  // ---- pass 2: weighted moments over every point with |y - x| < eps ----
  double nrm[3] = {x0, x1, x2};
  const double inv = 1.0 / sqrt(x0 * x0 + x1 * x1 + x2 * x2);
  nrm[0] *= inv, nrm[1] *= inv, nrm[2] *= inv;
  double t1[3], t2[3];
  tangent_frame(nrm, t1, t2);
  constexpr int D = 2 * OM;
  double mu[D + 1][D + 1];    // mu[a][b], a + b <= 2 OM
  double mf[OM + 1][OM + 1];  // data-weighted, a + b <= OM
  double mh[OM + 1][OM + 1];  // height-weighted
LPMX_UNROLL
  for (int a = 0; a <= D; ++a)
LPMX_UNROLL
    for (int b = 0; b <= D; ++b) mu[a][b] = 0.0;
LPMX_UNROLL
  for (int a = 0; a <= OM; ++a)
LPMX_UNROLL
    for (int b = 0; b <= OM; ++b) mf[a][b] = 0.0, mh[a][b] = 0.0;
  int count = 0;
  {
    int ring = (int)ceil(eps * c.inv_cell);
    if (ring < 1) ring = 1;
    const int i0 = ci - ring < 0 ? 0 : ci - ring, i1 = ci + ring >= c.G ? c.G - 1 : ci + ring;
    const int j0 = cj - ring < 0 ? 0 : cj - ring, j1 = cj + ring >= c.G ? c.G - 1 : cj + ring;
    const int k0 = ck - ring < 0 ? 0 : ck - ring, k1 = ck + ring >= c.G ? c.G - 1 : ck + ring;
    for (int a = i0; a <= i1; ++a)
      for (int b = j0; b <= j1; ++b) {
        const long base = ((long)a * c.G + b) * c.G;
        const int first = c.cell_start[base + k0], hi = c.cell_start[base + k1 + 1];
        for (int j = first; j < hi; ++j) {
          const double d0 = c.x[j] - x0, d1 = c.x[n + j] - x1, d2 = c.x[2 * n + j] - x2;
          const double d = d0 * d0 + d1 * d1 + d2 * d2;
          if (!(d < eps2)) continue;  // strictly inside the window (nanoflann radius search)
          ++count;
          const double u = (d0 * t1[0] + d1 * t1[1] + d2 * t1[2]) * ieps;
          const double v = (d0 * t2[0] + d1 * t2[1] + d2 * t2[2]) * ieps;
          const double hgt = d0 * nrm[0] + d1 * nrm[1] + d2 * nrm[2];
          const double w = power_weight(sqrt(u * u + v * v), p.weight_pwr);
          const double wf = w * c.f[j], wh = w * hgt;
          double pu[D + 1], pv[D + 1];
          pu[0] = 1.0, pv[0] = 1.0;
LPMX_UNROLL
          for (int q = 1; q <= D; ++q) pu[q] = pu[q - 1] * u, pv[q] = pv[q - 1] * v;
LPMX_UNROLL
          for (int qa = 0; qa <= D; ++qa)
LPMX_UNROLL
            for (int qb = 0; qb <= D - qa; ++qb) {
              const double m = pu[qa] * pv[qb];
              mu[qa][qb] += w * m;
              if (qa + qb <= OM) {
                mf[qa][qb] += wf * m;
                mh[qa][qb] += wh * m;
              }
            }
        }
      }
  }
  out.eps = eps;
  out.n_neighbors = count;

  // ---- assemble P^T W P and the right-hand sides in Compadre's basis order (n = 0..order, ay = 0..n) ----
  constexpr int NPM = (OM + 1) * (OM + 2) / 2;
  double M[NPM * (NPM + 1) / 2], rf[NPM], rh[NPM];
  {
    constexpr double ifact[9] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0};
    int r = 0;
LPMX_UNROLL
    for (int nr = 0; nr <= OM; ++nr)
LPMX_UNROLL
      for (int ayr = 0; ayr <= nr; ++ayr) {
        const int axr = nr - ayr;
        const double sr = ifact[axr] * ifact[ayr];
        rf[r] = mf[axr][ayr] * sr;
        rh[r] = mh[axr][ayr] * sr;
        int q = 0;
LPMX_UNROLL
        for (int nq = 0; nq <= OM; ++nq)
LPMX_UNROLL
          for (int ayq = 0; ayq <= nq; ++ayq) {
            const int axq = nq - ayq;
            if (q <= r) M[tri(r, q)] = mu[axr + axq][ayr + ayq] * (sr * (ifact[axq] * ifact[ayq]));
            ++q;
          }
        ++r;
      }
  }
  const int of = p.samples_order, oh = p.manifold_order;
  const int npf = np_of(of), nph = np_of(oh);
  // The leading np x np block of P^T W P is the lower-order system (total-degree ordering), and the Cholesky factor of
  // a leading block is the leading block of the factor: one factorisation serves both orders.
  if (!cholesky_factor(NPM, M)) {
    out.lap = NAN;
    return out;
  }
  // a lower-order fit is NOT the truncation of the higher-order one, so each right-hand side is solved in its own block
  cholesky_solve(npf, M, rf);
  cholesky_solve(nph, M, rh);
  out.lap = of >= 2 ? laplace_beltrami(rf, of, rh, oh, eps) : 0.0;
  return out;
}

// ---- scalar point evaluation at targets that are NOT cloud points (remeshing) ------------------------------------------------
// gmls::Neighborhoods(src, tgt, params) + ScalarPointEvaluation / PointSample of CompadreRemesh
// (src/mesh/lpm_compadre_remesh_impl.hpp:136-210): the value of the weighted least-squares Taylor fit at the chart
// origin, i.e. coefficient 0.  The manifold reconstruction does not enter a point evaluation (the chart origin is the
// target), so only the data fits are solved: NF fields share one moment matrix and one Cholesky factorisation.
constexpr int kInterpFields = 4;
struct Fields {
  const double* f[kInterpFields];  // sorted like the cloud
};

template <int OM, int KMAX>
LPMX_HD TargetResult interpolate_at_point(const Cloud& c, const Fields& fl, const Params& p, double x0, double x1, double x2,
                                          double* value /* kInterpFields */) {
  TargetResult out;
  const int n = c.n;
  const int ci = cell_coord(c, x0), cj = cell_coord(c, x1), ck = cell_coord(c, x2);
  const int K = p.min_neighbors < KMAX ? p.min_neighbors : KMAX;
  double best[KMAX];
  double rK2 = 1e300;
  for (int ring = 1;; ++ring) {
LPMX_UNROLL
    for (int q = 0; q < KMAX; ++q) best[q] = 1e300;
    rK2 = 1e300;
    const int i0 = ci - ring < 0 ? 0 : ci - ring, i1 = ci + ring >= c.G ? c.G - 1 : ci + ring;
    const int j0 = cj - ring < 0 ? 0 : cj - ring, j1 = cj + ring >= c.G ? c.G - 1 : cj + ring;
    const int k0 = ck - ring < 0 ? 0 : ck - ring, k1 = ck + ring >= c.G ? c.G - 1 : ck + ring;
    for (int a = i0; a <= i1; ++a)
      for (int b = j0; b <= j1; ++b) {
        const long base = ((long)a * c.G + b) * c.G;
        const int first = c.cell_start[base + k0], hi = c.cell_start[base + k1 + 1];
        for (int j = first; j < hi; ++j) {
          const double d0 = c.x[j] - x0, d1 = c.x[n + j] - x1, d2 = c.x[2 * n + j] - x2;
          double d = d0 * d0 + d1 * d1 + d2 * d2;
          if (d < rK2) {
LPMX_UNROLL
            for (int q = 0; q < KMAX; ++q) {
              const double lo = d < best[q] ? d : best[q];
              d = d < best[q] ? best[q] : d;
              best[q] = lo;
            }
            rK2 = 1e300;
LPMX_UNROLL
            for (int q = 0; q < KMAX; ++q)
              if (q == K - 1) rK2 = best[q];
          }
        }
      }
    const double reach = ring * c.cell;
    const bool whole_grid = i0 == 0 && j0 == 0 && k0 == 0 && i1 == c.G - 1 && j1 == c.G - 1 && k1 == c.G - 1;
    if ((rK2 < 1e299 && rK2 <= reach * reach) || whole_grid) break;
  }
  if (!(rK2 < 1e299)) {
    for (int q = 0; q < kInterpFields; ++q) value[q] = NAN;
    out.lap = NAN, out.eps = 0.0, out.n_neighbors = 0;
    return out;
  }
  const double eps = (rK2 > 0.0 ? sqrt(rK2) : 1e-14) * p.eps_multiplier;
  const double eps2 = eps * eps, ieps = 1.0 / eps;
  double nrm[3] = {x0, x1, x2};
  const double inv = 1.0 / sqrt(x0 * x0 + x1 * x1 + x2 * x2);
  nrm[0] *= inv, nrm[1] *= inv, nrm[2] *= inv;
  double t1[3], t2[3];
  tangent_frame(nrm, t1, t2);
  constexpr int D = 2 * OM;
  double mu[D + 1][D + 1];
  double mf[kInterpFields][OM + 1][OM + 1];
LPMX_UNROLL
  for (int a = 0; a <= D; ++a)
LPMX_UNROLL
    for (int b = 0; b <= D; ++b) mu[a][b] = 0.0;
LPMX_UNROLL
  for (int q = 0; q < kInterpFields; ++q)
LPMX_UNROLL
    for (int a = 0; a <= OM; ++a)
LPMX_UNROLL
      for (int b = 0; b <= OM; ++b) mf[q][a][b] = 0.0;
  int count = 0;
  {
    int ring = (int)ceil(eps * c.inv_cell);
    if (ring < 1) ring = 1;
    const int i0 = ci - ring < 0 ? 0 : ci - ring, i1 = ci + ring >= c.G ? c.G - 1 : ci + ring;
    const int j0 = cj - ring < 0 ? 0 : cj - ring, j1 = cj + ring >= c.G ? c.G - 1 : cj + ring;
    const int k0 = ck - ring < 0 ? 0 : ck - ring, k1 = ck + ring >= c.G ? c.G - 1 : ck + ring;
    for (int a = i0; a <= i1; ++a)
      for (int b = j0; b <= j1; ++b) {
        const long base = ((long)a * c.G + b) * c.G;
        const int first = c.cell_start[base + k0], hi = c.cell_start[base + k1 + 1];
        for (int j = first; j < hi; ++j) {
          const double d0 = c.x[j] - x0, d1 = c.x[n + j] - x1, d2 = c.x[2 * n + j] - x2;
          const double d = d0 * d0 + d1 * d1 + d2 * d2;
          if (!(d < eps2)) continue;
          ++count;
          const double u = (d0 * t1[0] + d1 * t1[1] + d2 * t1[2]) * ieps;
          const double v = (d0 * t2[0] + d1 * t2[1] + d2 * t2[2]) * ieps;
          const double w = power_weight(sqrt(u * u + v * v), p.weight_pwr);
          double wf[kInterpFields];
LPMX_UNROLL
          for (int q = 0; q < kInterpFields; ++q) wf[q] = w * fl.f[q][j];
          double pu[D + 1], pv[D + 1];
          pu[0] = 1.0, pv[0] = 1.0;
LPMX_UNROLL
          for (int q = 1; q <= D; ++q) pu[q] = pu[q - 1] * u, pv[q] = pv[q - 1] * v;
LPMX_UNROLL
          for (int qa = 0; qa <= D; ++qa)
LPMX_UNROLL
            for (int qb = 0; qb <= D - qa; ++qb) {
              const double m = pu[qa] * pv[qb];
              mu[qa][qb] += w * m;
              if (qa + qb <= OM) {
LPMX_UNROLL
                for (int q = 0; q < kInterpFields; ++q) mf[q][qa][qb] += wf[q] * m;
              }
            }
        }
      }
  }
  out.eps = eps;
  out.n_neighbors = count;
  out.lap = 0.0;
  constexpr int NPM = (OM + 1) * (OM + 2) / 2;
  double M[NPM * (NPM + 1) / 2], rf[kInterpFields][NPM];
  {
    constexpr double ifact[9] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0};
    int r = 0;
LPMX_UNROLL
    for (int nr = 0; nr <= OM; ++nr)
LPMX_UNROLL
      for (int ayr = 0; ayr <= nr; ++ayr) {
        const int axr = nr - ayr;
        const double sr = ifact[axr] * ifact[ayr];
LPMX_UNROLL
        for (int q = 0; q < kInterpFields; ++q) rf[q][r] = mf[q][axr][ayr] * sr;
        int q2 = 0;
LPMX_UNROLL
        for (int nq = 0; nq <= OM; ++nq)
LPMX_UNROLL
          for (int ayq = 0; ayq <= nq; ++ayq) {
            const int axq = nq - ayq;
            if (q2 <= r) M[tri(r, q2)] = mu[axr + axq][ayr + ayq] * (sr * (ifact[axq] * ifact[ayq]));
            ++q2;
          }
        ++r;
      }
  }
  const int npf = np_of(p.samples_order);
  if (!cholesky_factor(npf, M)) {
    for (int q = 0; q < kInterpFields; ++q) value[q] = NAN;
    return out;
  }
  for (int q = 0; q < kInterpFields; ++q) {
    cholesky_solve(npf, M, rf[q]);
    value[q] = rf[q][0];
  }
  return out;
}

// run-time (orders, min_neighbors) -> compile-time instance
LPMX_HD TargetResult laplacian_at_target_dispatch(const Cloud& c, const Params& p, int it) {
  const int om = p.samples_order > p.manifold_order ? p.samples_order : p.manifold_order;
  const bool small = p.min_neighbors <= 16;
  if (om <= 2) return small ? laplacian_at_target<2, 16>(c, p, it) : laplacian_at_target<2, kMaxK>(c, p, it);
  if (om == 3) return small ? laplacian_at_target<3, 16>(c, p, it) : laplacian_at_target<3, kMaxK>(c, p, it);
  return small ? laplacian_at_target<4, 16>(c, p, it) : laplacian_at_target<4, kMaxK>(c, p, it);
}

}  // namespace gmls
}  // namespace lpmx

#endif
