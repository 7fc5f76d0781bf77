// lpmx_plane.cu -- planar (PlaneGeometry) direct sums and time steppers: SURVEY.md 8(f) row 2.
//
// Reference (as coded):
//   Incompressible2DKernels<PlaneGeometry>::kernel_vals     src/lpm_incompressible2d_kernels.hpp:55-85
//   Incompressible2DPassiveSums / ActiveSums / Tendencies   src/lpm_incompressible2d_kernels.hpp:144-279
//   Incompressible2DRK2::advance_timestep_impl              src/lpm_incompressible2d_rk2_impl.hpp:75-172
//   planar_swe_sums_rhs_pse, PlanarSwePseDirectSumReducer   src/lpm_swe_kernels.hpp:393-445, :522-571
//   PlanarSWEVertexSums / PlanarSWEFaceSums                 src/lpm_swe_kernels.hpp:626-716, :787-870
//   pse::BivariateOrder8::laplacian                         src/lpm_pse.hpp:66-73
//   SWEVorticityDivergence{Height,Area}Tendencies<Plane>    src/lpm_swe_kernels.hpp:941-1069
//   SetSurfaceFromDepth / SetDepthAndSurfaceFromMassAndArea src/lpm_swe_kernels.hpp:1079-1140
//   CoriolisBetaPlane                                       src/lpm_coriolis.hpp:93-148
//   PlanarGaussianMountain / ZeroFunctor                    src/lpm_surface_gallery.hpp:41-61, :92-102
//   SWERK4<Seed,Topo>::advance_timestep                     src/lpm_swe_rk4_impl.hpp:203-445
//   SWE<Seed>::init_direct_sums (PlaneGeometry branch)      src/lpm_swe_impl.hpp:401-422
//
// One resident state serves both planar solvers.  Targets are the concatenated list vertices-then-faces, kept as
// structure-of-arrays rows of length nt; the target "coordinate" arrays have three rows (x0, x1, surface height) so
// that the pair-sum kernel reads the target's surface height (the PSE Laplacian differences it) through the same
// 3-row view it uses on the sphere.  A SWERK4 step is 4 pair-sum launches + 5 fused O(N) stage kernels (the
// reference: 8 team-policy sums + 56 BLAS-1/functor launches); an Incompressible2DRK2 step is 2 + 3.
// Source records are leaves only, ping-pong, written by the stage kernel that produces the state they describe.
#include <cfloat>
#include <cmath>
#include <new>

#include "lpmx_finalize.cuh"
#include "lpmx_internal.h"

using namespace lpmx;

namespace lpmx {

// Padding records sit far outside any mesh with zero strength: a = |x - y|^2 stays finite and non-zero, every
// accumulated term is an exact zero (a zero record at the origin would give 0 * log(0) for a target at the origin).
constexpr double kPlanePad = 0x1p100;

enum PlaneMode : int { kModeIc2d = 0, kModeSwe = 1 };

struct PlaneState {
  lpmx_handle_t h = nullptr;
  int mode = kModeSwe;
  int nv = 0, nf = 0, nt = 0, n_leaf = 0;
  int t0 = 0, t1 = 0;
  double eps = 0, pse_eps = 1;
  int topo = LPMX_TOPO_ZERO;
  bool has_state = false;
  void* slab = nullptr;
  double *X = nullptr, *Xw = nullptr;  // 3*nt each: x0, x1, surface height (state / work state)
  double* U = nullptr;                 // 2*nt
  double *Z, *S, *T, *Zw, *Sw, *Tw;    // vorticity, divergence, third (depth at vertices / area at faces) + work
  double* K[3];                        // stage increments 1..3, 5*nt each: x0, x1, zeta, sigma, third
  double *DD, *G11, *G12, *G21, *G22, *LAPS, *PSI, *PHI, *BOT, *DEPTH, *MASS;
  unsigned char* mask = nullptr;  // nf
  int* leaf_idx = nullptr;        // nf + 1
  int* self_idx = nullptr;        // nt + 1
  double* packed[2] = {nullptr, nullptr};
  size_t packed_doubles = 0;
  int cur = 0;
  std::vector<long> tgt_off, packed_off;
  int kind() const { return mode == kModeSwe ? kPlaneSwe : kPlaneVelPsi; }
  int rec() const { return kind_rec(kind()); }
};

struct PlaneArgs {
  PartView pv;
  int t0, n_local, nv, stage, mode, topo, do_velocity;
  int with_pot;  // the evaluation carried the two potentials (kPlaneSwe) or not (kPlaneSweNoPot)
  long nt;
  double dt, f0, beta, g, ap_scale;
  double *X, *Xw, *U, *Z, *S, *T, *Zw, *Sw, *Tw, *K0, *K1, *K2;
  double *DD, *G11, *G12, *G21, *G22, *LAPS, *PSI, *PHI, *BOT, *DEPTH, *MASS;
  const unsigned char* mask;
  const int* leaf_idx;
  double* packed_next;
};

__device__ __forceinline__ double plane_topo(int topo, double x0, double x1) {
  // PlanarGaussianMountain::operator() (lpm_surface_gallery.hpp:49-52): mtn_height * exp(-b * norm2(xy))
  if (topo == LPMX_TOPO_PLANAR_GAUSSIAN_MOUNTAIN) return 0.8 * exp(-5.0 * (x0 * x0 + x1 * x1));
  return 0.0;
}

// SetSurfaceFromDepth (vertices) / SetDepthAndSurfaceFromMassAndArea (unmasked faces): returns the surface height
// of the particle for the state (x, third); masked faces keep `old_surf` (the reference never touches them).
__device__ __forceinline__ double plane_surface(const PlaneArgs& a, long g, bool is_face, double x0, double x1,
                                                double third, double old_surf) {
  if (!is_face) {
    const double b = plane_topo(a.topo, x0, x1);
    a.BOT[g] = b;
    return third + b;
  }
  if (a.mask[g - a.nv]) return old_surf;
  const double hh = a.MASS[g] / third;
  a.DEPTH[g] = hh;
  const double b = plane_topo(a.topo, x0, x1);
  a.BOT[g] = b;
  return b + hh;
}

// leaf face -> source record of the next evaluation
__device__ __forceinline__ void plane_pack(const PlaneArgs& a, long g, double x0, double x1, double zeta, double sigma,
                                           double area, double surf) {
  if (g < a.nv || !a.packed_next) return;
  const long f = g - a.nv;
  if (a.mask[f]) return;
  const double two_pi = 2 * LPMX_PI;
  if (a.mode == kModeSwe) {
    double2* r2 = reinterpret_cast<double2*>(a.packed_next + kPlaneSweRec * (size_t)a.leaf_idx[f]);
    r2[0] = make_double2(x0, x1);
    r2[1] = make_double2((zeta * area) / two_pi, (sigma * area) / two_pi);
    r2[2] = make_double2(surf, area * a.ap_scale);
  } else {
    double2* r2 = reinterpret_cast<double2*>(a.packed_next + kPlaneIc2dRec * (size_t)a.leaf_idx[f]);
    r2[0] = make_double2(x0, x1);
    r2[1] = make_double2((zeta * area) / two_pi, 0.0);
  }
}

// SWEVorticityDivergence{Height,Area}Tendencies<PlaneGeometry> with CoriolisBetaPlane (already multiplied by dt)
__device__ __forceinline__ void plane_swe_tend(const PlaneArgs& a, bool is_face, double x1, double u1, double zeta,
                                               double sigma, double third, double ddot, double laps, double* dz,
                                               double* ds, double* d3) {
  const double f = a.f0 + a.beta * x1;
  const double dfdt = a.beta * u1;
  const double gfxu = -a.beta * u1;
  *dz = (-dfdt - (zeta + f) * sigma) * a.dt;
  *ds = (f * zeta + gfxu - ddot - a.g * laps) * a.dt;
  *d3 = is_face ? (sigma * third) * a.dt : (-sigma * third) * a.dt;
}

struct PlaneSweSums {
  double u0, u1, g11, g12, g21, g22, dd, lap, psi, phi;
};
// PlanarSWEVertexSums::operator() unpacking (:693-715) on the accumulators of Pair<kPlaneSwe> (POT) / Pair<kPlaneSweNoPot>
template <bool POT>
__device__ __forceinline__ PlaneSweSums plane_swe_finalize(const double* acc) {
  PlaneSweSums r;
  r.u0 = acc[0], r.u1 = acc[1];
  r.g11 = acc[2], r.g12 = acc[3], r.g21 = acc[4], r.g22 = acc[5];
  r.dd = r.g11 * r.g11 + 2 * r.g12 * r.g21 + r.g22 * r.g22;
  r.lap = acc[6];
  r.psi = POT ? -0.5 * acc[7] : 0.0;
  r.phi = POT ? -0.5 * acc[8] : 0.0;
  return r;
}
template <bool POT>
__device__ __forceinline__ void plane_swe_store(const PlaneArgs& a, long g, const PlaneSweSums& r, bool with_vel) {
  if (with_vel) {
    a.U[g] = r.u0;
    a.U[a.nt + g] = r.u1;
  }
  a.DD[g] = r.dd;
  a.G11[g] = r.g11, a.G12[g] = r.g12, a.G21[g] = r.g21, a.G22[g] = r.g22;
  a.LAPS[g] = r.lap;
  if (POT) {  // without the potentials the stored ones stay: a later evaluation of the same call overwrites them
    a.PSI[g] = r.psi;
    a.PHI[g] = r.phi;
  }
}

// SWERK4 stages.  stage 0: k1 from the current state, work state 2.  stage 1/2: sums at the work state -> k2/k3 ->
// work state 3/4.  stage 3: sums at work state 4 -> k4 -> final update, new surfaces, new source records.
// stage 4: sums of the new state (also SWE::init_direct_sums).
// The stream function and the velocity potential are outputs of stage 4 only (the reference's stages 1-3 compute them into
// views that stage 4 overwrites, src/lpm_swe_rk4_impl.hpp:268-275,319-326,371-378,428-441), and inside a multi-step call only
// the last step's are observable: every other evaluation runs the 7-accumulator kernel without the log (POT = false).
template <bool POT>
__global__ void plane_swe_rk4_stage_kernel(const PlaneArgs a) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.t0 + li;
  const long nt = a.nt;
  const bool is_face = g >= a.nv;
  if (a.stage == 0) {
    const double x0 = a.X[g], x1 = a.X[nt + g];
    const double u0 = a.U[g], u1 = a.U[nt + g];
    const double z = a.Z[g], s = a.S[g], t = a.T[g];
    double dz, ds, d3;
    plane_swe_tend(a, is_face, x1, u1, z, s, t, a.DD[g], a.LAPS[g], &dz, &ds, &d3);  // (:216-233)
    const double k0 = a.dt * u0, k1 = a.dt * u1;
    a.K0[g] = k0, a.K0[nt + g] = k1, a.K0[2 * nt + g] = dz, a.K0[3 * nt + g] = ds, a.K0[4 * nt + g] = d3;
    const double xw0 = x0 + 0.5 * k0, xw1 = x1 + 0.5 * k1;                            // (:235-256)
    const double zw = z + 0.5 * dz, sw = s + 0.5 * ds, tw = t + 0.5 * d3;
    const double surf = plane_surface(a, g, is_face, xw0, xw1, tw, a.X[2 * nt + g]);  // (:258-266)
    a.Xw[g] = xw0, a.Xw[nt + g] = xw1, a.Xw[2 * nt + g] = surf;
    a.Zw[g] = zw, a.Sw[g] = sw, a.Tw[g] = tw;
    plane_pack(a, g, xw0, xw1, zw, sw, tw, surf);
    return;
  }
  constexpr int NACC = POT ? 9 : 7;
  double acc[NACC];
  reduce_slots<NACC>(a.pv, li, acc);
  const PlaneSweSums r = plane_swe_finalize<POT>(acc);
  if (a.stage == 4) {  // (:416-441), SWE::init_direct_sums
    plane_swe_store<POT>(a, g, r, a.do_velocity != 0);
    return;
  }
  const double xw1 = a.Xw[nt + g];
  const double zw = a.Zw[g], sw = a.Sw[g], tw = a.Tw[g];
  double dz, ds, d3;
  plane_swe_tend(a, is_face, xw1, r.u1, zw, sw, tw, r.dd, r.lap, &dz, &ds, &d3);  // (:276-289, :327-340, :379-392)
  const double x0 = a.X[g], x1 = a.X[nt + g];
  const double z = a.Z[g], s = a.S[g], t = a.T[g];
  if (a.stage < 3) {
    double* K = a.stage == 1 ? a.K1 : a.K2;
    const double k0 = a.dt * r.u0, k1 = a.dt * r.u1;  // x2 / x3 = dt * vel (:294-297, :345-346)
    K[g] = k0, K[nt + g] = k1, K[2 * nt + g] = dz, K[3 * nt + g] = ds, K[4 * nt + g] = d3;
    const double c = a.stage == 2 ? 1.0 : 0.5;
    const double nx0 = x0 + c * k0, nx1 = x1 + c * k1;
    const double nz = z + c * dz, ns = s + c * ds, ntd = t + c * d3;
    const double surf = plane_surface(a, g, is_face, nx0, nx1, ntd, a.X[2 * nt + g]);
    a.Xw[g] = nx0, a.Xw[nt + g] = nx1, a.Xw[2 * nt + g] = surf;
    a.Zw[g] = nz, a.Sw[g] = ns, a.Tw[g] = ntd;
    plane_pack(a, g, nx0, nx1, nz, ns, ntd, surf);
    return;
  }
  // stage 3: SWERK4Update (:395-414).  x4 is never assigned in the reference (zero-initialised view): as coded.
  const double third = 1.0 / 3.0, sixth = 1.0 / 6.0;
  const double x4 = 0.0;
  const double nx0 = x0 + (sixth * (a.K0[g] + x4) + third * (a.K1[g] + a.K2[g]));
  const double nx1 = x1 + (sixth * (a.K0[nt + g] + x4) + third * (a.K1[nt + g] + a.K2[nt + g]));
  const double nz = z + (sixth * (a.K0[2 * nt + g] + dz) + third * (a.K1[2 * nt + g] + a.K2[2 * nt + g]));
  const double ns = s + (sixth * (a.K0[3 * nt + g] + ds) + third * (a.K1[3 * nt + g] + a.K2[3 * nt + g]));
  const double ntd = t + (sixth * (a.K0[4 * nt + g] + d3) + third * (a.K1[4 * nt + g] + a.K2[4 * nt + g]));
  const double surf = plane_surface(a, g, is_face, nx0, nx1, ntd, a.X[2 * nt + g]);  // (:415-423)
  a.X[g] = nx0, a.X[nt + g] = nx1, a.X[2 * nt + g] = surf;
  a.Z[g] = nz, a.S[g] = ns, a.T[g] = ntd;
  plane_pack(a, g, nx0, nx1, nz, ns, ntd, surf);
}

// Incompressible2DRK2 stages (lpm_incompressible2d_rk2_impl.hpp:75-172), PlaneGeometry.  T holds the (static) area.
__global__ void plane_ic2d_rk2_stage_kernel(const PlaneArgs a) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.t0 + li;
  const long nt = a.nt;
  if (a.stage == 0) {
    const double u0 = a.U[g], u1 = a.U[nt + g];
    const double k0 = a.dt * u0, k1 = a.dt * u1;  // (:77-82)
    const double kz = -(a.beta * u1);              // Incompressible2DTendencies: -dfdt(u), no dt (:85-94)
    a.K0[g] = k0, a.K0[nt + g] = k1, a.K0[2 * nt + g] = kz;
    const double zw = a.Z[g] + a.dt * kz;          // (:97-100)
    const double xw0 = a.X[g] + a.dt * u0, xw1 = a.X[nt + g] + a.dt * u1;  // (:102-112)
    a.Xw[g] = xw0, a.Xw[nt + g] = xw1, a.Zw[g] = zw;
    plane_pack(a, g, xw0, xw1, zw, 0.0, a.T[g], 0.0);
    return;
  }
  double acc[3];
  reduce_slots<3>(a.pv, li, acc);
  const double u0 = acc[0], u1 = acc[1];
  a.U[g] = u0, a.U[nt + g] = u1;
  a.PSI[g] = -0.5 * acc[2];
  if (a.stage == 1) {
    const double k0 = a.dt * u0, k1 = a.dt * u1;  // (:130-135)
    const double kz = -(a.beta * u1);              // (:138-147)
    const double nz = a.Z[g] + ((0.5 * a.dt) * a.K0[2 * nt + g] + (0.5 * a.dt) * kz);  // (:150-154)
    const double nx0 = a.X[g] + (0.5 * a.K0[g] + 0.5 * k0);                             // (:155-160)
    const double nx1 = a.X[nt + g] + (0.5 * a.K0[nt + g] + 0.5 * k1);
    a.Z[g] = nz, a.X[g] = nx0, a.X[nt + g] = nx1;
    plane_pack(a, g, nx0, nx1, nz, 0.0, a.T[g], 0.0);
  }
}

__global__ void plane_pack_state_kernel(const PlaneArgs a) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.t0 + li;
  plane_pack(a, g, a.X[g], a.X[a.nt + g], a.Z[g], a.S[g], a.T[g], a.X[2 * a.nt + g]);
}

__global__ void plane_fill_pad_kernel(double* packed, int rec, long n_rec) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n_rec) return;
  double* r = packed + i * rec;
  r[0] = kPlanePad, r[1] = kPlanePad;
  for (int k = 2; k < rec; ++k) r[k] = 0.0;
}

struct PlaneIo {  // device pointers of one side of a set/get (nullptr = absent)
  Vec3View xy, vel;
  double *vort, *div, *third, *mass, *depth, *surf, *bottom, *ddot, *g11, *g12, *g21, *g22, *laps, *psi, *phi;
};

__global__ void plane_import_kernel(int nv, int nf, PlaneIo p, PlaneIo f, PlaneArgs a) {
  const long nt = (long)nv + nf;
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= nt) return;
  const bool face = g >= nv;
  const long i = face ? g - nv : g;
  const PlaneIo& io = face ? f : p;
  for (int k = 0; k < 2; ++k) {
    a.X[k * nt + g] = io.xy(i, k);
    a.U[k * nt + g] = io.vel.p ? io.vel(i, k) : 0.0;
  }
  a.X[2 * nt + g] = io.surf ? io.surf[i] : 0.0;
  a.Z[g] = io.vort[i];
  a.S[g] = io.div ? io.div[i] : 0.0;
  a.T[g] = io.third ? io.third[i] : 0.0;  // depth (vertices) / area (faces)
  a.MASS[g] = (face && io.mass) ? io.mass[i] : 0.0;
  a.DEPTH[g] = face ? (io.depth ? io.depth[i] : 0.0) : a.T[g];
  a.BOT[g] = io.bottom ? io.bottom[i] : 0.0;
  a.DD[g] = io.ddot ? io.ddot[i] : 0.0;
  a.G11[g] = io.g11 ? io.g11[i] : 0.0;
  a.G12[g] = io.g12 ? io.g12[i] : 0.0;
  a.G21[g] = io.g21 ? io.g21[i] : 0.0;
  a.G22[g] = io.g22 ? io.g22[i] : 0.0;
  a.LAPS[g] = io.laps ? io.laps[i] : 0.0;
  a.PSI[g] = io.psi ? io.psi[i] : 0.0;
  a.PHI[g] = io.phi ? io.phi[i] : 0.0;
}

__global__ void plane_export_kernel(int nv, int nf, PlaneIo p, PlaneIo f, PlaneArgs a) {
  const long nt = (long)nv + nf;
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= nt) return;
  const bool face = g >= nv;
  const long i = face ? g - nv : g;
  const PlaneIo& io = face ? f : p;
  for (int k = 0; k < 2; ++k) {
    if (io.xy.p) io.xy(i, k) = a.X[k * nt + g];
    if (io.vel.p) io.vel(i, k) = a.U[k * nt + g];
  }
  if (io.surf) io.surf[i] = a.X[2 * nt + g];
  if (io.vort) io.vort[i] = a.Z[g];
  if (io.div) io.div[i] = a.S[g];
  if (io.third) io.third[i] = a.T[g];
  if (face && io.mass) io.mass[i] = a.MASS[g];
  if (io.depth) io.depth[i] = face ? a.DEPTH[g] : a.T[g];
  if (io.bottom) io.bottom[i] = a.BOT[g];
  if (io.ddot) io.ddot[i] = a.DD[g];
  if (io.g11) io.g11[i] = a.G11[g];
  if (io.g12) io.g12[i] = a.G12[g];
  if (io.g21) io.g21[i] = a.G21[g];
  if (io.g22) io.g22[i] = a.G22[g];
  if (io.laps) io.laps[i] = a.LAPS[g];
  if (io.psi) io.psi[i] = a.PSI[g];
  if (io.phi) io.phi[i] = a.PHI[g];
}

__global__ void plane_self_idx_kernel(int nv, int nf, const unsigned char* mask, const int* leaf_idx, int skip_self,
                                      int* self_idx) {
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= (long)nv + nf) return;
  int v = -1;
  if (g >= nv && skip_self && !mask[g - nv]) v = leaf_idx[g - nv];
  self_idx[g] = v;
}

static Vec3View make_view2(const double* p, int layout, long ld) {
  Vec3View v;
  v.p = const_cast<double*>(p);
  if (layout == LPMX_LAYOUT_LEFT) {
    v.si = 1;
    v.sk = ld;
  } else {
    v.si = 2;
    v.sk = 1;
  }
  return v;
}
static size_t vbytes2(int layout, long ld, int n) {
  return (layout == LPMX_LAYOUT_LEFT ? (size_t)(ld + n) : (size_t)2 * n) * sizeof(double);
}

static PlaneArgs plane_args(PlaneState* s, const SumPlan* plan, const double* partials, int t_lo, int t_hi, int stage,
                            double dt, double f0, double beta, double g, double* packed_next, int do_velocity) {
  PlaneArgs a;
  a.pv = plan ? part_view(*plan, partials) : PartView{nullptr, 0, 0, 0, 1, 1};
  a.t0 = t_lo, a.n_local = t_hi - t_lo, a.nv = s->nv, a.stage = stage, a.mode = s->mode, a.topo = s->topo;
  a.do_velocity = do_velocity;
  a.with_pot = 1;
  a.nt = s->nt;
  a.dt = dt, a.f0 = f0, a.beta = beta, a.g = g;
  a.ap_scale = 1.0 / (LPMX_PI * s->pse_eps * s->pse_eps);
  a.X = s->X, a.Xw = s->Xw, a.U = s->U, a.Z = s->Z, a.S = s->S, a.T = s->T, a.Zw = s->Zw, a.Sw = s->Sw, a.Tw = s->Tw;
  a.K0 = s->K[0], a.K1 = s->K[1], a.K2 = s->K[2];
  a.DD = s->DD, a.G11 = s->G11, a.G12 = s->G12, a.G21 = s->G21, a.G22 = s->G22, a.LAPS = s->LAPS, a.PSI = s->PSI,
  a.PHI = s->PHI, a.BOT = s->BOT, a.DEPTH = s->DEPTH, a.MASS = s->MASS;
  a.mask = s->mask, a.leaf_idx = s->leaf_idx, a.packed_next = packed_next;
  return a;
}

static int plane_alloc(PlaneState* s, lpmx_handle_t h, int mode, int nv, int nf, double eps, double pse_eps, int topo) {
  if (!h || nv < 0 || nf < 0) return LPMX_ERR_INVALID;
  if (topo != LPMX_TOPO_ZERO && topo != LPMX_TOPO_PLANAR_GAUSSIAN_MOUNTAIN)
    return set_error(h, LPMX_ERR_INVALID, "unknown topography id %d", topo);
  if (mode == kModeSwe && !(pse_eps > 0)) return set_error(h, LPMX_ERR_INVALID, "pse_eps must be positive");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  s->h = h, s->mode = mode, s->nv = nv, s->nf = nf, s->nt = nv + nf, s->eps = eps, s->pse_eps = pse_eps, s->topo = topo;
  const size_t nt = s->nt;
  s->packed_doubles = (size_t)s->rec() * (size_t)(round_up_chunk(nf) + kChunk);
  const size_t dbl = 3 * nt + 3 * nt + 2 * nt + 6 * nt + 15 * nt + 11 * nt + 2 * s->packed_doubles + 64;
  const size_t bytes = dbl * sizeof(double) + sizeof(int) * (size_t)(nf + 1 + nt + 1) + (size_t)nf + 512;
  LPMX_TRY(slab_alloc(h, &s->slab, bytes));
  LPMX_CUDA(h, cudaMemsetAsync(s->slab, 0, bytes, h->stream));
  double* p = (double*)s->slab;
  s->X = p, p += 3 * nt;
  s->Xw = p, p += 3 * nt;
  s->U = p, p += 2 * nt;
  double** one[] = {&s->Z, &s->S, &s->T, &s->Zw, &s->Sw, &s->Tw};
  for (double** q : one) *q = p, p += nt;
  for (int k = 0; k < 3; ++k) s->K[k] = p, p += 5 * nt;
  double** two[] = {&s->DD, &s->G11, &s->G12, &s->G21, &s->G22, &s->LAPS, &s->PSI, &s->PHI, &s->BOT, &s->DEPTH, &s->MASS};
  for (double** q : two) *q = p, p += nt;
  p = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);  // TMA source alignment
  s->packed[0] = p, p += s->packed_doubles;
  p = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);
  s->packed[1] = p, p += s->packed_doubles;
  int* ip = (int*)p;
  s->leaf_idx = ip, ip += nf + 1;
  s->self_idx = ip, ip += nt + 1;
  s->mask = (unsigned char*)ip;
  s->t0 = (int)(((long)h->rank * s->nt) / h->world);
  s->t1 = (int)(((long)(h->rank + 1) * s->nt) / h->world);
  return LPMX_OK;
}

static int plane_exchange_rows(PlaneState* s, double* base, int n_rows) {
  if (s->h->world == 1) return LPMX_OK;
  for (int r = 0; r < n_rows; ++r) LPMX_TRY(comm_allgatherv(s->h, base + (long)r * s->nt, s->tgt_off.data()));
  return LPMX_OK;
}
static int plane_exchange_packed(PlaneState* s, double* packed) {
  if (s->h->world == 1) return LPMX_OK;
  return comm_allgatherv(s->h, packed, s->packed_off.data());
}

}  // namespace lpmx

struct lpmx_plane_solver_s {
  PlaneState st;
};

namespace lpmx {

static int plane_set_state(PlaneState* s, const PlaneIo& pu, const PlaneIo& fu, const unsigned char* mask_user,
                           int layout, long pld, long ald) {
  lpmx_handle_t h = s->h;
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if ((s->nv > 0 && !pu.xy.p) || (s->nf > 0 && (!fu.xy.p || !fu.vort || !fu.third || !mask_user)))
    return set_error(h, LPMX_ERR_INVALID, "null state array (xy; active vort, area and mask are required)");
  if (layout == LPMX_LAYOUT_LEFT && (pld < s->nv || ald < s->nf))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  PlaneIo p{}, f{};
  const void* d = nullptr;
  char name[32];
  int counter = 0;
  auto in_vec = [&](Vec3View* dst, const Vec3View& user, long ld, int n) -> int {
    snprintf(name, sizeof name, "pl_in_%d", counter++);
    if (!user.p) {
      dst->p = nullptr;
      return LPMX_OK;
    }
    LPMX_TRY(stage_in(h, name, user.p, vbytes2(layout, ld, n), &d));
    *dst = make_view2((const double*)d, layout, ld);
    return LPMX_OK;
  };
  auto in_scl = [&](double** dst, const double* user, int n) -> int {
    snprintf(name, sizeof name, "pl_in_%d", counter++);
    if (!user) {
      *dst = nullptr;
      return LPMX_OK;
    }
    LPMX_TRY(stage_in(h, name, user, sizeof(double) * (size_t)n, &d));
    *dst = (double*)d;
    return LPMX_OK;
  };
  const PlaneIo* users[2] = {&pu, &fu};
  PlaneIo* devs[2] = {&p, &f};
  for (int side = 0; side < 2; ++side) {
    const PlaneIo& u = *users[side];
    PlaneIo& o = *devs[side];
    const int n = side ? s->nf : s->nv;
    const long ld = side ? ald : pld;
    LPMX_TRY(in_vec(&o.xy, u.xy, ld, n));
    LPMX_TRY(in_vec(&o.vel, u.vel, ld, n));
    double* const* us[] = {&u.vort, &u.div, &u.third, &u.mass, &u.depth, &u.surf, &u.bottom, &u.ddot,
                           &u.g11,  &u.g12, &u.g21,   &u.g22,  &u.laps,  &u.psi,  &u.phi};
    double** os[] = {&o.vort, &o.div, &o.third, &o.mass, &o.depth, &o.surf, &o.bottom, &o.ddot,
                     &o.g11,  &o.g12, &o.g21,   &o.g22,  &o.laps,  &o.psi,  &o.phi};
    for (int k = 0; k < 15; ++k) LPMX_TRY(in_scl(os[k], *us[k], n));
  }
  if (s->nv > 0 && !p.vort) {  // passive vorticity is optional for operator-level calls: zeros
    void* z = nullptr;
    LPMX_TRY(dev_buffer(h, "pl_zero", sizeof(double) * (size_t)s->nv, &z));
    LPMX_CUDA(h, cudaMemsetAsync(z, 0, sizeof(double) * (size_t)s->nv, h->stream));
    p.vort = (double*)z;
  }
  LPMX_TRY(stage_in(h, "pl_mask", mask_user, (size_t)s->nf, &d));
  if (s->nf > 0) LPMX_CUDA(h, cudaMemcpyAsync(s->mask, d, (size_t)s->nf, cudaMemcpyDeviceToDevice, h->stream));
  const int threads = 256, blocks = (s->nt + threads - 1) / threads;
  if (s->nt > 0) {
    plane_import_kernel<<<blocks, threads, 0, h->stream>>>(s->nv, s->nf, p, f,
                                                           plane_args(s, nullptr, nullptr, 0, s->nt, 0, 0, 0, 0, 0, nullptr, 1));
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  LPMX_TRY(scan_leaves(h, s->mask, s->nf, s->leaf_idx, &s->n_leaf));
  // Incompressible2DActiveSums: collocated = FloatingPoint<Real>::zero(eps) (lpm_incompressible2d_kernels.hpp:235);
  // PlanarSWEFaceSums: collocated = !(eps > 0) (lpm_swe_kernels.hpp:848)
  const int skip = s->mode == kModeSwe ? !(s->eps > 0) : (std::fabs(s->eps) < DBL_EPSILON);
  if (s->nt > 0) {
    plane_self_idx_kernel<<<blocks, threads, 0, h->stream>>>(s->nv, s->nf, s->mask, s->leaf_idx, skip, s->self_idx);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  const long n_rec = (long)(s->packed_doubles / s->rec());
  for (int b = 0; b < 2; ++b) {
    plane_fill_pad_kernel<<<(int)((n_rec + 255) / 256), 256, 0, h->stream>>>(s->packed[b], s->rec(), n_rec);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  const int W = h->world;
  s->tgt_off.assign(W + 1, 0);
  s->packed_off.assign(W + 1, 0);
  std::vector<int> leaf_host;
  if (W > 1 && s->nf > 0) {
    leaf_host.resize(s->nf);
    LPMX_CUDA(h, cudaMemcpyAsync(leaf_host.data(), s->leaf_idx, sizeof(int) * s->nf, cudaMemcpyDeviceToHost, h->stream));
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  for (int r = 0; r <= W; ++r) {
    const long t = ((long)r * s->nt) / W;
    s->tgt_off[r] = t;
    long ff = t - s->nv;
    if (ff < 0) ff = 0;
    long l = (ff >= s->nf) ? s->n_leaf : (W > 1 ? leaf_host[ff] : 0);
    if (r == W) l = s->n_leaf;
    s->packed_off[r] = (long)s->rec() * l;
  }
  s->t0 = (int)s->tgt_off[h->rank];
  s->t1 = (int)s->tgt_off[h->rank + 1];
  s->cur = 0;
  s->has_state = true;
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));  // staging buffers may be reused by the next call
  return LPMX_OK;
}

static int plane_get_state(PlaneState* s, const PlaneIo& pu, const PlaneIo& fu, int layout, long pld, long ald) {
  lpmx_handle_t h = s->h;
  if (!s->has_state) return set_error(h, LPMX_ERR_STATE, "get_state before set_state");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if (layout == LPMX_LAYOUT_LEFT && (pld < s->nv || ald < s->nf))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  if (h->world > 1) {  // every rank returns the full state
    LPMX_TRY(plane_exchange_rows(s, s->X, 3));
    LPMX_TRY(plane_exchange_rows(s, s->U, 2));
    for (double* row : {s->Z, s->S, s->T, s->DD, s->G11, s->G12, s->G21, s->G22, s->LAPS, s->PSI, s->PHI, s->BOT, s->DEPTH})
      LPMX_TRY(plane_exchange_rows(s, row, 1));
  }
  struct Out {
    void* user;
    void* dev;
    size_t bytes;
  };
  std::vector<Out> outs;
  PlaneIo p{}, f{};
  void* d = nullptr;
  char name[32];
  int counter = 0;
  auto out_vec = [&](Vec3View* dst, const Vec3View& user, long ld, int n) -> int {
    snprintf(name, sizeof name, "pl_out_%d", counter++);
    dst->p = nullptr;
    if (!user.p || n == 0) return LPMX_OK;
    LPMX_TRY(stage_out_begin(h, name, user.p, vbytes2(layout, ld, n), &d));
    *dst = make_view2((const double*)d, layout, ld);
    outs.push_back({user.p, d, vbytes2(layout, ld, n)});
    return LPMX_OK;
  };
  auto out_scl = [&](double** dst, double* user, int n) -> int {
    snprintf(name, sizeof name, "pl_out_%d", counter++);
    *dst = nullptr;
    if (!user || n == 0) return LPMX_OK;
    LPMX_TRY(stage_out_begin(h, name, user, sizeof(double) * (size_t)n, &d));
    *dst = (double*)d;
    outs.push_back({user, d, sizeof(double) * (size_t)n});
    return LPMX_OK;
  };
  const PlaneIo* users[2] = {&pu, &fu};
  PlaneIo* devs[2] = {&p, &f};
  for (int side = 0; side < 2; ++side) {
    const PlaneIo& u = *users[side];
    PlaneIo& o = *devs[side];
    const int n = side ? s->nf : s->nv;
    const long ld = side ? ald : pld;
    LPMX_TRY(out_vec(&o.xy, u.xy, ld, n));
    LPMX_TRY(out_vec(&o.vel, u.vel, ld, n));
    double* const us[] = {u.vort, u.div, u.third, u.mass, u.depth, u.surf, u.bottom, u.ddot,
                          u.g11,  u.g12, u.g21,   u.g22,  u.laps,  u.psi,  u.phi};
    double** os[] = {&o.vort, &o.div, &o.third, &o.mass, &o.depth, &o.surf, &o.bottom, &o.ddot,
                     &o.g11,  &o.g12, &o.g21,   &o.g22,  &o.laps,  &o.psi,  &o.phi};
    for (int k = 0; k < 15; ++k) LPMX_TRY(out_scl(os[k], us[k], n));
  }
  if (s->nt > 0) {
    const int threads = 256, blocks = (s->nt + threads - 1) / threads;
    plane_export_kernel<<<blocks, threads, 0, h->stream>>>(s->nv, s->nf, p, f,
                                                           plane_args(s, nullptr, nullptr, 0, s->nt, 0, 0, 0, 0, 0, nullptr, 1));
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  bool any_host = false;
  for (const Out& o : outs) {
    if (o.user != o.dev) any_host = true;
    LPMX_TRY(stage_out_end(h, o.user, o.dev, o.bytes));
  }
  if (any_host) LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

static int plane_launch_stage(PlaneState* s, const PlaneArgs& a) {
  lpmx_handle_t h = s->h;
  if (a.n_local <= 0) return LPMX_OK;
  const int threads = 128, blocks = (a.n_local + threads - 1) / threads;
  if (s->mode == kModeSwe && a.with_pot)
    plane_swe_rk4_stage_kernel<true><<<blocks, threads, 0, h->stream>>>(a);
  else if (s->mode == kModeSwe)
    plane_swe_rk4_stage_kernel<false><<<blocks, threads, 0, h->stream>>>(a);
  else
    plane_ic2d_rk2_stage_kernel<<<blocks, threads, 0, h->stream>>>(a);
  ++h->launches;
  return check_cuda(h, cudaGetLastError(), "planar stage kernel launch");
}

// pair sums of targets [lo, hi) (global indices of the concatenated list) at the coordinates in `tgt_base`
// (3 rows of nt) against the current source records
static int plane_pair_sum(PlaneState* s, double* tgt_base, int lo, int hi, int kind, SumPlan* plan, double** partials) {
  lpmx_handle_t h = s->h;
  LPMX_TRY(make_plan(h, kind, hi - lo, s->n_leaf, plan));
  void* part = nullptr;
  LPMX_TRY(dev_buffer(h, "plane_partials", plan_partials_bytes(*plan) + 256, &part));
  *partials = (double*)part;
  Vec3View v;
  v.p = tgt_base + lo;
  v.si = 1;
  v.sk = s->nt;
  return launch_pair_sum(h, *plan, v, s->self_idx + lo, s->packed[s->cur], s->eps * s->eps, *partials,
                         1.0 / (s->pse_eps * s->pse_eps));
}

// direct sums of the resident state for targets [lo, hi) intersected with this rank's range:
// SWE::init_direct_sums / Incompressible2D::init_direct_sums and the operator-level entry points
static int plane_eval_state(PlaneState* s, int lo, int hi, int do_velocity) {
  lpmx_handle_t h = s->h;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  PlaneArgs pk = plane_args(s, nullptr, nullptr, s->t0, s->t1, 0, 0, 0, 0, 0, s->packed[s->cur], 1);
  if (pk.n_local > 0) {
    const int threads = 256, blocks = (pk.n_local + threads - 1) / threads;
    plane_pack_state_kernel<<<blocks, threads, 0, h->stream>>>(pk);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  LPMX_TRY(plane_exchange_packed(s, s->packed[s->cur]));
  lo = std::max(lo, s->t0);
  hi = std::min(hi, s->t1);
  if (hi <= lo) return LPMX_OK;
  SumPlan plan;
  double* part = nullptr;
  LPMX_TRY(plane_pair_sum(s, s->X, lo, hi, s->kind(), &plan, &part));
  const int final_stage = s->mode == kModeSwe ? 4 : 2;
  return plane_launch_stage(s, plane_args(s, &plan, part, lo, hi, final_stage, 0, 0, 0, 0, nullptr, do_velocity));
}

static int plane_advance(PlaneState* s, double dt, double f0, double beta, double g, int n_steps) {
  lpmx_handle_t h = s->h;
  if (!s->has_state) return set_error(h, LPMX_ERR_STATE, "advance before set_state");
  if (n_steps < 0) return set_error(h, LPMX_ERR_INVALID, "negative step count");
  if (n_steps == 0 || s->nt == 0) return LPMX_OK;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  SumPlan plan;
  double* part = nullptr;
  const int n_eval = s->mode == kModeSwe ? 4 : 2;
  for (int step = 0; step < n_steps; ++step) {
    // stage 0: first increments from the current state, first work state and its source records
    LPMX_TRY(plane_launch_stage(s, plane_args(s, nullptr, nullptr, s->t0, s->t1, 0, dt, f0, beta, g, s->packed[s->cur ^ 1], 1)));
    for (int e = 1; e <= n_eval; ++e) {
      s->cur ^= 1;
      LPMX_TRY(plane_exchange_packed(s, s->packed[s->cur]));
      const bool last = e == n_eval;  // the last evaluation is at the new state
      // SWE: psi and phi are wanted from the new-state evaluation of the call's last step only
      const bool with_pot = s->mode != kModeSwe || (last && step + 1 == n_steps);
      LPMX_TRY(plane_pair_sum(s, last ? s->X : s->Xw, s->t0, s->t1, with_pot ? s->kind() : kPlaneSweNoPot, &plan, &part));
      PlaneArgs a = plane_args(s, &plan, part, s->t0, s->t1, e, dt, f0, beta, g, last ? nullptr : s->packed[s->cur ^ 1], 1);
      a.with_pot = with_pot ? 1 : 0;
      LPMX_TRY(plane_launch_stage(s, a));
    }
  }
  return LPMX_OK;
}

static PlaneIo io_of(const lpmx_plane_swe_passive_t* P, int layout, long ld) {
  PlaneIo o{};
  if (!P) return o;
  o.xy = make_view2(P->xy, layout, ld);
  o.vel = make_view2(P->vel, layout, ld);
  o.vort = P->vort, o.div = P->div, o.third = P->depth, o.surf = P->surf, o.bottom = P->bottom, o.ddot = P->ddot;
  o.g11 = P->du1dx1, o.g12 = P->du1dx2, o.g21 = P->du2dx1, o.g22 = P->du2dx2, o.laps = P->laps, o.psi = P->psi, o.phi = P->phi;
  return o;
}
static PlaneIo io_of(const lpmx_plane_swe_active_t* A, int layout, long ld) {
  PlaneIo o{};
  if (!A) return o;
  o.xy = make_view2(A->xy, layout, ld);
  o.vel = make_view2(A->vel, layout, ld);
  o.vort = A->vort, o.div = A->div, o.third = A->area, o.mass = A->mass, o.depth = A->depth, o.surf = A->surf;
  o.bottom = A->bottom, o.ddot = A->ddot, o.g11 = A->du1dx1, o.g12 = A->du1dx2, o.g21 = A->du2dx1, o.g22 = A->du2dx2;
  o.laps = A->laps, o.psi = A->psi, o.phi = A->phi;
  return o;
}

static void plane_free(lpmx_plane_solver_s* s) {
  if (!s) return;
  if (s->st.h && s->st.h->cached_plane == s) s->st.h->cached_plane = nullptr;
  if (s->st.slab) {
    cudaSetDevice(s->st.h->device);
    cudaStreamSynchronize(s->st.h->stream);
    slab_free(s->st.h, s->st.slab);
  }
  delete s;
}

static int plane_create(lpmx_handle_t h, int mode, int nv, int nf, double eps, double pse_eps, int topo,
                        lpmx_plane_solver_t* out) {
  if (!h || !out) return LPMX_ERR_INVALID;
  lpmx_plane_solver_s* s = new (std::nothrow) lpmx_plane_solver_s;
  if (!s) return LPMX_ERR_NOMEM;
  const int rc = plane_alloc(&s->st, h, mode, nv, nf, eps, pse_eps, topo);
  if (rc != LPMX_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return LPMX_OK;
}

// the handle's cached solver for the in-place and operator-level entry points
static int plane_cached(lpmx_handle_t h, int mode, int nv, int nf, double eps, double pse_eps, int topo,
                        lpmx_plane_solver_t* out) {
  lpmx_plane_solver_t s = h->cached_plane;
  if (!s || s->st.mode != mode || s->st.nv != nv || s->st.nf != nf || s->st.eps != eps || s->st.pse_eps != pse_eps ||
      s->st.topo != topo) {
    if (s) plane_free(s);
    h->cached_plane = nullptr;
    LPMX_TRY(plane_create(h, mode, nv, nf, eps, pse_eps, topo, &s));
    h->cached_plane = s;
  }
  *out = s;
  return LPMX_OK;
}

}  // namespace lpmx

extern "C" {

int lpmx_plane_swe_solver_create(lpmx_handle_t h, int n_passive, int n_active, double eps, double pse_eps, int topo,
                                 lpmx_plane_solver_t* s) {
  return plane_create(h, kModeSwe, n_passive, n_active, eps, pse_eps, topo, s);
}

int lpmx_plane_swe_solver_destroy(lpmx_plane_solver_t s) {
  plane_free(s);
  return LPMX_OK;
}

int lpmx_plane_swe_solver_set_state(lpmx_plane_solver_t s, const lpmx_plane_swe_passive_t* passive,
                                    const lpmx_plane_swe_active_t* active, int layout, long passive_ld, long active_ld) {
  if (!s) return LPMX_ERR_INVALID;
  if (!passive || !active) return set_error(s->st.h, LPMX_ERR_INVALID, "null state struct");
  return plane_set_state(&s->st, io_of(passive, layout, passive_ld), io_of(active, layout, active_ld), active->mask, layout,
                         passive_ld, active_ld);
}

int lpmx_plane_swe_solver_get_state(lpmx_plane_solver_t s, const lpmx_plane_swe_passive_t* passive,
                                    const lpmx_plane_swe_active_t* active, int layout, long passive_ld, long active_ld) {
  if (!s) return LPMX_ERR_INVALID;
  if (!passive || !active) return set_error(s->st.h, LPMX_ERR_INVALID, "null state struct");
  return plane_get_state(&s->st, io_of(passive, layout, passive_ld), io_of(active, layout, active_ld), layout, passive_ld,
                         active_ld);
}

int lpmx_plane_swe_solver_init_direct_sums(lpmx_plane_solver_t s, int do_velocity) {
  if (!s) return LPMX_ERR_INVALID;
  if (!s->st.has_state) return set_error(s->st.h, LPMX_ERR_STATE, "init_direct_sums before set_state");
  return plane_eval_state(&s->st, 0, s->st.nt, do_velocity);
}

int lpmx_plane_swe_solver_advance(lpmx_plane_solver_t s, double dt, double f0, double beta, double g, int n_steps) {
  if (!s) return LPMX_ERR_INVALID;
  return plane_advance(&s->st, dt, f0, beta, g, n_steps);
}

int lpmx_plane_swe_solver_interactions_per_eval(lpmx_plane_solver_t s, double* local, double* global) {
  if (!s) return LPMX_ERR_INVALID;
  const PlaneState& st = s->st;
  if (!st.has_state) return set_error(st.h, LPMX_ERR_STATE, "no state");
  const bool skip = !(st.eps > 0);
  // the n_leaf collocated self pairs are spread over the ranks that own leaf faces; attribute them to the global count
  if (global) *global = (double)st.nt * st.n_leaf - (skip ? st.n_leaf : 0);
  if (local) {
    const double share = st.nt > 0 ? (double)(st.t1 - st.t0) / st.nt : 0.0;
    *local = (double)(st.t1 - st.t0) * st.n_leaf - (skip ? share * st.n_leaf : 0.0);
  }
  return LPMX_OK;
}

int lpmx_swe_plane_rk4_step(lpmx_handle_t h, double dt, double f0, double beta, double g, double eps, double pse_eps,
                            int topo, int n_passive, const lpmx_plane_swe_passive_t* passive, int n_active,
                            const lpmx_plane_swe_active_t* active, int layout, long passive_ld, long active_ld,
                            int n_steps) {
  if (!h) return LPMX_ERR_INVALID;
  if (!passive || !active) return set_error(h, LPMX_ERR_INVALID, "null state struct");
  if ((n_passive > 0 && (!passive->vort || !passive->depth || !passive->vel || !passive->ddot || !passive->laps)) ||
      (n_active > 0 && (!active->mass || !active->vel || !active->ddot || !active->laps)))
    return set_error(h, LPMX_ERR_INVALID,
                     "vort, depth/mass, vel, ddot and laps are required (SWE::init_direct_sums output)");
  lpmx_plane_solver_t s = nullptr;
  LPMX_TRY(plane_cached(h, kModeSwe, n_passive, n_active, eps, pse_eps, topo, &s));
  const PlaneIo p = io_of(passive, layout, passive_ld), a = io_of(active, layout, active_ld);
  LPMX_TRY(plane_set_state(&s->st, p, a, active->mask, layout, passive_ld, active_ld));
  LPMX_TRY(plane_advance(&s->st, dt, f0, beta, g, n_steps));
  return plane_get_state(&s->st, p, a, layout, passive_ld, active_ld);
}

int lpmx_swe_plane_sums(lpmx_handle_t h, const double* tgt_xy, int tgt_layout, long tgt_ld, const double* tgt_surf,
                        int n_tgt, const double* src_xy, int src_layout, long src_ld, const double* src_vort,
                        const double* src_div, const double* src_area, const unsigned char* src_mask,
                        const double* src_surf, int n_src, double eps, double pse_eps, int targets_are_sources,
                        int do_velocity, const lpmx_plane_swe_sums_t* out) {
  if (!h) return LPMX_ERR_INVALID;
  if (!out) return set_error(h, LPMX_ERR_INVALID, "null output struct");
  if (n_src < 0 || n_tgt < 0) return set_error(h, LPMX_ERR_INVALID, "negative size");
  if (n_src > 0 && (!src_xy || !src_vort || !src_div || !src_area || !src_mask || !src_surf))
    return set_error(h, LPMX_ERR_INVALID, "null source array");
  if (targets_are_sources && n_tgt != n_src) return set_error(h, LPMX_ERR_INVALID, "collocated call needs n_tgt == n_src");
  if (!targets_are_sources && n_tgt > 0 && (!tgt_xy || !tgt_surf)) return set_error(h, LPMX_ERR_INVALID, "null target array");
  if (!targets_are_sources && tgt_layout != src_layout)
    return set_error(h, LPMX_ERR_INVALID, "targets and sources must share one layout");
  if (do_velocity && n_tgt > 0 && !out->vel) return set_error(h, LPMX_ERR_INVALID, "null velocity output");
  if (n_tgt == 0) return LPMX_OK;
  const int nv = targets_are_sources ? 0 : n_tgt;
  lpmx_plane_solver_t s = nullptr;
  LPMX_TRY(plane_cached(h, kModeSwe, nv, n_src, eps, pse_eps, LPMX_TOPO_ZERO, &s));
  PlaneIo p{}, a{};
  if (nv > 0) {
    p.xy = make_view2(tgt_xy, src_layout, tgt_ld);
    p.surf = const_cast<double*>(tgt_surf);
  }
  a.xy = make_view2(src_xy, src_layout, src_ld);
  a.vort = const_cast<double*>(src_vort), a.div = const_cast<double*>(src_div), a.third = const_cast<double*>(src_area);
  a.surf = const_cast<double*>(src_surf);
  LPMX_TRY(plane_set_state(&s->st, p, a, src_mask, src_layout, tgt_ld, src_ld));
  const int lo = targets_are_sources ? 0 : 0, hi = targets_are_sources ? n_src : nv;
  LPMX_TRY(plane_eval_state(&s->st, lo, hi, do_velocity));
  PlaneIo o{};
  o.vel = make_view2(do_velocity ? out->vel : nullptr, src_layout, targets_are_sources ? src_ld : tgt_ld);
  o.ddot = out->ddot, o.g11 = out->du1dx1, o.g12 = out->du1dx2, o.g21 = out->du2dx1, o.g22 = out->du2dx2;
  o.laps = out->laps, o.psi = out->psi, o.phi = out->phi;
  PlaneIo none{};
  return targets_are_sources ? plane_get_state(&s->st, none, o, src_layout, tgt_ld, src_ld)
                             : plane_get_state(&s->st, o, none, src_layout, tgt_ld, src_ld);
}

int lpmx_ic2d_plane_sums(lpmx_handle_t h, const double* tgt_xy, int tgt_layout, long tgt_ld, int n_tgt,
                         const double* src_xy, int src_layout, long src_ld, const double* src_vort,
                         const double* src_area, const unsigned char* src_mask, int n_src, double eps,
                         int targets_are_sources, double* out_vel, double* out_psi) {
  if (!h) return LPMX_ERR_INVALID;
  if (n_src < 0 || n_tgt < 0) return set_error(h, LPMX_ERR_INVALID, "negative size");
  if (n_src > 0 && (!src_xy || !src_vort || !src_area || !src_mask)) return set_error(h, LPMX_ERR_INVALID, "null source array");
  if (targets_are_sources && n_tgt != n_src) return set_error(h, LPMX_ERR_INVALID, "collocated call needs n_tgt == n_src");
  if (!targets_are_sources && n_tgt > 0 && !tgt_xy) return set_error(h, LPMX_ERR_INVALID, "null target array");
  if (!targets_are_sources && tgt_layout != src_layout)
    return set_error(h, LPMX_ERR_INVALID, "targets and sources must share one layout");
  if (!out_vel && n_tgt > 0) return set_error(h, LPMX_ERR_INVALID, "null output");
  if (n_tgt == 0) return LPMX_OK;
  const int nv = targets_are_sources ? 0 : n_tgt;
  lpmx_plane_solver_t s = nullptr;
  LPMX_TRY(plane_cached(h, kModeIc2d, nv, n_src, eps, 1.0, LPMX_TOPO_ZERO, &s));
  PlaneIo p{}, a{};
  if (nv > 0) p.xy = make_view2(tgt_xy, src_layout, tgt_ld);
  a.xy = make_view2(src_xy, src_layout, src_ld);
  a.vort = const_cast<double*>(src_vort), a.third = const_cast<double*>(src_area);
  LPMX_TRY(plane_set_state(&s->st, p, a, src_mask, src_layout, tgt_ld, src_ld));
  LPMX_TRY(plane_eval_state(&s->st, 0, targets_are_sources ? n_src : nv, 1));
  PlaneIo o{}, none{};
  o.vel = make_view2(out_vel, src_layout, targets_are_sources ? src_ld : tgt_ld);
  o.psi = out_psi;
  return targets_are_sources ? plane_get_state(&s->st, none, o, src_layout, tgt_ld, src_ld)
                             : plane_get_state(&s->st, o, none, src_layout, tgt_ld, src_ld);
}

int lpmx_ic2d_plane_rk2_step(lpmx_handle_t h, double dt, double f0, double beta, double eps, int n_passive,
                             double* passive_xy, double* passive_vort, double* passive_vel, double* passive_psi,
                             int n_active, double* active_xy, double* active_vort, double* active_vel,
                             double* active_psi, const double* active_area, const unsigned char* active_mask,
                             int layout, long passive_ld, long active_ld, int n_steps) {
  if (!h) return LPMX_ERR_INVALID;
  if ((n_passive > 0 && (!passive_xy || !passive_vort || !passive_vel)) ||
      (n_active > 0 && (!active_xy || !active_vort || !active_vel || !active_area || !active_mask)))
    return set_error(h, LPMX_ERR_INVALID, "null state array");
  lpmx_plane_solver_t s = nullptr;
  LPMX_TRY(plane_cached(h, kModeIc2d, n_passive, n_active, eps, 1.0, LPMX_TOPO_ZERO, &s));
  PlaneIo p{}, a{};
  p.xy = make_view2(passive_xy, layout, passive_ld), p.vel = make_view2(passive_vel, layout, passive_ld);
  p.vort = passive_vort, p.psi = passive_psi;
  a.xy = make_view2(active_xy, layout, active_ld), a.vel = make_view2(active_vel, layout, active_ld);
  a.vort = active_vort, a.psi = active_psi, a.third = const_cast<double*>(active_area);
  LPMX_TRY(plane_set_state(&s->st, p, a, active_mask, layout, passive_ld, active_ld));
  LPMX_TRY(plane_advance(&s->st, dt, f0, beta, 0.0, n_steps));
  a.third = nullptr;  // the area is an input only
  return plane_get_state(&s->st, p, a, layout, passive_ld, active_ld);
}

}  // extern "C"
