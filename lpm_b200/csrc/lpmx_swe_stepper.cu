// lpmx_swe_stepper.cu -- device-resident SWE state and the SWERK2 stepper on the sphere.
//
// Reference (as coded -- SURVEY.md 8(a) rows C5-C7):
//   SWERK2::advance_timestep_impl                   src/lpm_swe_rk2_impl.hpp:80-258
//   SWEVorticityDivergenceHeightTendencies          src/lpm_swe_kernels.hpp:941-999   (passive particles)
//   SWEVorticityDivergenceAreaTendencies            src/lpm_swe_kernels.hpp:1010-1069 (active particles)
//   SetSurfaceFromDepth                             src/lpm_swe_kernels.hpp:1079-1100
//   SetDepthAndSurfaceFromMassAndArea               src/lpm_swe_kernels.hpp:1110-1140
//   CoriolisSphere::f / dfdt / grad_f_cross_u       src/lpm_coriolis.hpp:154-195
//   SWE::init_direct_sums                           src/lpm_swe_impl.hpp:401-445
//
// The GMLS surface Laplacian of the reference (Compadre, host kd-tree; rk2_impl.hpp:134-154,233-252) is NOT part
// of this path: it enters through lpmx_swe_laplacian_fn.  Everything else of a step is 2 pair-sum launches + 3
// fused O(N) kernels: stage 0 builds the predictor state, its surfaces and its packed source records; stage 1
// turns the predictor sums into (u, ddot), the stage-2 tendencies, the Heun combine, the new surfaces and the new
// source records; stage 2 turns the final sums into (u, ddot) of the new state.
// Bottom topography: ZeroFunctor (the only one a config uses, examples/sphere_swe_tc2.cpp:41; the reference's
// setters default-construct their functor anyway, quirk C-v).
#include <cfloat>
#include <cmath>
#include <new>

#include "lpmx_finalize.cuh"
#include "lpmx_internal.h"

using namespace lpmx;

namespace lpmx {

constexpr int kSweRec = 6;

struct SweState {
  lpmx_handle_t h = nullptr;
  int nv = 0, nf = 0, nt = 0, n_leaf = 0;
  int t0 = 0, t1 = 0;
  double eps = 0;
  bool has_state = false;
  void* slab = nullptr;
  // SoA over the concatenated target list (vertices then faces)
  double *X, *U, *Xw, *K1x;                    // 3*nt each
  double *Z, *S, *T, *Zw, *Sw, *Tw;            // vorticity, divergence, third (depth at vertices / area at faces) + work
  double *K1z, *K1s, *K1t;                     // stage-1 increments
  double *DD, *LAPS, *SURF, *BOT, *DEPTH, *MASS;  // DEPTH/MASS are used for faces only (vertex depth lives in T)
  unsigned char* mask = nullptr;               // nf
  int* leaf_idx = nullptr;                     // nf + 1
  int* self_idx = nullptr;                     // nt + 1
  double* packed[2] = {nullptr, nullptr};
  int cur = 0;
  double* partials = nullptr;
  std::vector<long> tgt_off, packed_off;
  SumPlan plan;
  Vec3View local_view(double* base) const {
    Vec3View v;
    v.p = base + t0;
    v.si = 1;
    v.sk = nt;
    return v;
  }
};

struct SweStageArgs {
  PartView pv;
  int t0, n_local, nv, stage;
  long nt;
  double dt, Omega, g;
  double *X, *U, *Xw, *K1x, *Z, *S, *T, *Zw, *Sw, *Tw, *K1z, *K1s, *K1t, *DD, *LAPS, *SURF, *BOT, *DEPTH, *MASS;
  const unsigned char* mask;
  const int* leaf_idx;
  double* packed_next;
};

// one particle's tendencies (:941-1069): returns dzeta, dsigma, dthird (already multiplied by dt)
__device__ __forceinline__ void swe_tend(const SweStageArgs& a, bool is_face, const double* x, const double* u, double zeta,
                                         double sigma, double third, double ddot, double laps, double* dz, double* ds,
                                         double* d3) {
  const double f = 2 * a.Omega * x[2];
  *dz = (-(2 * a.Omega * u[2]) - (zeta + f) * sigma) * a.dt;
  const double gfxu = -2 * a.Omega * (-u[0] * x[1] + u[1] * x[0]);
  const double n2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  *ds = (f * zeta + gfxu - ddot - a.g * laps - n2) * a.dt;
  *d3 = is_face ? (sigma * third) * a.dt : (-sigma * third) * a.dt;
}

// surfaces (:1079-1140) for the state (third = depth at vertices, area at faces); ZeroFunctor bottom
__device__ __forceinline__ void swe_surfaces(const SweStageArgs& a, long g, bool is_face, double third) {
  if (!is_face) {
    a.BOT[g] = 0.0;
    a.SURF[g] = third + 0.0;
  } else if (!a.mask[g - a.nv]) {
    const double hh = a.MASS[g] / third;
    a.DEPTH[g] = hh;
    a.BOT[g] = 0.0;
    a.SURF[g] = 0.0 + hh;
  }
}

__device__ __forceinline__ void swe_pack(const SweStageArgs& a, long g, const double* x, double zeta, double sigma,
                                         double area) {
  if (g < a.nv || !a.packed_next) return;
  const long f = g - a.nv;
  if (a.mask[f]) return;
  double* rec = a.packed_next + kSweRec * (size_t)a.leaf_idx[f];
  double2* r2 = reinterpret_cast<double2*>(rec);
  r2[0] = make_double2(x[0], x[1]);
  r2[1] = make_double2(x[2], gamma_of(zeta, area));
  r2[2] = make_double2(gamma_of(sigma, area), 0.0);
}

__global__ void swe_rk2_stage_kernel(const SweStageArgs a) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.t0 + li;
  const long nt = a.nt;
  const bool is_face = g >= a.nv;
  if (a.stage == 0) {
    // stage 1 tendencies from the current state (:85-104) and the predictor state (:108-132)
    const double x[3] = {a.X[g], a.X[nt + g], a.X[2 * nt + g]};
    const double u[3] = {a.U[g], a.U[nt + g], a.U[2 * nt + g]};
    double dz, ds, d3, xw[3];
    swe_tend(a, is_face, x, u, a.Z[g], a.S[g], a.T[g], a.DD[g], a.LAPS[g], &dz, &ds, &d3);
    a.K1z[g] = dz, a.K1s[g] = ds, a.K1t[g] = d3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double x1 = a.dt * u[k];
      a.K1x[k * nt + g] = x1;
      xw[k] = x[k] + x1;
      a.Xw[k * nt + g] = xw[k];
    }
    const double zw = a.Z[g] + dz, sw = a.S[g] + ds, tw = a.T[g] + d3;
    a.Zw[g] = zw, a.Sw[g] = sw, a.Tw[g] = tw;
    swe_surfaces(a, g, is_face, tw);
    swe_pack(a, g, xw, zw, sw, tw);
    return;
  }
  double acc[15];
  reduce_slots<15>(a.pv, li, acc);
  if (a.stage == 1) {
    const double xw[3] = {a.Xw[g], a.Xw[nt + g], a.Xw[2 * nt + g]};
    double u[3], g9[9];
    const double dd = swe_finalize(acc, xw, u, g9);
    a.DD[g] = dd;  // the reference overwrites velocity and double dot here (:157-168)
    double dz2, ds2, d32;
    swe_tend(a, is_face, xw, u, a.Zw[g], a.Sw[g], a.Tw[g], dd, a.LAPS[g], &dz2, &ds2, &d32);  // (:170-185)
    double xn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a.U[k * nt + g] = u[k];
      const double x2 = a.dt * u[k];                                          // (:187-188)
      xn[k] = a.X[k * nt + g] + (0.5 * a.K1x[k * nt + g] + 0.5 * x2);          // Heun combine (:190-205)
      a.X[k * nt + g] = xn[k];
    }
    const double zn = a.Z[g] + (0.5 * a.K1z[g] + 0.5 * dz2);
    const double sn = a.S[g] + (0.5 * a.K1s[g] + 0.5 * ds2);
    const double tn = a.T[g] + (0.5 * a.K1t[g] + 0.5 * d32);
    a.Z[g] = zn, a.S[g] = sn, a.T[g] = tn;
    swe_surfaces(a, g, is_face, tn);                                           // (:207-216)
    swe_pack(a, g, xn, zn, sn, tn);
  } else {  // stage 2: velocity and double dot of the new state (:218-231)
    const double x[3] = {a.X[g], a.X[nt + g], a.X[2 * nt + g]};
    double u[3], g9[9];
    a.DD[g] = swe_finalize(acc, x, u, g9);
#pragma unroll
    for (int k = 0; k < 3; ++k) a.U[k * nt + g] = u[k];
  }
}

// SWE::init_direct_sums finalize: velocity only if do_velocity
__global__ void swe_init_sums_kernel(const SweStageArgs a, int do_velocity) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.t0 + li;
  const long nt = a.nt;
  double acc[15];
  reduce_slots<15>(a.pv, li, acc);
  const double x[3] = {a.X[g], a.X[nt + g], a.X[2 * nt + g]};
  double u[3], g9[9];
  a.DD[g] = swe_finalize(acc, x, u, g9);
  if (do_velocity)
    for (int k = 0; k < 3; ++k) a.U[k * nt + g] = u[k];
}

__global__ void swe_pack_state_kernel(const SweStageArgs a) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.t0 + li;
  const double x[3] = {a.X[g], a.X[a.nt + g], a.X[2 * a.nt + g]};
  swe_pack(a, g, x, a.Z[g], a.S[g], a.T[g]);
}

struct SweIo {  // device pointers of one side of a set/get (nullptr = absent)
  Vec3View xyz, vel;
  double *vort, *div, *third, *mass, *depth, *surf, *bottom, *ddot, *laps;
};

__global__ void swe_import_kernel(int nv, int nf, SweIo p, SweIo f, SweStageArgs a) {
  const long nt = (long)nv + nf;
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= nt) return;
  const bool face = g >= nv;
  const long i = face ? g - nv : g;
  const SweIo& io = face ? f : p;
  for (int k = 0; k < 3; ++k) {
    a.X[k * nt + g] = io.xyz(i, k);
    a.U[k * nt + g] = io.vel.p ? io.vel(i, k) : 0.0;
  }
  a.Z[g] = io.vort[i];
  a.S[g] = io.div ? io.div[i] : 0.0;
  a.T[g] = io.third[i];  // depth (vertices) / area (faces)
  a.MASS[g] = (face && io.mass) ? io.mass[i] : 0.0;
  a.DEPTH[g] = face ? (io.depth ? io.depth[i] : 0.0) : io.third[i];
  a.SURF[g] = io.surf ? io.surf[i] : 0.0;
  a.BOT[g] = io.bottom ? io.bottom[i] : 0.0;
  a.DD[g] = io.ddot ? io.ddot[i] : 0.0;
  a.LAPS[g] = io.laps ? io.laps[i] : 0.0;
}

__global__ void swe_export_kernel(int nv, int nf, SweIo p, SweIo f, SweStageArgs a) {
  const long nt = (long)nv + nf;
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= nt) return;
  const bool face = g >= nv;
  const long i = face ? g - nv : g;
  const SweIo& io = face ? f : p;
  for (int k = 0; k < 3; ++k) {
    if (io.xyz.p) io.xyz(i, k) = a.X[k * nt + g];
    if (io.vel.p) io.vel(i, k) = a.U[k * nt + g];
  }
  if (io.vort) io.vort[i] = a.Z[g];
  if (io.div) io.div[i] = a.S[g];
  if (io.third) io.third[i] = a.T[g];
  if (face && io.mass) io.mass[i] = a.MASS[g];
  if (io.depth) io.depth[i] = face ? a.DEPTH[g] : a.T[g];
  if (io.surf) io.surf[i] = a.SURF[g];
  if (io.bottom) io.bottom[i] = a.BOT[g];
  if (io.ddot) io.ddot[i] = a.DD[g];
  if (io.laps) io.laps[i] = a.LAPS[g];
}

__global__ void swe_self_idx_kernel(int nv, int nf, const unsigned char* mask, const int* leaf_idx, int skip_self,
                                    int* self_idx) {
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= (long)nv + nf) return;
  int v = -1;
  if (g >= nv && skip_self && !mask[g - nv]) v = leaf_idx[g - nv];
  self_idx[g] = v;
}

static SweStageArgs swe_args(SweState* s, int stage, double dt, double Omega, double g, double* packed_next, bool with_plan) {
  SweStageArgs a;
  a.pv = with_plan ? part_view(s->plan, s->partials) : PartView{nullptr, 0, 0, 0, 1, 1};
  a.t0 = s->t0, a.n_local = s->t1 - s->t0, a.nv = s->nv, a.stage = stage, a.nt = s->nt;
  a.dt = dt, a.Omega = Omega, a.g = g;
  a.X = s->X, a.U = s->U, a.Xw = s->Xw, a.K1x = s->K1x, a.Z = s->Z, a.S = s->S, a.T = s->T, a.Zw = s->Zw, a.Sw = s->Sw,
  a.Tw = s->Tw, a.K1z = s->K1z, a.K1s = s->K1s, a.K1t = s->K1t, a.DD = s->DD, a.LAPS = s->LAPS, a.SURF = s->SURF,
  a.BOT = s->BOT, a.DEPTH = s->DEPTH, a.MASS = s->MASS;
  a.mask = s->mask, a.leaf_idx = s->leaf_idx, a.packed_next = packed_next;
  return a;
}

static int swe_alloc(SweState* s, lpmx_handle_t h, int nv, int nf, double eps) {
  if (!h || nv < 0 || nf < 0) return LPMX_ERR_INVALID;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  s->h = h, s->nv = nv, s->nf = nf, s->nt = nv + nf, s->eps = eps;
  const size_t nt = s->nt;
  const size_t pk = kSweRec * (size_t)(round_up_chunk(nf) + kChunk);
  const size_t dbl = 12 * nt + 15 * nt + 2 * pk + 32;
  const size_t bytes = dbl * sizeof(double) + sizeof(int) * (size_t)(nf + 1 + nt + 1) + (size_t)nf + 512;
  LPMX_TRY(slab_alloc(h, &s->slab, bytes));
  LPMX_CUDA(h, cudaMemsetAsync(s->slab, 0, bytes, h->stream));
  double* p = (double*)s->slab;
  s->X = p, p += 3 * nt;
  s->U = p, p += 3 * nt;
  s->Xw = p, p += 3 * nt;
  s->K1x = p, p += 3 * nt;
  double** one[] = {&s->Z, &s->S, &s->T, &s->Zw, &s->Sw, &s->Tw, &s->K1z, &s->K1s, &s->K1t, &s->DD, &s->LAPS,
                    &s->SURF, &s->BOT, &s->DEPTH, &s->MASS};
  for (double** q : one) *q = p, p += nt;
  p = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);  // TMA source alignment
  s->packed[0] = p, p += pk;
  s->packed[1] = p, p += pk;
  int* ip = (int*)p;
  s->leaf_idx = ip, ip += nf + 1;
  s->self_idx = ip, ip += nt + 1;
  s->mask = (unsigned char*)ip;
  s->t0 = (int)(((long)h->rank * s->nt) / h->world);
  s->t1 = (int)(((long)(h->rank + 1) * s->nt) / h->world);
  return LPMX_OK;
}

static size_t vbytes(int layout, long ld, int n) {
  return (layout == LPMX_LAYOUT_LEFT ? (size_t)(2 * ld + n) : (size_t)3 * n) * sizeof(double);
}

static int swe_exchange_rows(SweState* s, double* base, int n_rows) {
  if (s->h->world == 1) return LPMX_OK;
  for (int r = 0; r < n_rows; ++r) LPMX_TRY(comm_allgatherv(s->h, base + (long)r * s->nt, s->tgt_off.data()));
  return LPMX_OK;
}

}  // namespace lpmx

struct lpmx_swe_solver_s {
  SweState st;
};

static int swe_set_state(lpmx_swe_solver_s* sv, const lpmx_swe_passive_t* P, const lpmx_swe_active_t* A, int layout,
                         long pld, long ald) {
  SweState* s = &sv->st;
  lpmx_handle_t h = s->h;
  if (!P || !A) return set_error(h, LPMX_ERR_INVALID, "null state struct");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if ((s->nv > 0 && (!P->xyz || !P->vort || !P->depth)) || (s->nf > 0 && (!A->xyz || !A->vort || !A->area || !A->mask)))
    return set_error(h, LPMX_ERR_INVALID, "null state array (xyz, vort, depth/area and mask are required)");
  if (layout == LPMX_LAYOUT_LEFT && (pld < s->nv || ald < s->nf))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  SweIo p{}, f{};
  const void* d = nullptr;
#define IN_VEC(io, field, name, user, ld, n)                              \
  LPMX_TRY(stage_in(h, name, user, vbytes(layout, ld, n), &d));           \
  io.field = make_view((const double*)d, layout, ld);
#define IN_SCL(io, field, name, user, n)                                   \
  LPMX_TRY(stage_in(h, name, user, sizeof(double) * (size_t)(n), &d));    \
  io.field = (double*)d;
  IN_VEC(p, xyz, "sw_pxyz", P->xyz, pld, s->nv)
  IN_VEC(p, vel, "sw_pvel", P->vel, pld, s->nv)
  IN_SCL(p, vort, "sw_pz", P->vort, s->nv)
  IN_SCL(p, div, "sw_ps", P->div, s->nv)
  IN_SCL(p, third, "sw_ph", P->depth, s->nv)
  IN_SCL(p, surf, "sw_psurf", P->surf, s->nv)
  IN_SCL(p, bottom, "sw_pbot", P->bottom, s->nv)
  IN_SCL(p, ddot, "sw_pdd", P->ddot, s->nv)
  IN_SCL(p, laps, "sw_plaps", P->laps, s->nv)
  IN_VEC(f, xyz, "sw_axyz", A->xyz, ald, s->nf)
  IN_VEC(f, vel, "sw_avel", A->vel, ald, s->nf)
  IN_SCL(f, vort, "sw_az", A->vort, s->nf)
  IN_SCL(f, div, "sw_as", A->div, s->nf)
  IN_SCL(f, third, "sw_aarea", A->area, s->nf)
  IN_SCL(f, mass, "sw_amass", A->mass, s->nf)
  IN_SCL(f, depth, "sw_adepth", A->depth, s->nf)
  IN_SCL(f, surf, "sw_asurf", A->surf, s->nf)
  IN_SCL(f, bottom, "sw_abot", A->bottom, s->nf)
  IN_SCL(f, ddot, "sw_add", A->ddot, s->nf)
  IN_SCL(f, laps, "sw_alaps", A->laps, s->nf)
#undef IN_VEC
#undef IN_SCL
  LPMX_TRY(stage_in(h, "sw_mask", A->mask, (size_t)s->nf, &d));
  if (s->nf > 0) LPMX_CUDA(h, cudaMemcpyAsync(s->mask, d, (size_t)s->nf, cudaMemcpyDeviceToDevice, h->stream));
  const int threads = 256, blocks = (s->nt + threads - 1) / threads;
  if (s->nt > 0) {
    swe_import_kernel<<<blocks, threads, 0, h->stream>>>(s->nv, s->nf, p, f, swe_args(s, 0, 0, 0, 0, nullptr, false));
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  LPMX_TRY(scan_leaves(h, s->mask, s->nf, s->leaf_idx, &s->n_leaf));
  // SphereFaceSums: collocated = FloatingPoint<Real>::zero(eps) (lpm_swe_kernels.hpp:913)
  const int skip = std::fabs(s->eps) < DBL_EPSILON;
  if (s->nt > 0) {
    swe_self_idx_kernel<<<blocks, threads, 0, h->stream>>>(s->nv, s->nf, s->mask, s->leaf_idx, skip, s->self_idx);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  const size_t pk_bytes = sizeof(double) * kSweRec * (size_t)(round_up_chunk(s->nf) + kChunk);
  LPMX_CUDA(h, cudaMemsetAsync(s->packed[0], 0, pk_bytes, h->stream));
  LPMX_CUDA(h, cudaMemsetAsync(s->packed[1], 0, pk_bytes, h->stream));
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
    s->packed_off[r] = kSweRec * l;
  }
  s->t0 = (int)s->tgt_off[h->rank];
  s->t1 = (int)s->tgt_off[h->rank + 1];
  LPMX_TRY(make_plan(h, kSwe, s->t1 - s->t0, s->n_leaf, &s->plan));
  void* part = nullptr;
  LPMX_TRY(dev_buffer(h, "swe_partials", plan_partials_bytes(s->plan) + 256, &part));
  s->partials = (double*)part;
  s->has_state = true;
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));  // staging buffers may be reused by the next call
  return LPMX_OK;
}

static int swe_get_state(lpmx_swe_solver_s* sv, const lpmx_swe_passive_t* P, const lpmx_swe_active_t* A, int layout,
                         long pld, long ald) {
  SweState* s = &sv->st;
  lpmx_handle_t h = s->h;
  if (!s->has_state) return set_error(h, LPMX_ERR_STATE, "get_state before set_state");
  if (!P || !A) return set_error(h, LPMX_ERR_INVALID, "null state struct");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  if (h->world > 1) {  // every rank returns the full state
    LPMX_TRY(swe_exchange_rows(s, s->X, 3));
    LPMX_TRY(swe_exchange_rows(s, s->U, 3));
    for (double* row : {s->Z, s->S, s->T, s->DD, s->LAPS, s->SURF, s->BOT, s->DEPTH}) LPMX_TRY(swe_exchange_rows(s, row, 1));
  }
  SweIo p{}, f{};
  struct Out {
    void* user;
    void* dev;
    size_t bytes;
  };
  std::vector<Out> outs;
  void* d = nullptr;
#define OUT_VEC(io, field, name, user, ld, n)                                     \
  if (user) {                                                                     \
    LPMX_TRY(stage_out_begin(h, name, user, vbytes(layout, ld, n), &d));          \
    io.field = make_view((const double*)d, layout, ld);                           \
    outs.push_back({user, d, vbytes(layout, ld, n)});                             \
  }
#define OUT_SCL(io, field, name, user, n)                                         \
  if (user) {                                                                     \
    LPMX_TRY(stage_out_begin(h, name, user, sizeof(double) * (size_t)(n), &d));   \
    io.field = (double*)d;                                                        \
    outs.push_back({user, d, sizeof(double) * (size_t)(n)});                      \
  }
  OUT_VEC(p, xyz, "so_pxyz", P->xyz, pld, s->nv)
  OUT_VEC(p, vel, "so_pvel", P->vel, pld, s->nv)
  OUT_SCL(p, vort, "so_pz", P->vort, s->nv)
  OUT_SCL(p, div, "so_ps", P->div, s->nv)
  OUT_SCL(p, third, "so_ph", P->depth, s->nv)
  OUT_SCL(p, surf, "so_psurf", P->surf, s->nv)
  OUT_SCL(p, bottom, "so_pbot", P->bottom, s->nv)
  OUT_SCL(p, ddot, "so_pdd", P->ddot, s->nv)
  OUT_SCL(p, laps, "so_plaps", P->laps, s->nv)
  OUT_VEC(f, xyz, "so_axyz", A->xyz, ald, s->nf)
  OUT_VEC(f, vel, "so_avel", A->vel, ald, s->nf)
  OUT_SCL(f, vort, "so_az", A->vort, s->nf)
  OUT_SCL(f, div, "so_as", A->div, s->nf)
  OUT_SCL(f, third, "so_aarea", A->area, s->nf)
  OUT_SCL(f, mass, "so_amass", A->mass, s->nf)
  OUT_SCL(f, depth, "so_adepth", A->depth, s->nf)
  OUT_SCL(f, surf, "so_asurf", A->surf, s->nf)
  OUT_SCL(f, bottom, "so_abot", A->bottom, s->nf)
  OUT_SCL(f, ddot, "so_add", A->ddot, s->nf)
  OUT_SCL(f, laps, "so_alaps", A->laps, s->nf)
#undef OUT_VEC
#undef OUT_SCL
  if (s->nt > 0) {
    const int threads = 256, blocks = (s->nt + threads - 1) / threads;
    swe_export_kernel<<<blocks, threads, 0, h->stream>>>(s->nv, s->nf, p, f, swe_args(s, 0, 0, 0, 0, nullptr, false));
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

static int swe_launch_stage(SweState* s, const SweStageArgs& a) {
  lpmx_handle_t h = s->h;
  if (a.n_local <= 0) return LPMX_OK;
  const int threads = 128, blocks = (a.n_local + threads - 1) / threads;
  swe_rk2_stage_kernel<<<blocks, threads, 0, h->stream>>>(a);
  ++h->launches;
  return check_cuda(h, cudaGetLastError(), "swe_rk2_stage_kernel launch");
}

static int swe_pair_sum(SweState* s, double* tgt_base) {
  const double kappa = 1.0 + s->eps * s->eps;
  // the partials live in a handle-wide grow-only buffer: another SWE solver on this handle with a larger plan may have
  // reallocated it since set_state, so the pointer is fetched again for every evaluation (as ensure_partials does for BVE/IC2D);
  // the stage kernel that follows reads it through s->partials
  void* part = nullptr;
  LPMX_TRY(dev_buffer(s->h, "swe_partials", plan_partials_bytes(s->plan) + 256, &part));
  s->partials = (double*)part;
  return launch_pair_sum(s->h, s->plan, s->local_view(tgt_base), s->self_idx + s->t0, s->packed[s->cur], kappa, s->partials);
}

static int swe_exchange_packed(SweState* s, double* packed) {
  if (s->h->world == 1) return LPMX_OK;
  return comm_allgatherv(s->h, packed, s->packed_off.data());
}

// hand the (work or new) coordinates and surface heights to the Laplacian provider
static int swe_call_laplacian(SweState* s, lpmx_swe_laplacian_fn fn, void* user, int stage, double* xbase) {
  if (!fn) return LPMX_OK;
  lpmx_handle_t h = s->h;
  if (h->world > 1) {  // the provider sees every particle
    LPMX_TRY(swe_exchange_rows(s, xbase, 3));
    LPMX_TRY(swe_exchange_rows(s, s->SURF, 1));
  }
  const int rc = fn(user, stage, (void*)h->stream, s->nv, xbase, s->SURF, s->LAPS, s->nf, xbase + s->nv, s->SURF + s->nv,
                    s->mask, s->LAPS + s->nv, (long)s->nt);
  if (rc != 0) return set_error(h, LPMX_ERR_INVALID, "surface-Laplacian callback failed (%d) at stage %d", rc, stage);
  return LPMX_OK;
}

extern "C" {

int lpmx_swe_solver_create(lpmx_handle_t h, int n_passive, int n_active, double eps, lpmx_swe_solver_t* out) {
  if (!h || !out) return LPMX_ERR_INVALID;
  lpmx_swe_solver_s* s = new (std::nothrow) lpmx_swe_solver_s;
  if (!s) return LPMX_ERR_NOMEM;
  const int rc = swe_alloc(&s->st, h, n_passive, n_active, eps);
  if (rc != LPMX_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return LPMX_OK;
}

int lpmx_swe_solver_destroy(lpmx_swe_solver_t s) {
  if (!s) return LPMX_OK;
  if (s->st.h && s->st.h->cached_swe == s) s->st.h->cached_swe = nullptr;
  if (s->st.slab) {
    cudaSetDevice(s->st.h->device);
    cudaStreamSynchronize(s->st.h->stream);
    slab_free(s->st.h, s->st.slab);
  }
  delete s;
  return LPMX_OK;
}

int lpmx_swe_solver_set_state(lpmx_swe_solver_t s, const lpmx_swe_passive_t* passive, const lpmx_swe_active_t* active,
                              int layout, long passive_ld, long active_ld) {
  if (!s) return LPMX_ERR_INVALID;
  return swe_set_state(s, passive, active, layout, passive_ld, active_ld);
}

int lpmx_swe_solver_get_state(lpmx_swe_solver_t s, const lpmx_swe_passive_t* passive, const lpmx_swe_active_t* active,
                              int layout, long passive_ld, long active_ld) {
  if (!s) return LPMX_ERR_INVALID;
  return swe_get_state(s, passive, active, layout, passive_ld, active_ld);
}

int lpmx_swe_solver_set_laplacian(lpmx_swe_solver_t sv, const double* passive_laps, const double* active_laps) {
  if (!sv) return LPMX_ERR_INVALID;
  SweState* s = &sv->st;
  lpmx_handle_t h = s->h;
  if (!s->has_state) return set_error(h, LPMX_ERR_STATE, "set_laplacian before set_state");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  if (passive_laps && s->nv > 0)
    LPMX_CUDA(h, cudaMemcpyAsync(s->LAPS, passive_laps, sizeof(double) * s->nv, cudaMemcpyDefault, h->stream));
  if (active_laps && s->nf > 0)
    LPMX_CUDA(h, cudaMemcpyAsync(s->LAPS + s->nv, active_laps, sizeof(double) * s->nf, cudaMemcpyDefault, h->stream));
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

int lpmx_swe_solver_init_direct_sums(lpmx_swe_solver_t sv, int do_velocity) {
  if (!sv) return LPMX_ERR_INVALID;
  SweState* s = &sv->st;
  lpmx_handle_t h = s->h;
  if (!s->has_state) return set_error(h, LPMX_ERR_STATE, "init_direct_sums before set_state");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  SweStageArgs a = swe_args(s, 0, 0, 0, 0, s->packed[s->cur], true);
  if (a.n_local > 0) {
    const int threads = 256, blocks = (a.n_local + threads - 1) / threads;
    swe_pack_state_kernel<<<blocks, threads, 0, h->stream>>>(a);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  LPMX_TRY(swe_exchange_packed(s, s->packed[s->cur]));
  LPMX_TRY(swe_pair_sum(s, s->X));
  a = swe_args(s, 0, 0, 0, 0, s->packed[s->cur], true);  // swe_pair_sum re-fetched the partials buffer
  if (a.n_local > 0) {
    const int threads = 128, blocks = (a.n_local + threads - 1) / threads;
    swe_init_sums_kernel<<<blocks, threads, 0, h->stream>>>(a, do_velocity);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  return LPMX_OK;
}

int lpmx_swe_solver_advance(lpmx_swe_solver_t sv, double dt, double Omega, double g, lpmx_swe_laplacian_fn laplacian,
                            void* user, int n_steps) {
  if (!sv) return LPMX_ERR_INVALID;
  SweState* s = &sv->st;
  lpmx_handle_t h = s->h;
  if (!s->has_state) return set_error(h, LPMX_ERR_STATE, "advance before set_state");
  if (n_steps < 0) return set_error(h, LPMX_ERR_INVALID, "negative step count");
  if (n_steps == 0 || s->nt == 0) return LPMX_OK;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  for (int step = 0; step < n_steps; ++step) {
    // stage 0: stage-1 tendencies, predictor state + surfaces + source records
    LPMX_TRY(swe_launch_stage(s, swe_args(s, 0, dt, Omega, g, s->packed[s->cur ^ 1], false)));
    s->cur ^= 1;
    LPMX_TRY(swe_exchange_packed(s, s->packed[s->cur]));
    LPMX_TRY(swe_call_laplacian(s, laplacian, user, 1, s->Xw));
    LPMX_TRY(swe_pair_sum(s, s->Xw));
    // stage 1: (u, ddot) at the predictor, stage-2 tendencies, Heun combine, new surfaces + source records
    LPMX_TRY(swe_launch_stage(s, swe_args(s, 1, dt, Omega, g, s->packed[s->cur ^ 1], true)));
    s->cur ^= 1;
    LPMX_TRY(swe_exchange_packed(s, s->packed[s->cur]));
    LPMX_TRY(swe_pair_sum(s, s->X));
    // stage 2: (u, ddot) of the new state, then the new state's Laplacian
    LPMX_TRY(swe_launch_stage(s, swe_args(s, 2, dt, Omega, g, nullptr, true)));
    LPMX_TRY(swe_call_laplacian(s, laplacian, user, 2, s->X));
  }
  return LPMX_OK;
}

int lpmx_swe_rk2_step(lpmx_handle_t h, double dt, double Omega, double g, double eps, int n_passive,
                      const lpmx_swe_passive_t* passive, int n_active, const lpmx_swe_active_t* active, int layout,
                      long passive_ld, long active_ld, lpmx_swe_laplacian_fn laplacian, void* user, int n_steps) {
  if (!h) return LPMX_ERR_INVALID;
  if (!passive || !active) return set_error(h, LPMX_ERR_INVALID, "null state struct");
  if ((n_passive > 0 && (!passive->vel || !passive->ddot)) || (n_active > 0 && (!active->vel || !active->ddot)))
    return set_error(h, LPMX_ERR_INVALID, "velocity and double-dot arrays are required (SWE::init_direct_sums output)");
  lpmx_swe_solver_t s = h->cached_swe;
  if (!s || s->st.nv != n_passive || s->st.nf != n_active || s->st.eps != eps) {
    if (s) lpmx_swe_solver_destroy(s);
    h->cached_swe = nullptr;
    LPMX_TRY(lpmx_swe_solver_create(h, n_passive, n_active, eps, &s));
    h->cached_swe = s;
  }
  LPMX_TRY(swe_set_state(s, passive, active, layout, passive_ld, active_ld));
  LPMX_TRY(lpmx_swe_solver_advance(s, dt, Omega, g, laplacian, user, n_steps));
  return swe_get_state(s, passive, active, layout, passive_ld, active_ld);
}

}  // extern "C"
