// lpmx_steppers.cu -- device-resident time steppers: BVERK4 and Incompressible2DRK2 on the sphere.
//
// Reference (as coded, quirks included -- SURVEY.md 8(a)):
//   BVERK4::advance_timestep            src/lpm_bve_rk4_impl.hpp:63-167  (38 launches per step)
//   BVERK4Update                        src/lpm_bve_rk4_impl.hpp:12-53
//   BVEVorticityTendency                src/lpm_bve_sphere_kernels.hpp:399-413
//   Incompressible2DRK2::advance_timestep_impl   src/lpm_incompressible2d_rk2_impl.hpp:75-172
//   Incompressible2DTendencies          src/lpm_incompressible2d_kernels.hpp:257-278
//
// Here a step is (pair-sum kernel + one fused O(N) stage kernel) per velocity evaluation: the stage
// kernel adds the partial moments, applies u = x cross M, forms the stage increments, the next
// stage's input state AND its packed source records (ping-pong buffers), so the reference's
// KokkosBlas scal/update calls and tendency/update functors never run as separate passes.
// Targets (vertices then faces) are stored structure-of-arrays.  With world > 1 each rank owns a contiguous range of the
// vertices and a contiguous range of the faces (1/world of each) and evaluates them as two index lists per stage: list A = its
// leaf faces (the sources it contributes), list B = its vertices and divided faces (never sources).  A goes first; the packed
// records its stage kernel writes are exchanged (one kernel over NVLink peer memory on the copy stream when the peer path is on,
// else the NCCL all-gather in line) WHILE the pair sum of list B runs, so the exchange is off the critical path.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "lpmx_finalize.cuh"
#include "lpmx_internal.h"

using namespace lpmx;

namespace lpmx {

// state common to both steppers: SoA arrays over the concatenated target list
struct SolverState {
  lpmx_handle_t h = nullptr;
  int nv = 0, nf = 0, nt = 0, n_leaf = 0;
  // Target lists.  world == 1: one "list" B = everything (gid[1] null: the identity), list A empty.  world > 1: the leaf faces
  // (in index order) and the non-sources (vertices, then divided faces in index order) are each cut into `world` equal
  // slices; rank r owns slice r of both: list A = gid[0] = its leaf faces, list B = gid[1] = its non-sources (global indices
  // into the concatenated SoA arrays).  `perm` (device, nt ints) is the concatenation [A_0, B_0, A_1, B_1, ...] of all
  // ranks' lists, p_off[r] the start of rank r's segment in it: the row gather of get_state runs in that order.
  bool split = false;
  int* gid[2] = {nullptr, nullptr};
  int n_part[2] = {0, 0};
  int* perm = nullptr;  // nt ints in the slab
  std::vector<long> p_off;
  std::vector<std::pair<int, int>> own_v_runs, own_f_runs;  // this rank's rows as [first, last) runs (sharded host I/O)
  // the lists depend on the mask only: a set_state with the same (host) mask bytes, rank and world keeps them (the in-place
  // steppers call set_state every step)
  std::vector<unsigned char> list_mask;
  int list_rank = -1, list_world = -1;
  bool has_state = false;
  void* slab = nullptr;
  double *X = nullptr, *U = nullptr, *Xw = nullptr, *Z = nullptr, *Zw = nullptr, *Psi = nullptr;
  double* K[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};  // K[s][0]=x incr (3*nt), [1]=zeta incr
  double* area = nullptr;
  unsigned char* mask = nullptr;
  int* leaf_idx = nullptr;
  int* self_idx = nullptr;
  double* packed[2] = {nullptr, nullptr};
  int cur = 0;
  int n_src_pad = 0;
  double* partials[2] = {nullptr, nullptr};  // per list
  // Merged evaluation (sharded solvers on small meshes): the two lists summed by ONE bank sequence over gid[0][0 .. n_A + n_B)
  // -- the lists are adjacent in `perm` -- into one accumulator array; the stage kernels of list A and list B then read their
  // parts of it through pv_m (set for the duration of the stage launches of such an evaluation).
  double* partials_m = nullptr;
  bool merged_now = false;
  PartView pv_m[2];
  std::vector<long> packed_off;    // world+1 offsets into packed (in doubles): rank r's leaves are [r L / W, (r+1) L / W)
  int n_local() const { return n_part[0] + n_part[1]; }
  Vec3View view(double* base) const {
    Vec3View v;
    v.p = base;
    v.si = 1;
    v.sk = nt;
    return v;
  }
};

static int solver_alloc(SolverState* s, lpmx_handle_t h, int nv, int nf, int n_k) {
  if (!h || nv < 0 || nf < 0) return LPMX_ERR_INVALID;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  s->h = h;
  s->nv = nv;
  s->nf = nf;
  s->nt = nv + nf;
  const long nt = s->nt;
  s->n_src_pad = round_up_chunk(nf);  // upper bound: every face a leaf
  // one slab: X U Xw (3nt each) Z Zw Psi (nt each) K (n_k * 4nt) area packed[2] | ints | mask
  size_t dbl = 9 * nt + 3 * nt + (size_t)n_k * 4 * nt + nf + 2 * kBveRec * (size_t)(s->n_src_pad + kChunk);
  size_t bytes = dbl * sizeof(double) + sizeof(int) * (size_t)(nf + 1 + nt + 1 + nt + 1) + (size_t)nf + 64 + 256;
  LPMX_TRY(slab_alloc(h, &s->slab, bytes));
  LPMX_CUDA(h, cudaMemsetAsync(s->slab, 0, bytes, h->stream));  // Kokkos views start at zero
  double* p = (double*)s->slab;
  s->X = p, p += 3 * nt;
  s->U = p, p += 3 * nt;
  s->Xw = p, p += 3 * nt;
  s->Z = p, p += nt;
  s->Zw = p, p += nt;
  s->Psi = p, p += nt;
  for (int k = 0; k < n_k; ++k) {
    s->K[k][0] = p, p += 3 * nt;
    s->K[k][1] = p, p += nt;
  }
  s->area = p, p += nf;
  p = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);  // TMA source alignment
  s->packed[0] = p, p += kBveRec * (size_t)(s->n_src_pad + kChunk);
  s->packed[1] = p, p += kBveRec * (size_t)(s->n_src_pad + kChunk);
  int* ip = (int*)p;
  s->leaf_idx = ip, ip += nf + 1;
  s->self_idx = ip, ip += nt + 1;
  s->perm = ip, ip += nt + 1;
  s->mask = (unsigned char*)ip;
  return LPMX_OK;
}

static void solver_free(SolverState* s) {
  if (s->slab) {
    cudaSetDevice(s->h->device);
    cudaStreamSynchronize(s->h->stream);
    slab_free(s->h, s->slab);
    s->slab = nullptr;
  }
}

// user views -> SoA state; also self index of every target
__global__ void import_state_kernel(int nv, int nf, Vec3View vx, const double* vz, Vec3View vu, Vec3View fx,
                                    const double* fz, Vec3View fu, double* X, double* Z, double* U) {
  const long nt = (long)nv + nf;
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= nt) return;
  const bool vert = g < nv;
  const long i = vert ? g : g - nv;
  const Vec3View& x = vert ? vx : fx;
  const Vec3View& u = vert ? vu : fu;
  const double* z = vert ? vz : fz;
  for (int k = 0; k < 3; ++k) {
    X[k * nt + g] = x(i, k);
    U[k * nt + g] = u.p ? u(i, k) : 0.0;
  }
  Z[g] = z[i];
}

__global__ void export_state_kernel(int nv, int nf, Vec3View vx, double* vz, Vec3View vu, double* vpsi, Vec3View fx,
                                    double* fz, Vec3View fu, double* fpsi, const double* X, const double* Z,
                                    const double* U, const double* Psi) {
  const long nt = (long)nv + nf;
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= nt) return;
  const bool vert = g < nv;
  const long i = vert ? g : g - nv;
  const Vec3View& x = vert ? vx : fx;
  const Vec3View& u = vert ? vu : fu;
  double* z = vert ? vz : fz;
  double* psi = vert ? vpsi : fpsi;
  for (int k = 0; k < 3; ++k) {
    if (x.p) x(i, k) = X[k * nt + g];
    if (u.p) u(i, k) = U[k * nt + g];
  }
  if (z) z[i] = Z[g];
  if (psi) psi[i] = Psi[g];
}

__global__ void self_idx_kernel(int nv, int nf, const unsigned char* mask, const int* leaf_idx, int skip_self,
                                int* self_idx) {
  const long g = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (g >= (long)nv + nf) return;
  int v = -1;
  if (g >= nv && skip_self) {
    const long f = g - nv;
    if (!mask[f]) v = leaf_idx[f];
  }
  self_idx[g] = v;
}

// write the packed record of face-target g (if it is a leaf)
__device__ __forceinline__ void pack_target(long g, int nv, const unsigned char* mask, const int* leaf_idx,
                                            const double* area, const double* x, double zeta, double* packed) {
  if (g < nv || !packed) return;
  const long f = g - nv;
  if (mask[f]) return;
  write_bve_record(packed + kBveRec * (size_t)leaf_idx[f], x, gamma_of(zeta, area[f]));
}

// pack the current state (X, Z) [or work state] of this rank's targets
__global__ void pack_state_kernel(const int* gid, int n_local, int nv, long nt, const double* X, const double* Z,
                                  const unsigned char* mask, const int* leaf_idx, const double* area, double* packed) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= n_local) return;
  const long g = gid ? gid[li] : li;
  const double x[3] = {X[g], X[nt + g], X[2 * nt + g]};
  pack_target(g, nv, mask, leaf_idx, area, x, Z[g], packed);
}

struct StageArgs {
  PartView pv;
  const int* gid;  // index list of this launch's targets (null: the identity)
  int n_local, nv;
  long nt;
  int stage;  // which evaluation just finished (1-based); 0 = prologue from U
  int more;   // another step follows (fuse its stage 1)
  double dt, Omega;
  double *X, *U, *Xw, *Z, *Zw, *Psi;
  double *K1x, *K1z, *K2x, *K2z, *K3x, *K3z;
  const double* area;
  const unsigned char* mask;
  const int* leaf_idx;
  double* packed_next;
};

// ---- BVE RK4 -------------------------------------------------------------------------------
// stage 0 (prologue) and the tail of stage 4 both do "stage 1 from the current velocity":
//   x1 = dt*u, zeta1 = -2 Omega u_z dt, work = state + 0.5 * increment   (:85-98)
__device__ __forceinline__ void bve_stage1(const StageArgs& a, long g, const double* u, double* xw, double* zw) {
  const long nt = a.nt;
  const double z1 = -2.0 * a.Omega * u[2] * a.dt;
  a.K1z[g] = z1;
  for (int k = 0; k < 3; ++k) {
    const double x1 = a.dt * u[k];
    a.K1x[k * nt + g] = x1;
    xw[k] = a.X[k * nt + g] + 0.5 * x1;
    a.Xw[k * nt + g] = xw[k];
  }
  *zw = a.Z[g] + 0.5 * z1;
  a.Zw[g] = *zw;
}

__global__ void bve_rk4_stage_kernel(const StageArgs a) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.gid ? a.gid[li] : li;
  const long nt = a.nt;
  double u[3], xw[3], zw;
  if (a.stage == 0) {
    for (int k = 0; k < 3; ++k) u[k] = a.U[k * nt + g];
    bve_stage1(a, g, u, xw, &zw);
    pack_target(g, a.nv, a.mask, a.leaf_idx, a.area, xw, zw, a.packed_next);
    return;
  }
  double M[3];
  reduce_slots<3>(a.pv, li, M);
  const double x[3] = {a.Xw[g], a.Xw[nt + g], a.Xw[2 * nt + g]};
  cross3(u, x, M);
  const double zk = -2.0 * a.Omega * u[2] * a.dt;  // BVEVorticityTendency
  if (a.stage == 1 || a.stage == 2) {
    double* Kx = a.stage == 1 ? a.K2x : a.K3x;
    double* Kz = a.stage == 1 ? a.K2z : a.K3z;
    const double c = a.stage == 1 ? 0.5 : 1.0;  // stage-3 input uses 0.5*x2, stage-4 input 1.0*x3 (:114-137)
    Kz[g] = zk;
    for (int k = 0; k < 3; ++k) {
      const double xk = a.dt * u[k];
      Kx[k * nt + g] = xk;
      xw[k] = a.X[k * nt + g] + c * xk;
      a.Xw[k * nt + g] = xw[k];
    }
    zw = a.Z[g] + c * zk;
    a.Zw[g] = zw;
  } else if (a.stage == 3) {
    // BVERK4Update (:12-53).  Faces get zeta4 in the zeta3 slot (:155-157, reference quirk A-i).
    const double sixth = 1.0 / 6.0, third = 1.0 / 3.0;
    for (int k = 0; k < 3; ++k) {
      const double x4 = a.dt * u[k];
      const double xn =
          a.X[k * nt + g] + (sixth * (a.K1x[k * nt + g] + x4) + third * (a.K2x[k * nt + g] + a.K3x[k * nt + g]));
      a.X[k * nt + g] = xn;
      a.Xw[k * nt + g] = xn;
      xw[k] = xn;
    }
    const double z3 = (g >= a.nv) ? zk : a.K3z[g];
    zw = a.Z[g] + (sixth * (a.K1z[g] + zk) + third * (a.K2z[g] + z3));
    a.Z[g] = zw;
    a.Zw[g] = zw;
  } else {  // stage 4: velocity of the new state (:159-164)
    for (int k = 0; k < 3; ++k) a.U[k * nt + g] = u[k];
    if (!a.more) return;
    bve_stage1(a, g, u, xw, &zw);
  }
  pack_target(g, a.nv, a.mask, a.leaf_idx, a.area, xw, zw, a.packed_next);
}

// ---- IC2D RK2 --------------------------------------------------------------------------------
// stage 1 from the current velocity (:77-110): x1 = dt*u, zeta1 = -2 Omega u_z (no dt),
// work = x + dt*u, zeta + dt*zeta1
__device__ __forceinline__ void ic2d_stage1(const StageArgs& a, long g, const double* u, double* xw, double* zw) {
  const long nt = a.nt;
  const double z1 = -(2.0 * a.Omega * u[2]);
  a.K1z[g] = z1;
  for (int k = 0; k < 3; ++k) {
    a.K1x[k * nt + g] = a.dt * u[k];
    xw[k] = a.X[k * nt + g] + a.dt * u[k];
    a.Xw[k * nt + g] = xw[k];
  }
  *zw = a.Z[g] + a.dt * z1;
  a.Zw[g] = *zw;
}

template <bool WITH_PSI>
__global__ void ic2d_rk2_stage_kernel(const StageArgs a) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= a.n_local) return;
  const long g = a.gid ? a.gid[li] : li;
  const long nt = a.nt;
  double u[3], xw[3], zw;
  if (a.stage == 0) {
    for (int k = 0; k < 3; ++k) u[k] = a.U[k * nt + g];
    ic2d_stage1(a, g, u, xw, &zw);
    pack_target(g, a.nv, a.mask, a.leaf_idx, a.area, xw, zw, a.packed_next);
    return;
  }
  constexpr int NACC = WITH_PSI ? 4 : 3;
  double M[NACC];
  reduce_slots<NACC>(a.pv, li, M);
  const double x[3] = {a.Xw[g], a.Xw[nt + g], a.Xw[2 * nt + g]};
  cross3(u, x, M);
  if (a.stage == 1) {
    // (:127-155) x2 = dt*u, zeta2 = -2 Omega u_z ; zeta += .5dt*z1 + .5dt*z2 ; x += .5*x1 + .5*x2
    const double z2 = -(2.0 * a.Omega * u[2]);
    const double hdt = 0.5 * a.dt;
    for (int k = 0; k < 3; ++k) {
      const double x2 = a.dt * u[k];
      const double xn = a.X[k * nt + g] + (0.5 * a.K1x[k * nt + g] + 0.5 * x2);
      a.X[k * nt + g] = xn;
      a.Xw[k * nt + g] = xn;
      xw[k] = xn;
      a.U[k * nt + g] = u[k];  // the reference overwrites the velocity view at this point too
    }
    zw = a.Z[g] + (hdt * a.K1z[g] + hdt * z2);
    a.Z[g] = zw;
    a.Zw[g] = zw;
  } else {  // stage 2: velocity and stream function of the new state (:157-170)
    for (int k = 0; k < 3; ++k) a.U[k * nt + g] = u[k];
    if (WITH_PSI) a.Psi[g] = M[NACC - 1];
    if (!a.more) return;
    ic2d_stage1(a, g, u, xw, &zw);
  }
  pack_target(g, a.nv, a.mask, a.leaf_idx, a.area, xw, zw, a.packed_next);
}

// psi-only finalize on the resident state (BVESphere::init_stream_fn)
__global__ void psi_out_kernel(PartView pv, const int* gid, int n_local, double* Psi) {
  const long li = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (li >= n_local) return;
  double p[1];
  reduce_slots<1>(pv, li, p);
  Psi[gid ? gid[li] : li] = p[0];
}

// the handle-wide grow-only partials buffer of one list (re-fetched per evaluation: another solver on the handle may have grown it)
static int ensure_partials(SolverState* s, int part, const SumPlan& plan) {
  const size_t need = plan_partials_bytes(plan) + 256;
  void* p = nullptr;
  LPMX_TRY(dev_buffer(s->h, part == 0 ? "solver_partials_a" : "solver_partials_b", need, &p));
  s->partials[part] = (double*)p;
  return LPMX_OK;
}

static int exchange_packed(SolverState* s, double* packed) {
  if (s->h->world == 1) return LPMX_OK;
  return comm_allgatherv(s->h, packed, s->packed_off.data());
}

// The overlapped exchange of a split evaluation.  begin(): called after the stage kernel of list A (this rank's leaf faces) has
// written its records; with the peer path the all-gather runs as ONE small kernel on the copy stream, ordered after that stage
// kernel by an event, while the compute stream goes on with the pair sum of list B; without it (NCCL) the all-gather is
// issued in line on the compute stream -- NCCL's kernels do not fit beside the persistent pair-sum CTAs, they would only run
// after them.  end(): the compute stream waits for the exchange before the next evaluation reads the records.
static int exchange_packed_begin(SolverState* s, double* packed, bool* async) {
  lpmx_handle_t h = s->h;
  *async = false;
  if (h->world == 1) return LPMX_OK;
  if (!peer_can_exchange(h, packed)) return comm_allgatherv(h, packed, s->packed_off.data());
  for (int i = 0; i < 2; ++i)
    if (!h->xchg_ev[i]) LPMX_CUDA(h, cudaEventCreateWithFlags(&h->xchg_ev[i], cudaEventDisableTiming));
  LPMX_CUDA(h, cudaEventRecord(h->xchg_ev[0], h->stream));
  LPMX_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->xchg_ev[0], 0));
  LPMX_TRY(comm_allgatherv(h, packed, s->packed_off.data(), h->copy_stream));
  LPMX_CUDA(h, cudaEventRecord(h->xchg_ev[1], h->copy_stream));
  *async = true;
  return LPMX_OK;
}
static int exchange_packed_end(SolverState* s, bool async) {
  if (async) LPMX_CUDA(s->h, cudaStreamWaitEvent(s->h->stream, s->h->xchg_ev[1], 0));
  return LPMX_OK;
}

// gather full-length SoA rows (n_rows rows of length nt) so every rank holds every target: each rank's rows are an index list,
// so the rows are gathered into list order (perm), all-gathered there (contiguous per rank) and scattered back
__global__ void perm_gather_kernel(const int* perm, long j0, long j1, const double* row, double* buf) {
  const long j = j0 + blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (j < j1) buf[j] = row[perm[j]];
}
__global__ void perm_scatter_kernel(const int* perm, long n, long skip0, long skip1, const double* buf, double* row) {
  const long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (j < n && (j < skip0 || j >= skip1)) row[perm[j]] = buf[j];
}
static int exchange_rows(SolverState* s, double* base, int n_rows) {
  lpmx_handle_t h = s->h;
  if (h->world == 1 || s->nt == 0) return LPMX_OK;
  void* bufv = nullptr;
  LPMX_TRY(dev_buffer(h, "xrows", sizeof(double) * (size_t)s->nt, &bufv));
  double* buf = (double*)bufv;
  const long j0 = s->p_off[h->rank], j1 = s->p_off[h->rank + 1];
  const int threads = 256;
  for (int r = 0; r < n_rows; ++r) {
    double* row = base + (long)r * s->nt;
    if (j1 > j0) {
      perm_gather_kernel<<<(int)((j1 - j0 + threads - 1) / threads), threads, 0, h->stream>>>(s->perm, j0, j1, row, buf);
      ++h->launches;
    }
    LPMX_TRY(comm_allgatherv(h, buf, s->p_off.data()));
    perm_scatter_kernel<<<(s->nt + threads - 1) / threads, threads, 0, h->stream>>>(s->perm, s->nt, j0, j1, buf, row);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  return LPMX_OK;
}

// Rows [r0, r1) of a Real*[3] (ncomp = 3) or Real* (ncomp = 1) array between a host array and its full-size device staging
// copy (same offsets on both sides): the sharded host I/O of lpmx_set_io_sharded.
static int copy_rows(lpmx_handle_t h, void* dev, const void* user, int layout, long ld, long r0, long r1, int ncomp,
                     bool to_device) {
  if (r1 <= r0 || !dev || !user) return LPMX_OK;
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  auto run = [&](long off, long len) -> int {
    char* d = (char*)dev + sizeof(double) * off;
    char* u = (char*)const_cast<void*>(user) + sizeof(double) * off;
    LPMX_CUDA(h, cudaMemcpyAsync(to_device ? (void*)d : (void*)u, to_device ? (const void*)u : (const void*)d,
                                 sizeof(double) * len, kind, h->stream));
    return LPMX_OK;
  };
  if (ncomp == 1) return run(r0, r1 - r0);
  if (layout == LPMX_LAYOUT_RIGHT) return run(3 * r0, 3 * (r1 - r0));
  for (int k = 0; k < 3; ++k) LPMX_TRY(run(k * ld + r0, r1 - r0));
  return LPMX_OK;
}

static int copy_runs(lpmx_handle_t h, void* dev, const void* user, int layout, long ld, const std::vector<std::pair<int, int>>& runs,
                     int ncomp, bool to_device) {
  for (const auto& r : runs) LPMX_TRY(copy_rows(h, dev, user, layout, ld, r.first, r.second, ncomp, to_device));
  return LPMX_OK;
}

// The class-balanced target lists of every rank (see SolverState): perm = [A_0, B_0, A_1, B_1, ...], p_off, and for rank `rank`
// the list sizes and its rows as runs.  `leaf(f)` tells whether face f is a leaf.  Mirrored by lpm_b200/partition.py.
template <class LeafFn>
static void build_target_lists(int nv, int nf, int n_leaf, int world, int rank, LeafFn leaf, std::vector<int>* perm,
                               std::vector<long>* p_off, int* n_a, int* n_b, std::vector<std::pair<int, int>>* v_runs,
                               std::vector<std::pair<int, int>>* f_runs) {
  std::vector<int> lf, ns;  // leaf faces; non-sources (global indices)
  lf.reserve(n_leaf), ns.reserve((size_t)nv + nf - n_leaf);
  for (int v = 0; v < nv; ++v) ns.push_back(v);
  for (int f = 0; f < nf; ++f) (leaf(f) ? lf : ns).push_back(nv + f);
  perm->clear(), perm->reserve((size_t)nv + nf);
  p_off->assign(world + 1, 0);
  std::vector<int> own;
  for (int r = 0; r < world; ++r) {
    (*p_off)[r] = (long)perm->size();
    const size_t a0 = (size_t)r * lf.size() / world, a1 = (size_t)(r + 1) * lf.size() / world;
    const size_t b0 = (size_t)r * ns.size() / world, b1 = (size_t)(r + 1) * ns.size() / world;
    perm->insert(perm->end(), lf.begin() + a0, lf.begin() + a1);
    perm->insert(perm->end(), ns.begin() + b0, ns.begin() + b1);
    if (r == rank) {
      *n_a = (int)(a1 - a0), *n_b = (int)(b1 - b0);
      own.assign(perm->end() - (*n_a + *n_b), perm->end());
    }
  }
  (*p_off)[world] = (long)perm->size();
  std::sort(own.begin(), own.end());
  v_runs->clear(), f_runs->clear();
  for (size_t i = 0; i < own.size();) {
    size_t j = i + 1;
    const bool vert = own[i] < nv;
    while (j < own.size() && own[j] == own[j - 1] + 1 && (own[j] < nv) == vert) ++j;
    if (vert)
      v_runs->push_back({own[i], own[j - 1] + 1});
    else
      f_runs->push_back({own[i] - nv, own[j - 1] + 1 - nv});
    i = j;
  }
}

static int solver_set_state(SolverState* s, const double* vx, const double* vz, const double* vu, const double* fx,
                            const double* fz, const double* fu, const double* fa, const unsigned char* fm, int layout,
                            long vld, long fld, int skip_self) {
  lpmx_handle_t h = s->h;
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  if ((s->nv > 0 && (!vx || !vz)) || (s->nf > 0 && (!fx || !fz || !fa || !fm)))
    return set_error(h, LPMX_ERR_INVALID, "null state array");
  if (layout == LPMX_LAYOUT_LEFT && (vld < s->nv || fld < s->nf))
    return set_error(h, LPMX_ERR_INVALID, "leading dimension smaller than extent");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  auto vb = [&](long ld, int n) {
    return (layout == LPMX_LAYOUT_LEFT ? (size_t)(2 * ld + n) : (size_t)3 * n) * sizeof(double);
  };
  const void *dvx, *dvz, *dvu, *dfx, *dfz, *dfu, *dfa, *dfm;
  // 1. area and mask (always in full: the leaf scan, the target lists and the conserved totals run over all faces)
  LPMX_TRY(stage_in(h, "st_fa", fa, sizeof(double) * s->nf, &dfa));
  LPMX_TRY(stage_in(h, "st_fm", fm, (size_t)s->nf, &dfm));
  if (s->nf > 0) {
    LPMX_CUDA(h, cudaMemcpyAsync(s->area, dfa, sizeof(double) * s->nf, cudaMemcpyDeviceToDevice, h->stream));
    LPMX_CUDA(h, cudaMemcpyAsync(s->mask, dfm, (size_t)s->nf, cudaMemcpyDeviceToDevice, h->stream));
  }
  LPMX_TRY(scan_leaves(h, s->mask, s->nf, s->leaf_idx, &s->n_leaf));
  s->n_src_pad = round_up_chunk(s->n_leaf);

  // 2. who owns what.  LPMX_FORCE_SPLIT=1: evaluate the two index lists on a single GPU as well (test hook: the list path of
  // the kernels and the A-then-B evaluation can then be checked on a one-GPU box; there is nothing to exchange)
  const int W = h->world;
  const char* fs = getenv("LPMX_FORCE_SPLIT");
  s->split = W > 1 || (fs && fs[0] == '1');
  s->packed_off.assign(W + 1, 0);
  for (int r = 0; r <= W; ++r) s->packed_off[r] = kBveRec * (((long)r * s->n_leaf) / W);
  if (!s->split) {
    s->own_v_runs.clear(), s->own_f_runs.clear();
    s->gid[0] = s->gid[1] = nullptr;
    s->n_part[0] = 0, s->n_part[1] = s->nt;
    s->p_off.assign(2, 0);
    s->p_off[1] = s->nt;
  } else if (!is_device_pointer(fm) && s->list_rank == h->rank && s->list_world == W && (int)s->list_mask.size() == s->nf &&
             (s->nf == 0 || memcmp(s->list_mask.data(), fm, (size_t)s->nf) == 0) && s->gid[0] != nullptr) {
    // same mask as last time: perm, p_off, the list sizes and the runs are still valid
  } else {
    std::vector<int> leaf_host(s->nf + 1, 0);
    if (s->nf > 0)
      LPMX_CUDA(h, cudaMemcpyAsync(leaf_host.data(), s->leaf_idx, sizeof(int) * s->nf, cudaMemcpyDeviceToHost, h->stream));
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
    leaf_host[s->nf] = s->n_leaf;
    std::vector<int> perm;
    build_target_lists(s->nv, s->nf, s->n_leaf, W, h->rank, [&](int f) { return leaf_host[f + 1] > leaf_host[f]; }, &perm,
                       &s->p_off, &s->n_part[0], &s->n_part[1], &s->own_v_runs, &s->own_f_runs);
    if (!perm.empty())
      LPMX_CUDA(h, cudaMemcpyAsync(s->perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, h->stream));
    LPMX_CUDA(h, cudaStreamSynchronize(h->stream));  // perm goes out of scope
    s->gid[0] = s->perm + s->p_off[h->rank];
    s->gid[1] = s->gid[0] + s->n_part[0];
    if (!is_device_pointer(fm)) {
      s->list_mask.assign(fm, fm + s->nf);
      s->list_rank = h->rank, s->list_world = W;
    } else {
      s->list_rank = -1;
    }
  }

  // 3. the state rows.  Sharded host I/O (lpmx_set_io_sharded, world > 1): only this rank's rows are read from the host arrays --
  // they are all it needs, since every rank packs the source records of its OWN leaf faces and the exchange distributes them.
  // The other rows of the device-side state are left as they are and never read.
  const bool shard_in = h->io_sharded && W > 1;
  auto in_rows = [&](const char* name, const double* user, size_t bytes, long ld, const std::vector<std::pair<int, int>>& runs,
                     int ncomp, const void** dev) -> int {
    if (!shard_in || !user || is_device_pointer(user)) return stage_in(h, name, user, bytes, dev);
    void* d = nullptr;
    LPMX_TRY(dev_buffer(h, name, bytes, &d));
    *dev = d;
    return copy_runs(h, d, user, layout, ld, runs, ncomp, true);
  };
  LPMX_TRY(in_rows("st_vx", vx, vb(vld, s->nv), vld, s->own_v_runs, 3, &dvx));
  LPMX_TRY(in_rows("st_vz", vz, sizeof(double) * s->nv, 0, s->own_v_runs, 1, &dvz));
  LPMX_TRY(in_rows("st_vu", vu, vb(vld, s->nv), vld, s->own_v_runs, 3, &dvu));
  LPMX_TRY(in_rows("st_fx", fx, vb(fld, s->nf), fld, s->own_f_runs, 3, &dfx));
  LPMX_TRY(in_rows("st_fz", fz, sizeof(double) * s->nf, 0, s->own_f_runs, 1, &dfz));
  LPMX_TRY(in_rows("st_fu", fu, vb(fld, s->nf), fld, s->own_f_runs, 3, &dfu));
  const int threads = 256;
  const int blocks = (s->nt + threads - 1) / threads;
  if (s->nt > 0) {
    import_state_kernel<<<blocks, threads, 0, h->stream>>>(
        s->nv, s->nf, make_view((const double*)dvx, layout, vld), (const double*)dvz,
        make_view((const double*)dvu, layout, vld), make_view((const double*)dfx, layout, fld), (const double*)dfz,
        make_view((const double*)dfu, layout, fld), s->X, s->Z, s->U);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
    self_idx_kernel<<<blocks, threads, 0, h->stream>>>(s->nv, s->nf, s->mask, s->leaf_idx, skip_self, s->self_idx);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  // zero both packed buffers once: the padding records must stay {0,0,0,0}
  const size_t pk_bytes = sizeof(double) * kBveRec * (size_t)(round_up_chunk(s->nf) + kChunk);
  LPMX_CUDA(h, cudaMemsetAsync(s->packed[0], 0, pk_bytes, h->stream));
  LPMX_CUDA(h, cudaMemsetAsync(s->packed[1], 0, pk_bytes, h->stream));
  s->has_state = true;
  if (!is_device_pointer(vx) || !is_device_pointer(fx)) LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

static int solver_get_state(SolverState* s, double* vx, double* vz, double* vu, double* vpsi, double* fx, double* fz,
                            double* fu, double* fpsi, int layout, long vld, long fld) {
  lpmx_handle_t h = s->h;
  if (!s->has_state) return set_error(h, LPMX_ERR_STATE, "get_state before set_state");
  if (layout != LPMX_LAYOUT_LEFT && layout != LPMX_LAYOUT_RIGHT) return set_error(h, LPMX_ERR_INVALID, "unknown layout");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  // every rank returns the full state -- unless the host I/O is sharded: then each rank writes back its own target rows only
  const bool shard_out = h->io_sharded && h->world > 1;
  if (!shard_out) {
    LPMX_TRY(exchange_rows(s, s->X, 3));
    LPMX_TRY(exchange_rows(s, s->U, 3));
    LPMX_TRY(exchange_rows(s, s->Z, 1));
    if (vpsi || fpsi) LPMX_TRY(exchange_rows(s, s->Psi, 1));
  }
  auto vb = [&](long ld, int n) {
    return (layout == LPMX_LAYOUT_LEFT ? (size_t)(2 * ld + n) : (size_t)3 * n) * sizeof(double);
  };
  void *dvx = nullptr, *dvz = nullptr, *dvu = nullptr, *dvp = nullptr, *dfx = nullptr, *dfz = nullptr, *dfu = nullptr,
       *dfp = nullptr;
  if (vx) LPMX_TRY(stage_out_begin(h, "go_vx", vx, vb(vld, s->nv), &dvx));
  if (vz) LPMX_TRY(stage_out_begin(h, "go_vz", vz, sizeof(double) * s->nv, &dvz));
  if (vu) LPMX_TRY(stage_out_begin(h, "go_vu", vu, vb(vld, s->nv), &dvu));
  if (vpsi) LPMX_TRY(stage_out_begin(h, "go_vp", vpsi, sizeof(double) * s->nv, &dvp));
  if (fx) LPMX_TRY(stage_out_begin(h, "go_fx", fx, vb(fld, s->nf), &dfx));
  if (fz) LPMX_TRY(stage_out_begin(h, "go_fz", fz, sizeof(double) * s->nf, &dfz));
  if (fu) LPMX_TRY(stage_out_begin(h, "go_fu", fu, vb(fld, s->nf), &dfu));
  if (fpsi) LPMX_TRY(stage_out_begin(h, "go_fp", fpsi, sizeof(double) * s->nf, &dfp));
  if (s->nt > 0) {
    const int threads = 256;
    const int blocks = (s->nt + threads - 1) / threads;
    export_state_kernel<<<blocks, threads, 0, h->stream>>>(
        s->nv, s->nf, make_view((double*)dvx, layout, vld), (double*)dvz, make_view((double*)dvu, layout, vld),
        (double*)dvp, make_view((double*)dfx, layout, fld), (double*)dfz, make_view((double*)dfu, layout, fld),
        (double*)dfp, s->X, s->Z, s->U, s->Psi);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  bool any_host = false;
  auto out = [&](void* user, void* dev, size_t bytes, long ld, const std::vector<std::pair<int, int>>& runs, int ncomp) -> int {
    if (user && user != dev) any_host = true;
    if (shard_out && user && user != dev) return copy_runs(h, dev, user, layout, ld, runs, ncomp, false);
    return stage_out_end(h, user, dev, bytes);
  };
  LPMX_TRY(out(vx, dvx, vb(vld, s->nv), vld, s->own_v_runs, 3));
  LPMX_TRY(out(vz, dvz, sizeof(double) * s->nv, 0, s->own_v_runs, 1));
  LPMX_TRY(out(vu, dvu, vb(vld, s->nv), vld, s->own_v_runs, 3));
  LPMX_TRY(out(vpsi, dvp, sizeof(double) * s->nv, 0, s->own_v_runs, 1));
  LPMX_TRY(out(fx, dfx, vb(fld, s->nf), fld, s->own_f_runs, 3));
  LPMX_TRY(out(fz, dfz, sizeof(double) * s->nf, 0, s->own_f_runs, 1));
  LPMX_TRY(out(fu, dfu, vb(fld, s->nf), fld, s->own_f_runs, 3));
  LPMX_TRY(out(fpsi, dfp, sizeof(double) * s->nf, 0, s->own_f_runs, 1));
  if (any_host) LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

static StageArgs stage_args(SolverState* s, int part, const SumPlan* plan, int stage, int more, double dt, double Omega,
                            double* packed_next) {
  StageArgs a;
  if (plan && s->merged_now)
    a.pv = s->pv_m[part];
  else if (plan)
    a.pv = part_view(*plan, s->partials[part]);
  else
    a.pv = PartView{nullptr, 0, 0, 0, 1, 1};
  a.gid = s->gid[part];
  a.n_local = s->n_part[part];
  a.nv = s->nv;
  a.nt = s->nt;
  a.stage = stage;
  a.more = more;
  a.dt = dt;
  a.Omega = Omega;
  a.X = s->X, a.U = s->U, a.Xw = s->Xw, a.Z = s->Z, a.Zw = s->Zw, a.Psi = s->Psi;
  a.K1x = s->K[0][0], a.K1z = s->K[0][1];
  a.K2x = s->K[1][0], a.K2z = s->K[1][1];
  a.K3x = s->K[2][0], a.K3z = s->K[2][1];
  a.area = s->area;
  a.mask = s->mask;
  a.leaf_idx = s->leaf_idx;
  a.packed_next = packed_next;
  return a;
}

// pack (X,Z) of the resident state into packed[cur] and exchange
static int pack_resident(SolverState* s) {
  lpmx_handle_t h = s->h;
  const int part = s->split ? 0 : 1;  // only leaf faces have records: list A when the lists are split
  const int n_local = s->n_part[part];
  if (n_local > 0) {
    const int threads = 256, blocks = (n_local + threads - 1) / threads;
    pack_state_kernel<<<blocks, threads, 0, h->stream>>>(s->gid[part], n_local, s->nv, s->nt, s->X, s->Z, s->mask,
                                                         s->leaf_idx, s->area, s->packed[s->cur]);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  return exchange_packed(s, s->packed[s->cur]);
}

// One evaluation over this rank's targets: per list, pair sum -> stage kernel (`stage_fn(part)` launches it); the records the
// stage kernel of list A wrote into `next` are exchanged while list B is summed (see exchange_packed_begin).
// `merged` (optional): a plan over both lists at once, used instead of plans[0..1] -- see SolverState::partials_m.  The exchange
// of list A's records then starts after the whole sum instead of beside list B's, the price of one launch sequence that fills
// the chip where two would not (make_merged_plan).
template <class StageFn>
static int eval_lists(SolverState* st, const SumPlan* plans, double* tgt_base, double kappa, double* next, StageFn stage_fn,
                      const SumPlan* merged = nullptr) {
  lpmx_handle_t h = st->h;
  bool async = false;
  if (merged) {
    void* pm = nullptr;
    LPMX_TRY(dev_buffer(h, "solver_partials_m", plan_partials_bytes(*merged) + 256, &pm));
    st->partials_m = (double*)pm;
    LPMX_TRY(launch_pair_sum(h, *merged, st->view(tgt_base), st->self_idx, st->packed[st->cur], kappa, st->partials_m, 0.0, st->gid[0]));
    st->pv_m[0] = part_view(*merged, st->partials_m);
    st->pv_m[1] = part_view(*merged, st->partials_m + st->n_part[0]);
    st->merged_now = true;
    int rc = LPMX_OK;
    for (int part = 0; part < 2 && rc == LPMX_OK; ++part) {
      rc = stage_fn(part);
      if (rc == LPMX_OK && part == 0 && next) rc = exchange_packed_begin(st, next, &async);
    }
    st->merged_now = false;
    LPMX_TRY(rc);
    if (next) {
      LPMX_TRY(exchange_packed_end(st, async));
      st->cur ^= 1;
    }
    return LPMX_OK;
  }
  for (int part = 0; part < 2; ++part) {
    if (st->n_part[part] > 0) {
      LPMX_TRY(ensure_partials(st, part, plans[part]));
      LPMX_TRY(launch_pair_sum(h, plans[part], st->view(tgt_base), st->self_idx, st->packed[st->cur], kappa, st->partials[part],
                               0.0, st->gid[part]));
      LPMX_TRY(stage_fn(part));
    }
    if (part == 0 && st->split && next) LPMX_TRY(exchange_packed_begin(st, next, &async));
  }
  if (next) {
    LPMX_TRY(exchange_packed_end(st, async));
    st->cur ^= 1;
  }
  return LPMX_OK;
}

// "stage 1 of the first step from the resident velocity" for both lists, then the exchange of the records it wrote
template <class Launch>
static int prologue_lists(SolverState* st, Launch launch) {
  for (int part = 0; part < 2; ++part)
    if (st->n_part[part] > 0) LPMX_TRY(launch(part));
  return exchange_packed(st, st->packed[st->cur]);
}

static int make_list_plans(SolverState* st, int kind, SumPlan* plans) {
  for (int part = 0; part < 2; ++part)
    LPMX_TRY(make_plan(st->h, kind, st->n_part[part], st->n_leaf, &plans[part]));
  return LPMX_OK;
}

// One velocity plan over both lists, where that gets the bank path and the lists by themselves do not (a rank's share of a small
// mesh: at cubed-7 on eight GPUs 12 288 + 16 384 targets per rank; the bank path wants a few dozen CTAs per launch).  Returns
// whether to use it.  LPMX_MERGE_LISTS=0 / 1 overrides.
static bool make_merged_plan(SolverState* st, const SumPlan* plans, SumPlan* merged) {
  if (!st->split || st->n_part[0] == 0 || st->n_part[1] == 0) return false;
  const char* e = getenv("LPMX_MERGE_LISTS");
  if (e && e[0] == '0') return false;
  if (make_plan(st->h, kVel, st->n_part[0] + st->n_part[1], st->n_leaf, merged) != LPMX_OK) return false;
  if (merged->shape != kShapeConstStream) return false;
  if (e && e[0] == '1') return true;
  // measured (r2z, one GPU, a rank's target counts x 98 304 sources): 28 672 targets merged 1.62 ms against 0.77 + 1.04 ms for the
  // two lists through the ring kernel (neither fills the banks' pipeline); 57 344 merged 3.07 ms, 114 688 merged 5.98 ms against
  // 3.53 / 6.91 ms.  From 150 000 targets per rank the lists fill the pipeline by themselves and keep the overlapped exchange.
  (void)plans;
  return st->n_part[0] + st->n_part[1] < 150000;
}

}  // namespace lpmx

struct lpmx_bve_solver_s {
  SolverState st;
  SumPlan plan_vel[2], plan_psi[2];  // per target list
  SumPlan plan_vel_m;                // both lists at once (make_merged_plan)
  bool merged = false;
};
struct lpmx_ic2d_solver_s {
  SolverState st;
  SumPlan plan_vel[2], plan_velpsi[2], plan_psi[2];  // per target list
  SumPlan plan_vel_m;                                // both lists at once (make_merged_plan)
  bool merged = false;
  double eps = 0;
  // Lazy stream function.  psi of the new state is an OUTPUT of a step that no later step reads (quirk B-i), and the fused
  // velocity + psi evaluation costs 2.4 x the velocity one.  advance() therefore ends with the velocity-only kernel and marks
  // psi stale unless somebody read psi since the previous advance (psi_demanded: then the next advance fuses it again, which
  // is cheaper than a separate pass); a reader of a stale psi (get_state with a psi pointer) triggers ONE psi-only pass over
  // the retained state.  Same per-pair arithmetic either way; the sums differ by summation order only.
  bool psi_stale = false;
  bool psi_demanded = true;
};

extern "C" {

// ------------------------------------------------------------------------------------------------
// BVE
// ------------------------------------------------------------------------------------------------
int lpmx_local_targets(lpmx_handle_t h, int n_first, int n_second, const unsigned char* mask_second, int* idx, int* n_sources,
                       int* n_other) {
  if (!h || n_first < 0 || n_second < 0 || (n_second > 0 && !mask_second) || !idx) return LPMX_ERR_INVALID;
  int n_leaf = 0;
  for (int f = 0; f < n_second; ++f) n_leaf += mask_second[f] ? 0 : 1;
  std::vector<int> perm;
  std::vector<long> p_off;
  std::vector<std::pair<int, int>> vr, fr;
  int na = 0, nb = 0;
  build_target_lists(n_first, n_second, n_leaf, h->world, h->rank, [&](int f) { return mask_second[f] == 0; }, &perm, &p_off, &na,
                     &nb, &vr, &fr);
  for (int k = 0; k < na + nb; ++k) idx[k] = perm[p_off[h->rank] + k];
  if (n_sources) *n_sources = na;
  if (n_other) *n_other = nb;
  return LPMX_OK;
}

int lpmx_bve_solver_create(lpmx_handle_t h, int n_verts, int n_faces, lpmx_bve_solver_t* out) {
  if (!h || !out) return LPMX_ERR_INVALID;
  lpmx_bve_solver_s* s = new (std::nothrow) lpmx_bve_solver_s;
  if (!s) return LPMX_ERR_NOMEM;
  const int rc = solver_alloc(&s->st, h, n_verts, n_faces, 3);
  if (rc != LPMX_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return LPMX_OK;
}

int lpmx_bve_solver_destroy(lpmx_bve_solver_t s) {
  if (!s) return LPMX_OK;
  if (s->st.h && s->st.h->cached_bve == s) s->st.h->cached_bve = nullptr;
  solver_free(&s->st);
  delete s;
  return LPMX_OK;
}

int lpmx_bve_solver_set_state(lpmx_bve_solver_t s, const double* vx, const double* vz, const double* vu,
                              const double* fx, const double* fz, const double* fu, const double* fa,
                              const unsigned char* fm, int layout, long vld, long fld) {
  if (!s) return LPMX_ERR_INVALID;
  LPMX_TRY(solver_set_state(&s->st, vx, vz, vu, fx, fz, fu, fa, fm, layout, vld, fld, /*skip_self=*/1));
  LPMX_TRY(make_list_plans(&s->st, kVel, s->plan_vel));
  s->merged = make_merged_plan(&s->st, s->plan_vel, &s->plan_vel_m);
  LPMX_TRY(make_list_plans(&s->st, kPsi, s->plan_psi));
  return LPMX_OK;
}

int lpmx_bve_solver_get_state(lpmx_bve_solver_t s, double* vx, double* vz, double* vu, double* fx, double* fz,
                              double* fu, int layout, long vld, long fld) {
  if (!s) return LPMX_ERR_INVALID;
  return solver_get_state(&s->st, vx, vz, vu, nullptr, fx, fz, fu, nullptr, layout, vld, fld);
}

int lpmx_bve_solver_interactions_per_eval(lpmx_bve_solver_t s, double* local, double* global) {
  if (!s || !s->st.has_state) return LPMX_ERR_INVALID;
  const SolverState& st = s->st;
  // every target against every leaf, minus the self pair of each leaf face target
  const double nl = (double)st.n_leaf;
  if (global) *global = (double)st.nt * nl - nl;
  if (local) {
    // leaves among this rank's face targets
    const long l0 = st.packed_off[st.h->rank] / kBveRec, l1 = st.packed_off[st.h->rank + 1] / kBveRec;
    *local = (double)st.n_local() * nl - (double)(st.h->world > 1 ? (l1 - l0) : st.n_leaf);
  }
  return LPMX_OK;
}

// a psi-only pass over the lists (BVESphere::init_stream_fn, the lazy psi of the IC2D solver): no records are written
static int psi_pass(SolverState* st, const SumPlan* plans, double kappa) {
  return eval_lists(st, plans, st->X, kappa, nullptr, [&](int part) -> int {
    const int n = st->n_part[part], threads = 128, blocks = (n + threads - 1) / threads;
    psi_out_kernel<<<blocks, threads, 0, st->h->stream>>>(part_view(plans[part], st->partials[part]), st->gid[part], n, st->Psi);
    ++st->h->launches;
    return check_cuda(st->h, cudaGetLastError(), "psi_out_kernel launch");
  });
}

static int bve_eval(lpmx_bve_solver_s* s, int stage, int more, double dt, double Omega) {
  SolverState& st = s->st;
  lpmx_handle_t h = st.h;
  const bool writes_next = !(stage == 4 && !more);
  double* next = writes_next ? st.packed[st.cur ^ 1] : nullptr;
  return eval_lists(&st, s->plan_vel, st.Xw, 1.0, next, [&](int part) -> int {
    const StageArgs a = stage_args(&st, part, &s->plan_vel[part], stage, more, dt, Omega, next);
    const int threads = 128, blocks = (a.n_local + threads - 1) / threads;
    bve_rk4_stage_kernel<<<blocks, threads, 0, h->stream>>>(a);
    ++h->launches;
    return check_cuda(h, cudaGetLastError(), "bve_rk4_stage_kernel launch");
  }, s->merged ? &s->plan_vel_m : nullptr);
}

int lpmx_bve_solver_init_velocity(lpmx_bve_solver_t s) {
  if (!s) return LPMX_ERR_INVALID;
  SolverState& st = s->st;
  if (!st.has_state) return set_error(st.h, LPMX_ERR_STATE, "init_velocity before set_state");
  LPMX_CUDA(st.h, cudaSetDevice(st.h->device));
  // work state := state
  LPMX_CUDA(st.h, cudaMemcpyAsync(st.Xw, st.X, sizeof(double) * 3 * (size_t)st.nt, cudaMemcpyDeviceToDevice, st.h->stream));
  LPMX_CUDA(st.h, cudaMemcpyAsync(st.Zw, st.Z, sizeof(double) * (size_t)st.nt, cudaMemcpyDeviceToDevice, st.h->stream));
  LPMX_TRY(pack_resident(&st));
  return bve_eval(s, 4, 0, 0.0, 0.0);
}

int lpmx_bve_solver_stream_fn(lpmx_bve_solver_t s, double* vert_psi, double* face_psi) {
  if (!s) return LPMX_ERR_INVALID;
  SolverState& st = s->st;
  lpmx_handle_t h = st.h;
  if (!st.has_state) return set_error(h, LPMX_ERR_STATE, "stream_fn before set_state");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  LPMX_TRY(pack_resident(&st));
  LPMX_TRY(psi_pass(&st, s->plan_psi, 1.0));
  return solver_get_state(&st, nullptr, nullptr, nullptr, vert_psi, nullptr, nullptr, nullptr, face_psi,
                          LPMX_LAYOUT_RIGHT, 0, 0);
}

int lpmx_bve_solver_advance(lpmx_bve_solver_t s, double dt, double Omega, int n_steps) {
  if (!s) return LPMX_ERR_INVALID;
  SolverState& st = s->st;
  lpmx_handle_t h = st.h;
  if (!st.has_state) return set_error(h, LPMX_ERR_STATE, "advance before set_state");
  if (n_steps < 0) return set_error(h, LPMX_ERR_INVALID, "negative step count");
  if (n_steps == 0 || st.nt == 0) return LPMX_OK;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  // prologue: stage 1 of the first step from the resident velocity
  LPMX_TRY(prologue_lists(&st, [&](int part) -> int {
    const StageArgs a = stage_args(&st, part, nullptr, 0, 1, dt, Omega, st.packed[st.cur]);
    const int threads = 128, blocks = (a.n_local + threads - 1) / threads;
    bve_rk4_stage_kernel<<<blocks, threads, 0, h->stream>>>(a);
    ++h->launches;
    return check_cuda(h, cudaGetLastError(), "bve_rk4_stage_kernel launch");
  }));
  for (int step = 0; step < n_steps; ++step) {
    const int more = step + 1 < n_steps;
    for (int stage = 1; stage <= 4; ++stage) LPMX_TRY(bve_eval(s, stage, more, dt, Omega));
  }
  return LPMX_OK;
}

int lpmx_bve_rk4_step(lpmx_handle_t h, double dt, double Omega, int n_verts, double* vx, double* vz, double* vu,
                      int n_faces, double* fx, double* fz, double* fu, const double* fa, const unsigned char* fm,
                      int layout, long vld, long fld, int n_steps) {
  if (!h) return LPMX_ERR_INVALID;
  if ((n_verts > 0 && !vu) || (n_faces > 0 && !fu)) return set_error(h, LPMX_ERR_INVALID, "null velocity array");
  lpmx_bve_solver_t s = h->cached_bve;
  if (!s || s->st.nv != n_verts || s->st.nf != n_faces) {
    if (s) lpmx_bve_solver_destroy(s);
    h->cached_bve = nullptr;
    LPMX_TRY(lpmx_bve_solver_create(h, n_verts, n_faces, &s));
    h->cached_bve = s;
  }
  LPMX_TRY(lpmx_bve_solver_set_state(s, vx, vz, vu, fx, fz, fu, fa, fm, layout, vld, fld));
  LPMX_TRY(lpmx_bve_solver_advance(s, dt, Omega, n_steps));
  return lpmx_bve_solver_get_state(s, vx, vz, vu, fx, fz, fu, layout, vld, fld);
}

// ------------------------------------------------------------------------------------------------
// IC2D
// ------------------------------------------------------------------------------------------------
int lpmx_ic2d_solver_create(lpmx_handle_t h, int n_passive, int n_active, double eps, lpmx_ic2d_solver_t* out) {
  if (!h || !out) return LPMX_ERR_INVALID;
  lpmx_ic2d_solver_s* s = new (std::nothrow) lpmx_ic2d_solver_s;
  if (!s) return LPMX_ERR_NOMEM;
  s->eps = eps;
  const int rc = solver_alloc(&s->st, h, n_passive, n_active, 1);
  if (rc != LPMX_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return LPMX_OK;
}

int lpmx_ic2d_solver_destroy(lpmx_ic2d_solver_t s) {
  if (!s) return LPMX_OK;
  if (s->st.h && s->st.h->cached_ic2d == s) s->st.h->cached_ic2d = nullptr;
  solver_free(&s->st);
  delete s;
  return LPMX_OK;
}

int lpmx_ic2d_solver_set_state(lpmx_ic2d_solver_t s, const double* px, const double* pz, const double* pu,
                               const double* ax, const double* az, const double* au, const double* aa,
                               const unsigned char* am, int layout, long pld, long ald) {
  if (!s) return LPMX_ERR_INVALID;
  // Incompressible2DActiveSums skips the self term only when |eps| < DBL_EPSILON (:235)
  const int skip = std::fabs(s->eps) < DBL_EPSILON;
  LPMX_TRY(solver_set_state(&s->st, px, pz, pu, ax, az, au, aa, am, layout, pld, ald, skip));
  LPMX_TRY(make_list_plans(&s->st, kVel, s->plan_vel));
  s->merged = make_merged_plan(&s->st, s->plan_vel, &s->plan_vel_m);
  LPMX_TRY(make_list_plans(&s->st, kVelPsi, s->plan_velpsi));
  LPMX_TRY(make_list_plans(&s->st, kPsi, s->plan_psi));
  s->psi_stale = false;
  return LPMX_OK;
}

// one psi-only pass (kPsi) over the retained state: what the last evaluation of advance() left out
static int ic2d_refresh_psi(lpmx_ic2d_solver_s* s) {
  SolverState& st = s->st;
  lpmx_handle_t h = st.h;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  LPMX_TRY(pack_resident(&st));
  LPMX_TRY(psi_pass(&st, s->plan_psi, 1.0 + s->eps * s->eps));
  s->psi_stale = false;
  return LPMX_OK;
}

int lpmx_ic2d_solver_get_state(lpmx_ic2d_solver_t s, double* px, double* pz, double* pu, double* ppsi, double* ax,
                               double* az, double* au, double* apsi, int layout, long pld, long ald) {
  if (!s) return LPMX_ERR_INVALID;
  if (ppsi || apsi) {
    s->psi_demanded = true;
    if (s->psi_stale) LPMX_TRY(ic2d_refresh_psi(s));
  }
  return solver_get_state(&s->st, px, pz, pu, ppsi, ax, az, au, apsi, layout, pld, ald);
}

int lpmx_ic2d_solver_lazy_stream_fn(lpmx_ic2d_solver_t s, int demand_next) {
  if (!s) return LPMX_ERR_INVALID;
  s->psi_demanded = demand_next != 0;
  return LPMX_OK;
}

// one evaluation: stage 1 = predictor (velocity only; its psi is overwritten by the corrector in
// the reference, quirk B-i), stage 2 = velocity + stream function at the new state.  The same argument runs across
// steps: inside a multi-step call the psi of every step but the last is overwritten by the next step before anyone can
// read it, so only the final evaluation of the call uses the (2.4 x dearer) velocity + psi kernel.
static int ic2d_eval(lpmx_ic2d_solver_s* s, int stage, int more, double dt, double Omega, bool want_psi = true) {
  SolverState& st = s->st;
  lpmx_handle_t h = st.h;
  const bool with_psi = (stage == 2) && !more && want_psi;
  const SumPlan* plans = with_psi ? s->plan_velpsi : s->plan_vel;
  const double kappa = 1.0 + s->eps * s->eps;
  const bool writes_next = !(stage == 2 && !more);
  double* next = writes_next ? st.packed[st.cur ^ 1] : nullptr;
  return eval_lists(&st, plans, st.Xw, kappa, next, [&](int part) -> int {
    const StageArgs a = stage_args(&st, part, &plans[part], stage, more, dt, Omega, next);
    const int threads = 128, blocks = (a.n_local + threads - 1) / threads;
    if (with_psi)
      ic2d_rk2_stage_kernel<true><<<blocks, threads, 0, h->stream>>>(a);
    else
      ic2d_rk2_stage_kernel<false><<<blocks, threads, 0, h->stream>>>(a);
    ++h->launches;
    return check_cuda(h, cudaGetLastError(), "ic2d_rk2_stage_kernel launch");
  }, (!with_psi && s->merged) ? &s->plan_vel_m : nullptr);
}

int lpmx_ic2d_solver_init_direct_sums(lpmx_ic2d_solver_t s) {
  if (!s) return LPMX_ERR_INVALID;
  SolverState& st = s->st;
  if (!st.has_state) return set_error(st.h, LPMX_ERR_STATE, "init_direct_sums before set_state");
  LPMX_CUDA(st.h, cudaSetDevice(st.h->device));
  LPMX_CUDA(st.h, cudaMemcpyAsync(st.Xw, st.X, sizeof(double) * 3 * (size_t)st.nt, cudaMemcpyDeviceToDevice, st.h->stream));
  LPMX_CUDA(st.h, cudaMemcpyAsync(st.Zw, st.Z, sizeof(double) * (size_t)st.nt, cudaMemcpyDeviceToDevice, st.h->stream));
  LPMX_TRY(pack_resident(&st));
  return ic2d_eval(s, 2, 0, 0.0, 0.0);
}

int lpmx_ic2d_solver_advance(lpmx_ic2d_solver_t s, double dt, double Omega, int n_steps) {
  if (!s) return LPMX_ERR_INVALID;
  SolverState& st = s->st;
  lpmx_handle_t h = st.h;
  if (!st.has_state) return set_error(h, LPMX_ERR_STATE, "advance before set_state");
  if (n_steps < 0) return set_error(h, LPMX_ERR_INVALID, "negative step count");
  if (n_steps == 0 || st.nt == 0) return LPMX_OK;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  LPMX_TRY(prologue_lists(&st, [&](int part) -> int {
    const StageArgs a = stage_args(&st, part, nullptr, 0, 1, dt, Omega, st.packed[st.cur]);
    const int threads = 128, blocks = (a.n_local + threads - 1) / threads;
    ic2d_rk2_stage_kernel<false><<<blocks, threads, 0, h->stream>>>(a);
    ++h->launches;
    return check_cuda(h, cudaGetLastError(), "ic2d_rk2_stage_kernel launch");
  }));
  const bool eager_psi = s->psi_demanded;  // somebody read psi since the last advance: fuse it into the final evaluation
  for (int step = 0; step < n_steps; ++step) {
    const int more = step + 1 < n_steps;
    LPMX_TRY(ic2d_eval(s, 1, more, dt, Omega));
    LPMX_TRY(ic2d_eval(s, 2, more, dt, Omega, eager_psi));
  }
  s->psi_stale = !eager_psi;
  s->psi_demanded = false;
  return LPMX_OK;
}

int lpmx_ic2d_solver_totals(lpmx_ic2d_solver_t s, double* total_vorticity, double* total_kinetic_energy,
                            double* total_enstrophy) {
  if (!s) return LPMX_ERR_INVALID;
  SolverState& st = s->st;
  lpmx_handle_t h = st.h;
  if (!st.has_state) return set_error(h, LPMX_ERR_STATE, "totals before set_state");
  LPMX_CUDA(h, cudaSetDevice(h->device));
  // every rank sums over all faces: gather the rows the other ranks own first
  LPMX_TRY(exchange_rows(&st, st.U, 3));
  LPMX_TRY(exchange_rows(&st, st.Z, 1));
  double out[3] = {0, 0, 0};
  if (st.nf > 0) {
    Vec3View u = st.view(st.U);
    u.p += st.nv;
    LPMX_TRY(ic2d_totals_device(h, st.nf, st.Z + st.nv, u, st.area, st.mask, out));
  }
  if (total_vorticity) *total_vorticity = out[0];
  if (total_enstrophy) *total_enstrophy = 0.5 * out[1];
  if (total_kinetic_energy) *total_kinetic_energy = 0.5 * out[2];
  return LPMX_OK;
}

int lpmx_ic2d_rk2_step(lpmx_handle_t h, double dt, double Omega, double eps, int n_passive, double* px, double* pz,
                       double* pu, double* ppsi, int n_active, double* ax, double* az, double* au, double* apsi,
                       const double* aa, const unsigned char* am, int layout, long pld, long ald, int n_steps) {
  if (!h) return LPMX_ERR_INVALID;
  if ((n_passive > 0 && !pu) || (n_active > 0 && !au)) return set_error(h, LPMX_ERR_INVALID, "null velocity array");
  lpmx_ic2d_solver_t s = h->cached_ic2d;
  if (!s || s->st.nv != n_passive || s->st.nf != n_active || s->eps != eps) {
    if (s) lpmx_ic2d_solver_destroy(s);
    h->cached_ic2d = nullptr;
    LPMX_TRY(lpmx_ic2d_solver_create(h, n_passive, n_active, eps, &s));
    h->cached_ic2d = s;
  }
  LPMX_TRY(lpmx_ic2d_solver_set_state(s, px, pz, pu, ax, az, au, aa, am, layout, pld, ald));
  s->psi_demanded = (ppsi != nullptr) || (apsi != nullptr);  // the caller's psi views are outputs of this very call
  LPMX_TRY(lpmx_ic2d_solver_advance(s, dt, Omega, n_steps));
  return lpmx_ic2d_solver_get_state(s, px, pz, pu, ppsi, ax, az, au, apsi, layout, pld, ald);
}

}  // extern "C"
