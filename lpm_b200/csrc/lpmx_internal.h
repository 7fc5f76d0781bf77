// lpmx_internal.h -- shared declarations of the engine's translation units (not installed).
#ifndef LPMX_INTERNAL_H
#define LPMX_INTERNAL_H

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/lpmx.h"

namespace lpmx {

// ------------------------------------------------------------------------------------------------
// Pair-sum kernel configuration
// ------------------------------------------------------------------------------------------------
constexpr int kStages = 4;          // TMA ring depth
constexpr int kChunk = 256;         // sources per TMA stage

enum PairKind : int {
  kVel = 0,     // moment  M = sum Gamma y / d                    (BVE + IC2D velocity)
  kVelPsi = 1,  // M and   P = sum Gamma log d                    (IC2D velocity + stream function)
  kPsi = 2,     // P only                                         (BVE stream function)
  kSwe = 3,     // Mz, Ms, G[9]                                   (spherical SWE 12-tuple)
  kPlaneVelPsi = 4,  // u0, u1, sum G log a                       (planar IC2D velocity + stream function)
  kPlaneSwe = 5,     // u0, u1, du[4], lap, sum Gz log a, sum Gs log a   (planar SWE 9-tuple with the PSE Laplacian)
  kPlaneSweNoPot = 6,  // u0, u1, du[4], lap: the same without the two potentials (inner SWERK4 evaluations, where the
                       // reference computes psi and phi only to overwrite them before anyone can read them)
};
constexpr int kind_nacc(int k) {
  return k == kVel ? 3 : k == kVelPsi ? 4 : k == kPsi ? 1 : k == kSwe ? 15 : k == kPlaneVelPsi ? 3 : k == kPlaneSweNoPot ? 7 : 9;
}
// doubles per packed source record:
//   BVE / IC2D kinds  {y0, y1, y2, G*y0, G*y1, G*y2, G, 0}   (G = -zeta*A/(4 pi); 64 bytes)
//   SWE               {y0, y1, y2, Gz, Gs, 0}                 (48 bytes)
//   planar IC2D       {y0, y1, G, 0}                          (G = zeta*A/(2 pi); 32 bytes)
//   planar SWE        {y0, y1, Gz, Gs, s, A/(pi pse_eps^2)}   (Gz = zeta*A/(2 pi), Gs = sigma*A/(2 pi); 48 bytes)
constexpr int kBveRec = 8;
constexpr int kPlaneIc2dRec = 4;
constexpr int kPlaneSweRec = 6;
constexpr int kind_rec(int k) {
  return k == kSwe ? 6 : k == kPlaneVelPsi ? kPlaneIc2dRec : (k == kPlaneSwe || k == kPlaneSweNoPot) ? kPlaneSweRec : kBveRec;
}
constexpr bool kind_is_plane(int k) { return k == kPlaneVelPsi || k == kPlaneSwe || k == kPlaneSweNoPot; }

// strided accessor for Real*[3] views: element (i,k) at p[i*si + k*sk]
struct Vec3View {
  double* p;
  long si, sk;
  __host__ __device__ double& operator()(long i, int k) const { return p[i * si + k * sk]; }
};
inline Vec3View make_view(const double* p, int layout, long ld) {
  Vec3View v;
  v.p = const_cast<double*>(p);
  if (layout == LPMX_LAYOUT_LEFT) {
    v.si = 1;
    v.sk = ld;
  } else {
    v.si = 3;
    v.sk = 1;
  }
  return v;
}

// Work decomposition of one pair-sum launch (stream-K over target blocks x source chunks).
struct SumPlan {
  int kind;
  int shape;        // index of the kernel instance (lpmx_kernels.cu: kShapes)
  int T;            // targets per thread
  int tb;           // targets per CTA block = T * compute lanes per CTA
  int n_tgt;        // targets evaluated by this launch
  int n_tb;         // target blocks
  int n_src_pad;    // packed sources incl. zero padding (multiple of kChunk)
  int n_sc;         // source chunks
  int grid;         // persistent CTAs
  int max_slots;    // partial-sum slots per target block
  long n_tgt_pad;   // n_tb * tb
  size_t smem_bytes;
  // Constant-bank plans only (shape == kShapeConstStream): the first cs_n_const targets go through the bank in launches of
  // cs_ctas CTAs (whole waves of the chip); the other rem.n_tgt targets -- what would leave a last wave mostly empty -- go through
  // the stream-K ring kernel (plan `rem`), whose slots are folded into the bank path's accumulators before the stage kernel
  // reads them (lpmx_const_stream.cu).
  int cs_ctas = 0;
  int cs_n_const = 0;
  int cs_batch = 0;  // source records per bank launch
  struct Rem {
    int shape = 0, T = 0, tb = 0, n_tgt = 0, n_tb = 0, grid = 0, max_slots = 0;
    long n_tgt_pad = 0;
    size_t smem_bytes = 0;
  } rem;
};

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// ---- peer exchange (lpmx_peer.cu): slabs mapped into every rank of the box with CUDA IPC ----
constexpr int kMaxPeers = 8;
struct PeerRegion {
  void* local = nullptr;  // base of this rank's allocation
  size_t bytes = 0;
  void* peer[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // rank q's allocation, mapped here
};
struct PeerState {
  bool enabled = false;
  unsigned long long epoch = 0;       // exchanges issued so far (identical on every rank)
  unsigned long long timeout_ns = 0;  // deadline of every spin in the exchange kernel
  unsigned long long* flags_local = nullptr;
  unsigned long long* flags_peer[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* host_err = nullptr;      // host-mapped: 1 + rank the kernel gave up waiting for, 0 = fine
  int* host_err_dev = nullptr;  // device alias of host_err
  int dead_peer = 0;            // sticky copy of *host_err: once a wait expired every later exchange fails fast
  std::vector<PeerRegion> regions;
  std::vector<void*> graveyard;  // exported allocations whose release waits for lpmx_destroy
};

}  // namespace lpmx

struct lpmx_handle_s {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t cs_stream = nullptr;  // bank refills of the constant-bank path (the copy stream may be busy with an exchange)
  std::string err;
  long launches = 0;
  int rank = 0, world = 1;
  bool io_sharded = false;    // lpmx_set_io_sharded: host arrays carry only this rank's target rows
  cudaEvent_t xchg_ev[2] = {nullptr, nullptr};  // overlapped exchange of the sharded steppers (compute -> copy stream -> compute)
  void* nccl_comm = nullptr;  // ncclComm_t when lpmx_comm_init succeeded
  void* nccl_lib = nullptr;   // dlopen handle
  lpmx::PeerState* peer = nullptr;  // lpmx_comm_enable_peer_exchange
  int const_stream = -1;            // lpmx_pair_sum_const_stream: -1 = LPMX_CONST_STREAM from the environment
  long cs_launches = 0;             // bank-kernel launches so far (lpmx_const_stream_launch_count)
  void* cs_graph_cache = nullptr;   // captured launch sequences of the constant-bank path (lpmx_const_stream.cu)
  std::vector<cudaEvent_t> cs_events;  // lpmx_const_stream.cu: repack, filled[kCsBanks], summed[kCsBanks]
  std::map<std::string, lpmx::DevBuf> bufs;  // named scratch buffers
  std::map<std::string, lpmx::DevBuf> pinned;  // named pinned host staging buffers
  // optional per-launch timing of the pair-sum kernel (lpmx_profile_enable)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  double prof_pairs = 0;
  // cached one-shot solvers for the in-place stepper entry points
  lpmx_bve_solver_t cached_bve = nullptr;
  lpmx_ic2d_solver_t cached_ic2d = nullptr;
  lpmx_swe_solver_t cached_swe = nullptr;
  lpmx_plane_solver_t cached_plane = nullptr;
};

namespace lpmx {

int set_error(lpmx_handle_t h, int code, const char* fmt, ...);
int check_cuda(lpmx_handle_t h, cudaError_t e, const char* what);
#define LPMX_CUDA(h, call)                                          \
  do {                                                              \
    int _rc = ::lpmx::check_cuda((h), (call), #call);               \
    if (_rc != LPMX_OK) return _rc;                                 \
  } while (0)
#define LPMX_TRY(call)               \
  do {                               \
    int _rc = (call);                \
    if (_rc != LPMX_OK) return _rc;  \
  } while (0)

// scratch management
int dev_buffer(lpmx_handle_t h, const char* name, size_t bytes, void** out);
int pinned_buffer(lpmx_handle_t h, const char* name, size_t bytes, void** out);
bool is_device_pointer(const void* p);

// Staged argument: if `user` is a host pointer, a device copy is made (inputs) or reserved
// (outputs) in a named scratch buffer; if it is a device pointer it is used in place.
int stage_in(lpmx_handle_t h, const char* name, const void* user, size_t bytes, const void** dev);
int stage_out_begin(lpmx_handle_t h, const char* name, void* user, size_t bytes, void** dev);
int stage_out_end(lpmx_handle_t h, void* user, const void* dev, size_t bytes);

// ---- pair-sum engine (lpmx_kernels.cu) ----
int make_plan(lpmx_handle_t h, int kind, int n_tgt, int n_src, SumPlan* plan, bool allow_const_stream = true, int force_T = 0);
int make_best_ring_plan(lpmx_handle_t h, int n_tgt, int n_src, SumPlan* plan);  // velocity kind, least modelled time over the shapes
size_t plan_partials_bytes(const SumPlan& p);
// modelled duration [s] of the ring kernel's launch for `p` (padded work / measured rate of the shape + a fixed ramp): what
// the constant-bank planner compares its own launches with
double ring_plan_seconds(const SumPlan& p);
// tgt: target coordinates of the n_tgt targets of this launch (view indexed from 0);
// self_idx: compact source index of each target's own particle or -1 (may be nullptr);
// packed: n_src_pad records of kind_rec(kind) doubles; partials: plan_partials_bytes.
// kappa: 1 + eps^2 on the sphere, eps^2 in the plane; aux: 1 / pse_eps^2 for kPlaneSwe (unused otherwise).
// Planar kinds read target rows (x0, x1, surface height) through the same 3-row view.
// tgt_map (optional): the launch's targets are the elements tgt_map[0 .. plan.n_tgt) of `tgt` / `self_idx`
int launch_pair_sum(lpmx_handle_t h, const SumPlan& plan, Vec3View tgt, const int* self_idx, const double* packed,
                    double kappa, double* partials, double aux = 0.0, const int* tgt_map = nullptr);

// ---- velocity pair sum through the constant bank (lpmx_const_stream.cu; opt-in) ----
constexpr int kShapeConstStream = 1000;  // SumPlan::shape of a launch that takes this path
// fills *p and returns true when the path is switched on and the launch is large enough for it (kVel only)
bool make_const_plan(lpmx_handle_t h, int n_tgt, int n_src, SumPlan* p);
int launch_const_stream(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed, double kappa,
                        double* partials, const int* tgt_map);
constexpr int kCsMaxThreads = 544;  // 16 compute warps + the prefetch warp (T = 3); per T: cs_max_threads, lpmx_const_bank.cuh
// the split of n_tgt targets x n_src sources between the bank path and the ring kernel with the least modelled time:
// *T_out targets per thread x *nw_out warps, *ctas_out CTAs per bank launch covering the first *n_const_out targets;
// returns the modelled seconds of the whole evaluation (bank launches + remainder), *ring_s_out those of the ring kernel alone
double pick_const_split(lpmx_handle_t h, int num_sms, int n_tgt, int n_src, int* T_out, int* nw_out, int* ctas_out, int* n_const_out,
                        double* ring_s_out);
// the banks (lpmx_const_bank.cu compiled kCsBanks times, lpm_b200/build.py: N_CONST_BANKS)
constexpr int kCsBanks = 24;
#define LPMX_CS_BANK_LIST(X) \
  X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23)
namespace cs { struct CsArgs; }
#define LPMX_CS_DECLARE(k)                                                                                               \
  cudaError_t cs_bank_launch_##k(int T, int grid, int threads, cudaStream_t stream, const cs::CsArgs& a, int pdl); \
  cudaError_t cs_bank_fill_##k(const double* records, int n_rec, cudaStream_t stream);
LPMX_CS_BANK_LIST(LPMX_CS_DECLARE)
#undef LPMX_CS_DECLARE
int const_stream_mode(lpmx_handle_t h);
void const_stream_teardown(lpmx_handle_t h);

// exclusive scan of !mask -> leaf_idx, returns number of unmasked sources (host sync)
int scan_leaves(lpmx_handle_t h, const unsigned char* mask_dev, int n, int* leaf_idx_dev, int* n_leaves);
int round_up_chunk(int n);
int fp64_probe(lpmx_handle_t h, double* tflops, double* ms_out);

// leaf totals [sum zeta A, sum zeta^2 A, sum |u|^2 A] of n faces -> host (lpmx_diagnostics.cu; synchronises the stream)
int ic2d_totals_device(lpmx_handle_t h, int n, const double* zeta, Vec3View u, const double* area, const unsigned char* mask,
                       double* out3_host);

// In-place allgatherv of doubles on the handle's stream: rank r owns elements
// [offsets[r], offsets[r+1]) of `base`; after the call every rank holds all of them.
// No-op for world == 1.  (lpmx_core.cu; NCCL broadcasts grouped into one launch.)
int comm_allgatherv(lpmx_handle_t h, double* base, const long* offsets, cudaStream_t stream = nullptr);  // null: h->stream

// Solver slabs.  Without the peer exchange these are cudaMalloc / cudaFree; with it the slab is also mapped
// into every rank (collective call) so that comm_allgatherv can store into the peers directly.  (lpmx_peer.cu)
int slab_alloc(lpmx_handle_t h, void** out, size_t bytes);
void slab_free(lpmx_handle_t h, void* p);
int peer_enable(lpmx_handle_t h, int enable);
bool peer_can_exchange(lpmx_handle_t h, const double* base);
int peer_allgatherv(lpmx_handle_t h, double* base, const long* offsets, cudaStream_t stream = nullptr);
int peer_check_error(lpmx_handle_t h);
void peer_teardown(lpmx_handle_t h);

}  // namespace lpmx

#endif
