// lpmx_kernels.cu -- the O(N^2) pair-sum kernel family for sm_100a and its launch planning.
//
// One kernel template serves the three direct-sum families of the reference:
//   kVel     BVE / IC2D velocity      biot_savart, lpm_sphere_functions.hpp:44-57;
//                                     kernel_vals r[0..2], lpm_incompressible2d_kernels.hpp:35-47
//   kVelPsi  IC2D velocity + psi      kernel_vals r[0..3]
//   kPsi     BVE stream function      greens_fn, lpm_sphere_functions.hpp:21-29
//   kSwe     SWE 12-tuple             sphere_swe_velocity_sums, lpm_swe_kernels.hpp:334-362
//
// Design:
//   * sources are pre-packed, leaves only, as 64-byte records {y, Gamma*y, Gamma, 0} (48 bytes
//     {y, Gamma_zeta, Gamma_sigma, 0} for kSwe), zero-padded to a multiple of kChunk;
//   * a dedicated producer warp streams kChunk-source tiles into a kStages-deep shared-memory
//     ring with 1-D bulk TMA (cp.async.bulk ... mbarrier::complete_tx) -- full/empty mbarriers,
//     no __syncthreads in the steady state;
//   * every compute thread keeps T targets in registers and reads each source once per T
//     targets with broadcast LDS.128;
//   * the cross product / projection is linear in the source, so it is pulled out of the sum:
//     per pair only r = 1/d and M += r*(Gamma*y) are evaluated (9 FP64-pipe instructions for the
//     24-flop reference pair), and u = x cross M happens once per target in the finalize kernel;
//   * the reciprocal is MUFU.RCP64H (rcp.approx.ftz.f64, 20-bit seed) plus one cubic
//     Newton step: r = r0*(1 + e + e^2), e = 1 - d*r0, |error| <= |e|^3 ~ 2^-57;
//   * work = (target block) x (source chunk) items in target-block-major order, split evenly
//     over a persistent grid (stream-K).  A CTA flushes its register accumulators to a partial
//     slot whenever it leaves a target block; the O(N) finalize kernels add the slots in a fixed
//     order (deterministic) and fuse all per-target algebra of the RK stage.
#include <cub/device/device_scan.cuh>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <initializer_list>

#include "lpmx_internal.h"
#include "lpmx_pair_kernel.cuh"

namespace lpmx {

// ------------------------------------------------------------------------------------------------
// planning + launch
// ------------------------------------------------------------------------------------------------
int round_up_chunk(int n) { return ((n + kChunk - 1) / kChunk) * kChunk; }

// One launchable kernel instance.
struct Shape {
  int kind, T, nw, per_sm;
  int (*launch)(lpmx_handle_t, const SumPlan&, const SumArgs&);
};

template <class C>
static int launch_cfg(lpmx_handle_t h, const SumPlan& p, const SumArgs& a) {
  auto kern = pair_sum_kernel<C>;
  // per template instance AND per device: the attribute belongs to the function in the device's context, and one process
  // may hold handles on several GPUs (lpmx_create(device_id))
  // one bit per device, atomic: handles may be driven from several host threads; setting the attribute twice is harmless
  static std::atomic<unsigned long long> attr_set{0ull};
  const int dev = h->device;
  const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
  if (!bit || !(attr_set.load(std::memory_order_acquire) & bit)) {
    LPMX_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    if (bit) attr_set.fetch_or(bit, std::memory_order_release);
  }
  kern<<<p.grid, C::THREADS, p.smem_bytes, h->stream>>>(a);
  ++h->launches;
  return check_cuda(h, cudaGetLastError(), "pair_sum_kernel launch");
}

#define LPMX_SHAPE(K, T, NW, MINB, UNR) \
  { K, T, NW, MINB, &launch_cfg<PairCfg<K, T, NW, MINB, UNR>> }
// Per kind: preferred (largest target block) first; smaller blocks keep all SMs busy on small meshes.
// Shapes were picked with tools/tune_pair_sum.cu on a B200 (profiles/r1b_tune_shapes.txt, profiles/r1f_tune_psi.txt): the FP64 pipe sustains
// one DFMA per 2 cycles per SM sub-partition only while an instruction reads <= 2 distinct 64-bit register
// operands (3 distinct: 3 cycles), so the winner is the shape with the most operand reuse across the T
// targets of a thread that still leaves 8 warps per SM to cover MUFU/LDS latency.
static const Shape kShapes[] = {
    LPMX_SHAPE(kVel, 6, 8, 1, 4),     LPMX_SHAPE(kVel, 4, 8, 2, 2),    LPMX_SHAPE(kVel, 2, 8, 2, 2),
    LPMX_SHAPE(kVel, 1, 8, 2, 2),
    LPMX_SHAPE(kVelPsi, 4, 8, 1, 2),  LPMX_SHAPE(kVelPsi, 2, 8, 2, 2), LPMX_SHAPE(kVelPsi, 1, 8, 2, 2),
    LPMX_SHAPE(kPsi, 8, 8, 1, 2),     LPMX_SHAPE(kPsi, 4, 8, 2, 2),    LPMX_SHAPE(kPsi, 2, 8, 2, 2),
    LPMX_SHAPE(kPsi, 1, 8, 2, 2),
    LPMX_SHAPE(kSwe, 2, 8, 1, 2),     LPMX_SHAPE(kSwe, 1, 8, 1, 2),
    LPMX_SHAPE(kPlaneVelPsi, 4, 8, 1, 2), LPMX_SHAPE(kPlaneVelPsi, 2, 8, 2, 2), LPMX_SHAPE(kPlaneVelPsi, 1, 8, 2, 2),
    LPMX_SHAPE(kPlaneSwe, 2, 8, 1, 2),    LPMX_SHAPE(kPlaneSwe, 1, 8, 2, 2),
    LPMX_SHAPE(kPlaneSweNoPot, 2, 8, 1, 2), LPMX_SHAPE(kPlaneSweNoPot, 1, 8, 2, 2),
};
constexpr int kNumShapes = sizeof(kShapes) / sizeof(kShapes[0]);

// Use the largest target block whose item count still gives every resident CTA a few items.
static int pick_shape(int kind, int num_sms, int n_tgt, int n_sc) {
  int last = -1;
  for (int i = 0; i < kNumShapes; ++i) {
    if (kShapes[i].kind != kind) continue;
    last = i;
    const long tb = (long)kShapes[i].T * kShapes[i].nw * 32;
    const long n_tb = (n_tgt + tb - 1) / tb;
    if (n_tb * n_sc >= 8L * kShapes[i].per_sm * num_sms) return i;
  }
  return last;
}

int make_plan(lpmx_handle_t h, int kind, int n_tgt, int n_src, SumPlan* p, bool allow_const_stream, int force_T) {
  if (n_tgt < 0 || n_src < 0) return set_error(h, LPMX_ERR_INVALID, "negative size");
  *p = SumPlan();
  if (kind == kVel && allow_const_stream && make_const_plan(h, n_tgt, n_src, p)) return LPMX_OK;  // sources through the constant bank
  *p = SumPlan();
  p->kind = kind;
  p->n_tgt = n_tgt;
  p->n_src_pad = round_up_chunk(n_src);
  p->n_sc = p->n_src_pad / kChunk;
  p->shape = -1;
  if (force_T > 0) {
    for (int i = 0; i < kNumShapes; ++i)
      if (kShapes[i].kind == kind && kShapes[i].T == force_T) p->shape = i;
  } else {
    p->shape = pick_shape(kind, h->num_sms, n_tgt, p->n_sc > 0 ? p->n_sc : 1);
  }
  if (p->shape < 0) return set_error(h, LPMX_ERR_INVALID, "no kernel for kind %d", kind);
  const Shape& sh = kShapes[p->shape];
  p->T = sh.T;
  p->tb = sh.T * sh.nw * 32;
  p->n_tb = (n_tgt + p->tb - 1) / p->tb;
  p->n_tgt_pad = (long)p->n_tb * p->tb;
  const long n_items = (long)p->n_tb * p->n_sc;
  long g = (long)sh.per_sm * h->num_sms;
  if (g > n_items) g = n_items;
  if (g < 1) g = 1;
  p->grid = (int)g;
  int ms = 1;
  if (n_items > 0) {
    // slots needed by the widest target block
    for (int tb = 0; tb < p->n_tb; ++tb) {
      const int c0 = cta_of_item((long)tb * p->n_sc, p->grid, n_items);
      const int c1 = cta_of_item((long)(tb + 1) * p->n_sc - 1, p->grid, n_items);
      if (c1 - c0 + 1 > ms) ms = c1 - c0 + 1;
    }
  }
  p->max_slots = ms;
  p->smem_bytes = pair_smem_bytes(kind, sh.T, sh.nw * 32);
  return LPMX_OK;
}

// The velocity shape with the least modelled time for a SMALL target set (the remainder of a constant-bank plan: a few
// thousand targets, where the general rule of pick_shape -- 8 items per resident CTA -- falls through to T = 1).
int make_best_ring_plan(lpmx_handle_t h, int n_tgt, int n_src, SumPlan* p) {
  double best = -1.0;
  SumPlan q;
  for (int T : {6, 4, 2, 1}) {
    if (make_plan(h, kVel, n_tgt, n_src, &q, false, T) != LPMX_OK) continue;
    const double t = ring_plan_seconds(q);
    if (best < 0 || t < best) best = t, *p = q;
  }
  return best < 0 ? set_error(h, LPMX_ERR_INVALID, "no velocity kernel") : LPMX_OK;
}

// Modelled duration of a ring-kernel launch: padded pairs over the measured rate of the shape (velocity kind, one B200:
// profiles/r1b_tune_shapes.txt, r1h_size_sweep.txt; T = 2 / 1 shapes from the small-mesh rows), stretched by the item
// quantisation of the persistent grid, plus a fixed ramp (launch, first tile, flush).  Only the constant-bank planner uses it,
// to decide which targets are better served by which kernel; it does not have to be better than ~10 %.
double ring_plan_seconds(const SumPlan& p) {
  if (p.n_tgt <= 0 || p.n_sc <= 0) return 0.0;
  const double rate = p.T >= 6 ? 1.637e12 : p.T >= 4 ? 1.48e12 : p.T >= 2 ? 1.4e12 : 1.3e12;  // T = 1: r2s, 2 048 targets in 0.187 ms
  const long n_items = (long)p.n_tb * p.n_sc;
  const double per_cta = (double)n_items / p.grid;
  const double quant = std::ceil(per_cta) / per_cta;
  return (double)p.n_tgt_pad * (double)p.n_src_pad / rate * quant + 25e-6;
}

size_t plan_partials_bytes(const SumPlan& p) {
  size_t b = (size_t)p.max_slots * kind_nacc(p.kind) * (size_t)p.n_tgt_pad * sizeof(double);
  // constant-bank plan with a ring remainder: the remainder's slots live behind the bank path's accumulators
  if (p.shape == kShapeConstStream && p.rem.n_tgt > 0)
    b += (size_t)p.rem.max_slots * kind_nacc(p.kind) * (size_t)p.rem.n_tgt_pad * sizeof(double);
  return b;
}

// the ring kernel on the remainder of a constant-bank plan: targets [cs_n_const, n_tgt) of the launch's views, slots behind
// the bank path's accumulators (lpmx_const_stream.cu folds them in)
int launch_ring_remainder(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed, double kappa,
                          double* rem_partials, const int* tgt_map) {
  SumPlan r;
  r.kind = p.kind;
  r.shape = p.rem.shape;
  r.T = p.rem.T;
  r.tb = p.rem.tb;
  r.n_tgt = p.rem.n_tgt;
  r.n_tb = p.rem.n_tb;
  r.n_src_pad = p.n_src_pad;
  r.n_sc = p.n_sc;
  r.grid = p.rem.grid;
  r.max_slots = p.rem.max_slots;
  r.n_tgt_pad = p.rem.n_tgt_pad;
  r.smem_bytes = p.rem.smem_bytes;
  SumArgs a;
  a.tgt = tgt;
  if (tgt_map) {  // the list's tail; the views stay (the list holds indices into them)
    a.tgt_map = tgt_map + p.cs_n_const;
    a.self_idx = self_idx;
  } else {
    a.tgt.p = tgt.p + (long)p.cs_n_const * tgt.si;
    a.tgt_map = nullptr;
    a.self_idx = self_idx ? self_idx + p.cs_n_const : nullptr;
  }
  a.packed = packed;
  a.part = rem_partials;
  a.n_tgt = r.n_tgt;
  a.n_tb = r.n_tb;
  a.n_sc = r.n_sc;
  a.n_tgt_pad = r.n_tgt_pad;
  a.kappa = kappa;
  a.aux = 0.0;
  return kShapes[r.shape].launch(h, r, a);
}

int launch_pair_sum(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed,
                    double kappa, double* partials, double aux, const int* tgt_map) {
  if (p.n_tgt == 0) return LPMX_OK;
  if (p.n_sc == 0) {
    // no sources: every partial sum is zero
    LPMX_CUDA(h, cudaMemsetAsync(partials, 0, plan_partials_bytes(p), h->stream));
    return LPMX_OK;
  }
  SumArgs a;
  a.tgt = tgt;
  a.tgt_map = tgt_map;
  a.self_idx = self_idx;
  a.packed = packed;
  a.part = partials;
  a.n_tgt = p.n_tgt;
  a.n_tb = p.n_tb;
  a.n_sc = p.n_sc;
  a.n_tgt_pad = p.n_tgt_pad;
  a.kappa = kappa;
  a.aux = aux;
  const bool cs = p.shape == kShapeConstStream;
  if (!h->profile) return cs ? launch_const_stream(h, p, tgt, self_idx, packed, kappa, partials, tgt_map) : kShapes[p.shape].launch(h, p, a);
  if (h->prof_used == h->prof_events.size()) {
    cudaEvent_t e0, e1;
    LPMX_CUDA(h, cudaEventCreate(&e0));
    LPMX_CUDA(h, cudaEventCreate(&e1));
    h->prof_events.push_back({e0, e1});
  }
  auto& ev = h->prof_events[h->prof_used++];
  LPMX_CUDA(h, cudaEventRecord(ev.first, h->stream));
  const int rc = cs ? launch_const_stream(h, p, tgt, self_idx, packed, kappa, partials, tgt_map) : kShapes[p.shape].launch(h, p, a);
  LPMX_CUDA(h, cudaEventRecord(ev.second, h->stream));
  h->prof_pairs += (double)p.n_tgt * (double)p.n_src_pad;
  return rc;
}

// ------------------------------------------------------------------------------------------------
// leaf scan (Faces::scan_leaves, mesh/lpm_faces_impl.hpp:100-124)
// ------------------------------------------------------------------------------------------------
__global__ void leaf_flags_kernel(const unsigned char* __restrict__ mask, int n, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = mask[i] ? 0 : 1;
}

int scan_leaves(lpmx_handle_t h, const unsigned char* mask_dev, int n, int* leaf_idx_dev, int* n_leaves) {
  *n_leaves = 0;
  if (n == 0) return LPMX_OK;
  leaf_flags_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(mask_dev, n, leaf_idx_dev);
  ++h->launches;
  LPMX_CUDA(h, cudaGetLastError());
  int* in = leaf_idx_dev;  // in-place exclusive scan
  size_t tmp_bytes = 0;
  LPMX_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, leaf_idx_dev, n, h->stream));
  void* tmp = nullptr;
  LPMX_TRY(dev_buffer(h, "scan_tmp", tmp_bytes + 16, &tmp));
  LPMX_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, leaf_idx_dev, n, h->stream));
  ++h->launches;
  int last_idx = 0;
  unsigned char last_mask = 0;
  LPMX_CUDA(h, cudaMemcpyAsync(&last_idx, leaf_idx_dev + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  LPMX_CUDA(h, cudaMemcpyAsync(&last_mask, mask_dev + (n - 1), 1, cudaMemcpyDeviceToHost, h->stream));
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_leaves = last_idx + (last_mask ? 0 : 1);
  return LPMX_OK;
}

// ------------------------------------------------------------------------------------------------
// FP64 peak probe: 8 independent DFMA chains per thread, no memory traffic.  v = fma(v, a, v) with `a` a kernel
// parameter: SASS `DFMA R, R, c[0x0][..], R` -- ONE distinct register operand per instruction, the multiplier from the
// constant bank (profiles/r1b_fp64_pipe_probe.txt: 98 % of 64 lanes/clk/SM; operands served by the register reuse cache
// already cost ~8 %, which is what the round-1 probe `fma(v, a, b)` with a, b in registers measured: 34.2 TFLOP/s).  The
// trip loop is unrolled 16 times (3 loop instructions per 128 DFMAs).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a) {
  double v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, v[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += v[k];
  if (s == 123.456) out[0] = s;  // never true; keeps the loop alive
}

int fp64_probe(lpmx_handle_t h, double* tflops, double* ms_out) {
  void* d = nullptr;
  LPMX_TRY(dev_buffer(h, "probe", 64, &d));
  const int blocks = h->num_sms * 8, threads = 256, iters = 1 << 16;
  cudaEvent_t e0, e1;
  LPMX_CUDA(h, cudaEventCreate(&e0));
  LPMX_CUDA(h, cudaEventCreate(&e1));
  double best = 1e30;
  for (int rep = 0; rep < 12; ++rep) {  // the first launches also ramp the SM clock up from idle
    LPMX_CUDA(h, cudaEventRecord(e0, h->stream));
    dfma_probe_kernel<<<blocks, threads, 0, h->stream>>>((double*)d, iters, 0x1p-60);
    ++h->launches;
    LPMX_CUDA(h, cudaEventRecord(e1, h->stream));
    LPMX_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    LPMX_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 3 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  if (tflops) *tflops = flops / (best * 1e-3) * 1e-12;
  if (ms_out) *ms_out = best;
  return LPMX_OK;
}

}  // namespace lpmx
