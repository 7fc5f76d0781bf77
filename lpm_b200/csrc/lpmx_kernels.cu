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
//   * sources are pre-packed, leaves only, as 32-byte records {y0,y1,y2,Gamma} (48 bytes
//     {y, Gamma_zeta, Gamma_sigma, 0} for kSwe), zero-padded to a multiple of kChunk;
//   * a dedicated producer warp streams kChunk-source tiles into a kStages-deep shared-memory
//     ring with 1-D bulk TMA (cp.async.bulk ... mbarrier::complete_tx) -- full/empty mbarriers,
//     no __syncthreads in the steady state;
//   * every compute thread keeps T targets in registers and reads each source once per T
//     targets with broadcast LDS.128;
//   * the cross product / projection is linear in the source, so it is pulled out of the sum:
//     per pair only w = Gamma/d and M += w*y are evaluated (10 FP64-pipe instructions for the
//     24-flop reference pair), and u = x cross M happens once per target in the finalize kernel;
//   * the reciprocal is MUFU.RCP64H (rcp.approx.ftz.f64, 20-bit seed) plus one cubic
//     Newton step: r = r0*(1 + e + e^2), e = 1 - d*r0, |error| <= |e|^3 ~ 2^-57;
//   * work = (target block) x (source chunk) items in target-block-major order, split evenly
//     over a persistent grid (stream-K).  A CTA flushes its register accumulators to a partial
//     slot whenever it leaves a target block; the O(N) finalize kernels add the slots in a fixed
//     order (deterministic) and fuse all per-target algebra of the RK stage.
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstdarg>

#include "lpmx_internal.h"

namespace lpmx {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA, SASS UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 20-bit reciprocal seed (MUFU.RCP64H)
__device__ __forceinline__ double rcp_seed(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
}

// ------------------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------------------
struct SumArgs {
  Vec3View tgt;
  const int* self_idx;
  const double* packed;
  double* part;
  int n_tgt;
  int n_tb;
  int n_sc;
  long n_tgt_pad;
  double kappa;  // 1 + eps^2
};

__host__ __device__ __forceinline__ int cta_of_item(long item, int grid, long n_items) {
  return (int)(((item + 1) * (long)grid - 1) / n_items);
}

// Per-kind pair bodies.  x = target, (y, g..) = source record, acc = this target's accumulators.
// CHECK: compare the source's global compact index with the target's own (self) index.
template <int KIND, bool CHECK>
struct Pair;

template <bool CHECK>
struct Pair<kVel, CHECK> {
  __device__ __forceinline__ static void apply(const double* x, const double* /*kx*/, double kappa, const double* s,
                                               int j, int self, double* acc) {
    const double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    const double r0 = rcp_seed(d);
    const double g = s[3] * r0;
    const double e = fma(-d, r0, 1.0);
    const double p = fma(e, e, e);
    double w = fma(g, p, g);
    if (CHECK) w = (j == self) ? 0.0 : w;
    acc[0] = fma(w, s[0], acc[0]);
    acc[1] = fma(w, s[1], acc[1]);
    acc[2] = fma(w, s[2], acc[2]);
  }
};

template <bool CHECK>
struct Pair<kVelPsi, CHECK> {
  __device__ __forceinline__ static void apply(const double* x, const double*, double kappa, const double* s, int j,
                                               int self, double* acc) {
    double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    double gam = s[3];
    if (CHECK) {
      const bool me = (j == self);
      d = me ? 1.0 : d;
      gam = me ? 0.0 : gam;
    }
    const double r0 = rcp_seed(d);
    const double g = gam * r0;
    const double e = fma(-d, r0, 1.0);
    const double p = fma(e, e, e);
    const double w = fma(g, p, g);
    acc[0] = fma(w, s[0], acc[0]);
    acc[1] = fma(w, s[1], acc[1]);
    acc[2] = fma(w, s[2], acc[2]);
    acc[3] = fma(gam, log(d), acc[3]);
  }
};

template <bool CHECK>
struct Pair<kPsi, CHECK> {
  __device__ __forceinline__ static void apply(const double* x, const double*, double kappa, const double* s, int j,
                                               int self, double* acc) {
    double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    double gam = s[3];
    if (CHECK) {
      const bool me = (j == self);
      d = me ? 1.0 : d;
      gam = me ? 0.0 : gam;
    }
    acc[0] = fma(gam, log(d), acc[0]);
  }
};

// kSwe accumulators: [0..2] Mz = sum Gz y/d, [3..5] Ms = sum Gs y/d, [6..14] G (row-major):
//   G_ab += (Gz/d^2) c_a q_b - (Gs/d^2) q_a p_b,  c = x cross y, q = kappa x - y, p = y - (x.y) x
// The 1/d parts of the coded gradient polynomials ([y]x / d and (x.y) P / d) are linear in y and
// are rebuilt from Mz and Ms in the finalize kernel.
template <bool CHECK>
struct Pair<kSwe, CHECK> {
  __device__ __forceinline__ static void apply(const double* x, const double* kx, double kappa, const double* s,
                                               int j, int self, double* acc) {
    double d = fma(-x[0], s[0], fma(-x[1], s[1], fma(-x[2], s[2], kappa)));
    double gz = s[3], gs = s[4];
    if (CHECK) {
      const bool me = (j == self);
      d = me ? 1.0 : d;
      gz = me ? 0.0 : gz;
      gs = me ? 0.0 : gs;
    }
    const double xy = kappa - d;
    const double r0 = rcp_seed(d);
    const double e = fma(-d, r0, 1.0);
    const double pp = fma(e, e, e);
    const double r = fma(r0, pp, r0);
    const double wz = gz * r, ws = gs * r;
    acc[0] = fma(wz, s[0], acc[0]);
    acc[1] = fma(wz, s[1], acc[1]);
    acc[2] = fma(wz, s[2], acc[2]);
    acc[3] = fma(ws, s[0], acc[3]);
    acc[4] = fma(ws, s[1], acc[4]);
    acc[5] = fma(ws, s[2], acc[5]);
    const double wz2 = wz * r, ws2 = ws * r;
    double c[3], q[3], p[3];
    c[0] = fma(x[1], s[2], -(x[2] * s[1]));
    c[1] = fma(x[2], s[0], -(x[0] * s[2]));
    c[2] = fma(x[0], s[1], -(x[1] * s[0]));
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      q[b] = kx[b] - s[b];
      p[b] = fma(-xy, x[b], s[b]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double cz = wz2 * c[a];
      const double qs = ws2 * q[a];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        acc[6 + 3 * a + b] = fma(cz, q[b], acc[6 + 3 * a + b]);
        acc[6 + 3 * a + b] = fma(-qs, p[b], acc[6 + 3 * a + b]);
      }
    }
  }
};

template <int KIND, int T, bool CHECK>
__device__ __forceinline__ void chunk_loop(const double (*x)[3], const double (*kx)[3], double kappa,
                                           const double* __restrict__ sp, int j0, const int* self,
                                           double (*acc)[kind_nacc(KIND)]) {
  constexpr int REC = kind_rec(KIND);
#pragma unroll 2
  for (int j = 0; j < kChunk; ++j) {
    double s[REC];
    const double2* s2 = reinterpret_cast<const double2*>(sp + (size_t)j * REC);
#pragma unroll
    for (int v = 0; v < REC / 2; ++v) {
      const double2 t = s2[v];
      s[2 * v] = t.x;
      s[2 * v + 1] = t.y;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) Pair<KIND, CHECK>::apply(x[t], kx[t], kappa, s, j0 + j, self[t], acc[t]);
  }
}

template <int KIND, int T>
__global__ void __launch_bounds__(kCtaThreads, (KIND == kVel ? 2 : 1)) pair_sum_kernel(const SumArgs a) {
  constexpr int REC = kind_rec(KIND);
  constexpr int NACC = kind_nacc(KIND);
  constexpr uint32_t kStageBytes = kChunk * REC * sizeof(double);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kStageBytes);
  uint64_t* empty = full + kStages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kComputeWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const long n_items = (long)a.n_tb * a.n_sc;
  const int grid = gridDim.x;
  const long it0 = ((long)blockIdx.x * n_items) / grid;
  const long it1 = ((long)(blockIdx.x + 1) * n_items) / grid;

  if (warp == kComputeWarps) {
    // ---- producer warp: one lane streams source chunks through the ring ----
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int sc = (int)(it0 % a.n_sc);
      for (long it = it0; it < it1; ++it) {
        mbar_wait(empty + s, ph ^ 1u);
        mbar_arrive_expect_tx(full + s, kStageBytes);
        tma_load_1d(stage + (size_t)s * kChunk * REC, a.packed + (size_t)sc * kChunk * REC, kStageBytes, full + s);
        if (++sc == a.n_sc) sc = 0;
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    return;
  }

  // ---- compute warps ----
  const int tid = threadIdx.x;  // 0 .. kLanesPerCta-1
  constexpr int TB = T * kLanesPerCta;
  int s = 0;
  uint32_t ph = 0;
  long it = it0;
  while (it < it1) {
    const int tb = (int)(it / a.n_sc);
    int sc = (int)(it - (long)tb * a.n_sc);
    const long it_end = min(it1, (long)(tb + 1) * a.n_sc);

    double x[T][3], kx[T][3], acc[T][NACC];
    int self[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const long tg = (long)tb * TB + t * kLanesPerCta + tid;
      const bool valid = tg < a.n_tgt;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        x[t][k] = valid ? a.tgt(tg, k) : 0.0;
        kx[t][k] = a.kappa * x[t][k];
      }
      self[t] = (valid && a.self_idx) ? a.self_idx[tg] : -1;
#pragma unroll
      for (int q = 0; q < NACC; ++q) acc[t][q] = 0.0;
    }

    for (; it < it_end; ++it, ++sc) {
      mbar_wait(full + s, ph);
      const double* sp = stage + (size_t)s * kChunk * REC;
      const int j0 = sc * kChunk;
      bool hit = false;
#pragma unroll
      for (int t = 0; t < T; ++t) hit |= (unsigned)(self[t] - j0) < (unsigned)kChunk;
      if (__any_sync(0xffffffffu, hit))
        chunk_loop<KIND, T, true>(x, kx, a.kappa, sp, j0, self, acc);
      else
        chunk_loop<KIND, T, false>(x, kx, a.kappa, sp, j0, self, acc);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      if (++s == kStages) {
        s = 0;
        ph ^= 1u;
      }
    }

    // flush this CTA's contribution to target block tb into its slot
    const int slot = blockIdx.x - cta_of_item((long)tb * a.n_sc, grid, n_items);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const long tg = (long)tb * TB + t * kLanesPerCta + tid;
#pragma unroll
      for (int q = 0; q < NACC; ++q) a.part[((long)slot * NACC + q) * a.n_tgt_pad + tg] = acc[t][q];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// planning + launch
// ------------------------------------------------------------------------------------------------
int round_up_chunk(int n) { return ((n + kChunk - 1) / kChunk) * kChunk; }

static int pick_T(int kind, int num_sms, int n_tgt, int n_sc) {
  // Use the largest T whose item count still gives every SM's two CTAs a few items each.
  const int t_max = (kind == kVel) ? 4 : (kind == kSwe ? 2 : 2);
  for (int T = t_max; T > 1; T /= 2) {
    const long n_tb = (n_tgt + (long)T * kLanesPerCta - 1) / ((long)T * kLanesPerCta);
    if (n_tb * n_sc >= 8L * num_sms) return T;
  }
  return 1;
}

int make_plan(lpmx_handle_t h, int kind, int n_tgt, int n_src, SumPlan* p) {
  if (n_tgt < 0 || n_src < 0) return set_error(h, LPMX_ERR_INVALID, "negative size");
  p->kind = kind;
  p->n_tgt = n_tgt;
  p->n_src_pad = round_up_chunk(n_src);
  p->n_sc = p->n_src_pad / kChunk;
  p->T = pick_T(kind, h->num_sms, n_tgt, p->n_sc > 0 ? p->n_sc : 1);
  p->tb = p->T * kLanesPerCta;
  p->n_tb = (n_tgt + p->tb - 1) / p->tb;
  p->n_tgt_pad = (long)p->n_tb * p->tb;
  const long n_items = (long)p->n_tb * p->n_sc;
  const int per_sm = (kind == kVel) ? 2 : 1;
  long g = (long)per_sm * h->num_sms;
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
  p->smem_bytes = (size_t)kStages * kChunk * kind_rec(kind) * sizeof(double) + 2 * kStages * sizeof(uint64_t);
  return LPMX_OK;
}

size_t plan_partials_bytes(const SumPlan& p) {
  return (size_t)p.max_slots * kind_nacc(p.kind) * (size_t)p.n_tgt_pad * sizeof(double);
}

template <int KIND, int T>
static int launch_impl(lpmx_handle_t h, const SumPlan& p, const SumArgs& a) {
  auto kern = pair_sum_kernel<KIND, T>;
  static bool attr_set = false;  // per template instance
  if (!attr_set) {
    LPMX_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    attr_set = true;
  }
  kern<<<p.grid, kCtaThreads, p.smem_bytes, h->stream>>>(a);
  ++h->launches;
  return check_cuda(h, cudaGetLastError(), "pair_sum_kernel launch");
}

int launch_pair_sum(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed,
                    double kappa, double* partials) {
  if (p.n_tgt == 0) return LPMX_OK;
  if (p.n_sc == 0) {
    // no sources: every partial sum is zero
    LPMX_CUDA(h, cudaMemsetAsync(partials, 0, plan_partials_bytes(p), h->stream));
    return LPMX_OK;
  }
  SumArgs a;
  a.tgt = tgt;
  a.self_idx = self_idx;
  a.packed = packed;
  a.part = partials;
  a.n_tgt = p.n_tgt;
  a.n_tb = p.n_tb;
  a.n_sc = p.n_sc;
  a.n_tgt_pad = p.n_tgt_pad;
  a.kappa = kappa;
#define LPMX_DISPATCH(K)                                  \
  switch (p.T) {                                          \
    case 1: return launch_impl<K, 1>(h, p, a);            \
    case 2: return launch_impl<K, 2>(h, p, a);            \
    default: break;                                       \
  }
  switch (p.kind) {
    case kVel:
      if (p.T == 4) return launch_impl<kVel, 4>(h, p, a);
      LPMX_DISPATCH(kVel);
      break;
    case kVelPsi: LPMX_DISPATCH(kVelPsi); break;
    case kPsi: LPMX_DISPATCH(kPsi); break;
    case kSwe: LPMX_DISPATCH(kSwe); break;
    default: break;
  }
#undef LPMX_DISPATCH
  return set_error(h, LPMX_ERR_INVALID, "no kernel for kind %d T %d", p.kind, p.T);
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
// FP64 peak probe: 8 independent DFMA chains per thread, no memory traffic
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a, double b) {
  double v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, b);
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
  for (int rep = 0; rep < 4; ++rep) {
    LPMX_CUDA(h, cudaEventRecord(e0, h->stream));
    dfma_probe_kernel<<<blocks, threads, 0, h->stream>>>((double*)d, iters, 0.999999, 1e-7);
    ++h->launches;
    LPMX_CUDA(h, cudaEventRecord(e1, h->stream));
    LPMX_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0;
    LPMX_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  if (tflops) *tflops = flops / (best * 1e-3) * 1e-12;
  if (ms_out) *ms_out = best;
  return LPMX_OK;
}

}  // namespace lpmx
