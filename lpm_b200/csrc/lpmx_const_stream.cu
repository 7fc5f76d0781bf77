// lpmx_const_stream.cu -- the velocity pair sum with the source records streamed through the CONSTANT bank, so that
// they reach the DFMAs as uniform-register operands (SASS: LDCU.64 + DFMA R, R, UR, R).
// Measured (profiles/r2b_*): icos-8 (2.40 M targets x 1.31 M sources) 7.604 -> 7.008 s per BVERK4 step, 1.657e12 -> 1.798e12
// interactions/s (+8.5 %); cubed-7 (229 376 targets) 54.9 -> 66.4 ms (-21 %: one launch per 640 sources is ~80 us of work
// there, all CTAs of a launch share the sources so a single wave is either quantised or unbalanced over the 4 schedulers).
// Hence the default (mode -1 "auto"): used for velocity launches with >= 1e6 targets per rank (LPMX_CONST_MIN_TARGETS), the
// stream-K ring kernel everywhere else.  LPMX_CONST_STREAM=0 turns it off, =1 forces overlapped copies, =2 copies on the
// compute stream; lpmx_pair_sum_const_stream() sets the same per handle.
//
// Why: pair_sum_kernel (lpmx_pair_kernel.cuh) is bound by FP64 issue and reaches 81 % of the pipe because 5 of its 9
// DFMAs per pair read a third distinct register operand (profiles/README.md, "Why 81 %").  A source record is the same
// for every thread; read from c[3][..] it costs no register-file port.  tools/const_probe.cu measured the body below at
// 1.82e12 pairs/s per fully loaded B200 against 1.65e12 for the shared-memory ring (profiles/r1j_const_bank_probe.txt).
//
// How: the 64 KB bank holds two halves of 640 records x 48 B {y0, y1, y2, G*y0, G*y1, G*y2}.  One launch sums ONE half
// into every target of this rank (accumulators live in slot 0 of the partials buffer between launches, 24 B per target
// per launch -- three orders of magnitude below the FP64 time), while a device-to-device copy on the handle's copy
// stream fills the other half for the next launch.  All CTAs of a launch read the same sources, so the work cannot be
// split over sources the way the stream-K kernel does; the chip is balanced by the launch shape instead: one CTA per SM,
// and (T = 5..7 targets per thread) x 8 warps chosen so that waves x 148 x 32 x T x NW covers the targets with the least
// excess.  That needs >= ~2e5 targets per rank; smaller launches keep the ring kernel.
//
// Same arithmetic per pair as Pair<kVel> (bit-identical terms); per target the terms are added in source order, so the
// sums differ from the stream-K kernel's by round-off only (the tolerance of every parity test covers both).
#include <cstdlib>
#include <mutex>

#include "lpmx_const_stream_body.h"
#include "lpmx_internal.h"

namespace lpmx {

namespace {

using cs::CsArgs;
constexpr int kCsHalf = cs::kHalf;  // records per half of the bank
constexpr int kCsRec = cs::kRec;    // doubles per record
constexpr int kCsMaxThreads = 384;
__constant__ double c_src[2 * kCsHalf * kCsRec];  // 61 440 B of the 64 KB bank

// the kernel body's platform on the GPU (lpmx_const_stream_body.h)
struct CsDevice {
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int bid() const { return blockIdx.x; }
  __device__ __forceinline__ int n_threads() const { return blockDim.x; }
  __device__ __forceinline__ bool any_sync(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
  __device__ __forceinline__ double src(int i) const { return c_src[i]; }  // warp-uniform index: LDCU, uniform-register operand
  __device__ __forceinline__ double rcp_seed(double d) const {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
  }
};

template <int T>
__global__ void __launch_bounds__(kCsMaxThreads, 1) pair_sum_const_kernel(const CsArgs a) {
  CsDevice pf;
  cs::body<T>(pf, a);
}

// packed 64-byte records -> 48-byte records, zero-padded to whole halves
__global__ void cs_repack_kernel(const double* __restrict__ packed, int n_src_pad, double* __restrict__ out, long n_out) {
  const long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  double v[kCsRec] = {0, 0, 0, 0, 0, 0};
  if (j < n_src_pad) {
    const double2* r = reinterpret_cast<const double2*>(packed + (size_t)kBveRec * j);
    const double2 a = r[0], b = r[1], c = r[2];
    v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y, v[4] = c.x, v[5] = c.y;
  }
#pragma unroll
  for (int k = 0; k < kCsRec; ++k) out[(size_t)kCsRec * j + k] = v[k];
}

typedef void (*cs_kernel_t)(const CsArgs);
cs_kernel_t cs_kernel_for(int T) {
  switch (T) {
    case 4: return pair_sum_const_kernel<4>;
    case 5: return pair_sum_const_kernel<5>;
    case 6: return pair_sum_const_kernel<6>;
    case 7: return pair_sum_const_kernel<7>;
    case 8: return pair_sum_const_kernel<8>;
    default: return nullptr;
  }
}

}  // namespace

// -1 auto (overlapped copies, large launches only), 0 off, 1 overlapped copies, 2 copies on the compute stream
int const_stream_mode(lpmx_handle_t h) {
  if (h->const_stream >= 0) return h->const_stream;
  static const int env = [] {
    const char* e = getenv("LPMX_CONST_STREAM");
    const int v = e ? atoi(e) : -1;
    return (v < -1 || v > 2) ? -1 : v;
  }();
  return env;
}

// The bank is ONE __constant__ array per device (module scope), ordered only by the using handle's streams and events: two
// handles on the same device must not stream through it at once.  The first handle to take the path on a device owns the
// bank until lpmx_destroy; any other handle on that device keeps the ring kernel.
static std::mutex g_bank_mutex;
static lpmx_handle_t g_bank_owner[64] = {};
static bool claim_bank(lpmx_handle_t h) {
  if (h->device < 0 || h->device >= 64) return false;
  std::lock_guard<std::mutex> lk(g_bank_mutex);
  if (!g_bank_owner[h->device]) g_bank_owner[h->device] = h;
  return g_bank_owner[h->device] == h;
}
static void release_bank(lpmx_handle_t h) {
  if (h->device < 0 || h->device >= 64) return;
  std::lock_guard<std::mutex> lk(g_bank_mutex);
  if (g_bank_owner[h->device] == h) g_bank_owner[h->device] = nullptr;
}

// One CTA per SM; a launch takes ~ waves x T x NW while the FP64 pipe is the limit (>= 8 warps).  Least excess wins,
// ties go to T = 6 (the measured shape), then to more warps.  LPMX_CONST_SHAPE="T,NW" overrides (tuning).
void pick_const_shape(int num_sms, int n_tgt, int* T_out, int* nw_out, int* grid_out) {
  int bt = 6, bnw = 8;
  long best = -1;
  static int ft = 0, fnw = 0;
  static bool parsed = false;
  if (!parsed) {
    parsed = true;
    const char* e = getenv("LPMX_CONST_SHAPE");
    if (e && sscanf(e, "%d,%d", &ft, &fnw) == 2 && cs_kernel_for(ft) && fnw >= 1 && fnw * 32 <= kCsMaxThreads) {
    } else {
      ft = fnw = 0;
    }
  }
  if (ft) {
    bt = ft, bnw = fnw;
  } else {
    // warps per CTA in multiples of 4: with 9-11 warps two of the SM's four schedulers carry one warp more and the CTA waits
    // for them (measured: T = 5 with 10 warps 66.4 ms where 8 balanced warps would take 54; profiles/r2b_shape_sweep.txt)
    // measured at icos-8 (profiles/r2e_icos8_const_shapes.txt): T = 5 / 8 warps 1.797e12, T = 6 / 8 warps 1.776e12 interactions/s,
    // but T = 8 / 8 warps 1.649e12 and T = 6 / 12 warps 1.588e12 (below the ring kernel's 1.657e12): 8 warps, T <= 7
    const int order[3] = {5, 6, 7};
    for (int oi = 0; oi < 3; ++oi) {
      const int T = order[oi];
      for (int nw = 8; nw >= 8; nw -= 4) {
        const long tb = (long)T * nw * 32;
        const long ctas = (n_tgt + tb - 1) / tb;
        const long waves = (ctas + num_sms - 1) / num_sms;
        const long cost = waves * T * nw;
        if (best < 0 || cost < best) best = cost, bt = T, bnw = nw;
      }
    }
  }
  const long tb = (long)bt * bnw * 32;
  *T_out = bt;
  *nw_out = bnw;
  *grid_out = (int)((n_tgt + tb - 1) / tb);
}

bool make_const_plan(lpmx_handle_t h, int n_tgt, int n_src, SumPlan* p) {
  const int mode = const_stream_mode(h);
  if (mode == 0) return false;
  // auto: only where a launch (640 sources x all targets of the rank) is long against its fixed costs -- launch gap, target
  // loads, the read-modify-write of the accumulators, wave quantisation: measured break-even ~1e6 targets (file header).
  // forced (1 / 2): wherever one launch can fill the chip at all.
  long min_tgt = mode < 0 ? 1000000L : (long)h->num_sms * 32 * 8 * 5;
  if (const char* e = getenv("LPMX_CONST_MIN_TARGETS")) min_tgt = atol(e);  // parity tests on small meshes
  if ((long)n_tgt < min_tgt || n_tgt < 1 || n_src < 4 * kCsHalf) return false;
  int T, nw, grid;
  pick_const_shape(h->num_sms, n_tgt, &T, &nw, &grid);
  if (mode < 0) {
    // auto: the path is worth +8.5 % of a launch that fills its waves; a rank whose targets leave the last wave mostly empty
    // (1.2e6 targets: 7 waves at 90.6 %) is better served by the stream-K kernel, which has no wave quantisation
    const long tb = (long)T * nw * 32;
    const long waves = (grid + h->num_sms - 1) / h->num_sms;
    if ((double)n_tgt < 0.95 * (double)(waves * h->num_sms * tb)) return false;
  }
  if (!claim_bank(h)) return false;
  p->kind = kVel;
  p->shape = kShapeConstStream;
  p->T = T;
  p->tb = T * nw * 32;
  p->n_tgt = n_tgt;
  p->n_tb = grid;  // CTAs of one launch
  p->n_src_pad = round_up_chunk(n_src);
  p->n_sc = p->n_src_pad / kChunk;
  p->grid = 1;  // what the finalize kernels see: every target block was summed by "CTA 0", i.e. slot 0 only
  p->max_slots = 1;
  p->n_tgt_pad = (long)p->n_tb * p->tb;
  p->smem_bytes = 0;
  return true;
}

int launch_const_stream(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed, double kappa,
                        double* partials) {
  const int mode = const_stream_mode(h) == 2 ? 2 : 1;
  const long n_batches = ((long)p.n_src_pad + kCsHalf - 1) / kCsHalf;
  const long n_out = n_batches * kCsHalf;
  void* stage_v = nullptr;
  LPMX_TRY(dev_buffer(h, "const_stage", sizeof(double) * kCsRec * (size_t)n_out, &stage_v));
  const double* stage = (const double*)stage_v;
  cs_repack_kernel<<<(int)((n_out + 255) / 256), 256, 0, h->stream>>>(packed, p.n_src_pad, (double*)stage_v, n_out);
  ++h->launches;
  LPMX_CUDA(h, cudaGetLastError());
  if (!h->cs_events[0]) {
    for (int i = 0; i < 5; ++i) LPMX_CUDA(h, cudaEventCreateWithFlags(&h->cs_events[i], cudaEventDisableTiming));
  }
  cudaEvent_t ev_repack = h->cs_events[0];
  cudaEvent_t* ev_copied = &h->cs_events[1];  // [half]
  cudaEvent_t* ev_summed = &h->cs_events[3];  // [half]
  const size_t half_bytes = sizeof(double) * kCsRec * kCsHalf;
  cs_kernel_t kern = cs_kernel_for(p.T);
  if (!kern) return set_error(h, LPMX_ERR_STATE, "no constant-bank kernel for T = %d", p.T);
  const int threads = p.tb / p.T;
  cudaStream_t cps = mode == 1 ? h->copy_stream : h->stream;
  // copy of batch b into half b & 1; in the overlapped mode it waits for the launch that last read that half
  auto copy_batch = [&](long b) -> int {
    const int half = (int)(b & 1);
    if (mode == 1 && b >= 2) LPMX_CUDA(h, cudaStreamWaitEvent(cps, ev_summed[half], 0));
    LPMX_CUDA(h, cudaMemcpyToSymbolAsync(c_src, stage + (size_t)b * kCsHalf * kCsRec, half_bytes, (size_t)half * half_bytes,
                                         cudaMemcpyDeviceToDevice, cps));
    if (mode == 1) LPMX_CUDA(h, cudaEventRecord(ev_copied[half], cps));
    return LPMX_OK;
  };
  if (mode == 1) {
    // everything queued so far on the compute stream (the repack, and any earlier launch still reading the bank)
    LPMX_CUDA(h, cudaEventRecord(ev_repack, h->stream));
    LPMX_CUDA(h, cudaStreamWaitEvent(cps, ev_repack, 0));
    LPMX_TRY(copy_batch(0));
    if (n_batches > 1) LPMX_TRY(copy_batch(1));
  }
  CsArgs a;
  a.tgt = tgt.p;
  a.tgt_si = tgt.si;
  a.tgt_sk = tgt.sk;
  a.self_idx = self_idx;
  a.acc = partials;
  a.n_tgt_pad = p.n_tgt_pad;
  a.n_tgt = p.n_tgt;
  a.kappa = kappa;
  for (long b = 0; b < n_batches; ++b) {
    const int half = (int)(b & 1);
    if (mode == 1)
      LPMX_CUDA(h, cudaStreamWaitEvent(h->stream, ev_copied[half], 0));
    else
      LPMX_TRY(copy_batch(b));
    a.half = half;
    a.j0 = (int)(b * kCsHalf);
    a.first = b == 0 ? 1 : 0;
    kern<<<p.n_tb, threads, 0, h->stream>>>(a);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
    if (mode == 1) {
      LPMX_CUDA(h, cudaEventRecord(ev_summed[half], h->stream));
      if (b + 2 < n_batches) LPMX_TRY(copy_batch(b + 2));
    }
  }
  return LPMX_OK;
}

void const_stream_teardown(lpmx_handle_t h) {
  release_bank(h);
  for (int i = 0; i < 5; ++i)
    if (h->cs_events[i]) {
      cudaEventDestroy(h->cs_events[i]);
      h->cs_events[i] = nullptr;
    }
}

}  // namespace lpmx

extern "C" int lpmx_const_stream_shape(int num_sms, int n_tgt, int* T, int* n_warps, int* grid) {
  if (num_sms < 1 || n_tgt < 1 || !T || !n_warps || !grid) return LPMX_ERR_INVALID;
  lpmx::pick_const_shape(num_sms, n_tgt, T, n_warps, grid);
  return LPMX_OK;
}

extern "C" int lpmx_pair_sum_const_stream(lpmx_handle_t h, int mode) {
  if (!h || mode < -1 || mode > 2) return LPMX_ERR_INVALID;
  h->const_stream = mode;
  if (mode == 0) lpmx::release_bank(h);
  return LPMX_OK;
}
