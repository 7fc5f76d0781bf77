// lpmx_const_stream.cu -- the velocity pair sum with the source records streamed through CONSTANT banks, so that they reach
// the DFMAs as uniform-register operands (SASS: LDCU.64 + DFMA R, R, UR, R).  Host side: planning, the launch sequence of one
// evaluation, its CUDA graph.  Device side: lpmx_const_bank.cuh (one bank + its kernels, compiled once per bank) and
// lpmx_const_stream_body.h (the kernel body, shared with a host model).  DESIGN.md section 4.1b has the measurements.
//
// Why: pair_sum_kernel (lpmx_pair_kernel.cuh) is bound by FP64 issue and reaches 80 % of the pipe because 5 of its 9 DFMAs
// per pair read a third distinct register operand (profiles/README.md, "Why 81 %").  A source record is the same for every
// thread; read from c[3][..] it costs no register-file port: 94-96 % of the FP64 peak in issued DFMAs over whole evaluations.
//
// How:
//   * BANKS ARE MODULES.  lpmx_const_bank.cu is compiled kCsBanks = 24 times; every object is its own module with its own
//     64 KB user constant bank of cs::kBatch = 1 280 records x 48 B {y0, y1, y2, G*y0, G*y1, G*y2}.  A launch of bank b's
//     kernel sums that bank into all targets of the launch (accumulators live in slot 0 of the partials buffer between
//     launches, added with RED.ADD.F64); device-to-device copies on the handle's cs_stream refill a bank once the launch
//     that read it has completed.
//   * PIPELINED LAUNCHES (programmatic dependent launch, LPMX_CONST_PDL, default on).  CTAs of launch b + 1 take the slots
//     the CTAs of launch b leave and sum their own bank beside them; only their reduction into the accumulators waits for
//     launch b, so the additions per target keep the launch order (bit-identical to the serial sequence).  No waves to fill,
//     no remainder: every target goes through the banks in CTAs of 6 x 128 targets + a prefetch warp, three per SM, and as
//     many launches are in flight as it takes to fill the chip's 444 CTA slots -- one per bank in rotation, which is why a
//     rank's share of a small mesh (38 CTAs per launch at cubed-7 on eight GPUs) needs a dozen banks and more.
//   * ONE GRAPH LAUNCH PER EVALUATION: the ~80 (160) launches, as many refills and their event edges are captured the second
//     time a sequence comes by and replayed from then on (LPMX_CONST_GRAPH).
//   * WITHOUT THE PIPELINING (LPMX_CONST_PDL=0) all CTAs of a launch start together, so the launch has to fill whole waves:
//     pick_const_split then takes waves x 148 CTAs of 8 warps through two banks and hands the remainder to the ring kernel,
//     whose slots cs_fold_kernel adds into slot 0 (the state of r2s; kept, and tested, as the fallback).
// LPMX_CONST_STREAM=0 turns the path off, =1 forces it wherever a launch exists (refills overlapped), =2 forces it with the
// refills on the compute stream; lpmx_pair_sum_const_stream() sets the same per handle.
//
// Same arithmetic per pair as Pair<kVel> (bit-identical terms); per target the terms are added in source order in blocks of
// 1 280 (640), so the sums differ from the stream-K kernel's by round-off only (the tolerance of every parity test covers both).
#include <cstdlib>
#include <mutex>
#include <vector>

#include "lpmx_const_stream_body.h"
#include "lpmx_internal.h"

namespace lpmx {

// lpmx_kernels.cu
int launch_ring_remainder(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed, double kappa,
                          double* rem_partials, const int* tgt_map);

namespace {

using cs::CsArgs;
constexpr int kCsBatch = cs::kBatch;  // records per bank
constexpr int kCsRec = cs::kRec;      // doubles per record

struct Bank {
  cudaError_t (*launch)(int, int, int, cudaStream_t, const CsArgs&, int);
  cudaError_t (*fill)(const double*, int, cudaStream_t);
};
#define LPMX_CS_ENTRY(k) {cs_bank_launch_##k, cs_bank_fill_##k},
const Bank kBanks[kCsBanks] = {LPMX_CS_BANK_LIST(LPMX_CS_ENTRY)};
#undef LPMX_CS_ENTRY

// packed 64-byte records -> 48-byte records, zero-padded to whole banks
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

// remainder targets: the ring kernel's slots (in slot order, as the stage kernels would add them) -> the bank path's layout
__global__ void cs_fold_kernel(const double* __restrict__ rem, long rem_pad, int n_rem, int rem_tb, int rem_n_sc, int rem_grid,
                               long rem_items, double* __restrict__ acc, long n_tgt_pad, int n_const) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rem) return;
  const int tb = i / rem_tb;
  const int c0 = (int)((((long)tb * rem_n_sc + 1) * (long)rem_grid - 1) / rem_items);
  const int c1 = (int)(((((long)(tb + 1) * rem_n_sc - 1) + 1) * (long)rem_grid - 1) / rem_items);
  double m[3] = {0.0, 0.0, 0.0};
  for (int slot = 0; slot <= c1 - c0; ++slot)
#pragma unroll
    for (int k = 0; k < 3; ++k) m[k] += rem[((long)slot * 3 + k) * rem_pad + i];
#pragma unroll
  for (int k = 0; k < 3; ++k) acc[(long)k * n_tgt_pad + n_const + i] = m[k];
}

}  // namespace

// -1 auto (overlapped refills, where the planner's model beats the ring kernel), 0 off, 1 forced with overlapped refills,
// 2 forced with the refills on the compute stream
int const_stream_mode(lpmx_handle_t h) {
  if (h->const_stream >= 0) return h->const_stream;
  static const int env = [] {
    const char* e = getenv("LPMX_CONST_STREAM");
    const int v = e ? atoi(e) : -1;
    return (v < -1 || v > 2) ? -1 : v;
  }();
  return env;
}

// The banks are module-scope __constant__ arrays, one set per device, ordered only by the using handle's streams and
// events: two handles on the same device must not stream through them at once.  The first handle to take the path on a device
// owns the banks until lpmx_destroy; any other handle on that device keeps the ring kernel.
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

// Measured constants of the model (one B200): a bank launch with every CTA full runs at 88.6 % of the FP64 pipe
// (profiles/r2e_icos8_const_shapes.txt: T = 5 and 6 within 1 %, T = 7 slightly below), i.e. 64 / 9 * 0.886 = 6.3 pairs per
// cycle and SM; its fixed cost (launch gap, target loads, the read-modify-write of the accumulators) is ~6 us
// (profiles/r2b_shape_sweep.txt: two waves of T = 6 per 640-record launch took 163.8 us at cubed-7).
static double const_launch_seconds(int waves, int T, int nw) {
  const double pairs_per_sm = (double)waves * T * nw * 32 * kCsBatch;
  const double eff = T == 7 ? 0.80 : 0.886;  // T = 7: one CTA per SM, no second CTA to cover its stalls (r2e, r2r)
  return pairs_per_sm * 9.0 / (64.0 * eff) / 1.965e9 + 6e-6;
}

// LPMX_CONST_SHAPE="T,NW[,PERSM]" pins the shape (tuning): T targets per thread, NW compute warps, PERSM CTAs per SM and wave
static void forced_shape(int* ft, int* fnw, int* fps) {  // parsed per call: tools switch shapes within one process
  int t = 0, nw = 0, ps = 1;
  const char* e = getenv("LPMX_CONST_SHAPE");
  const int n = e ? sscanf(e, "%d,%d,%d", &t, &nw, &ps) : 0;
  if (n < 3) ps = 1;
  if (!(n >= 2 && t >= 3 && t <= 8 && nw >= 1 && (nw + 1) * 32 <= kCsMaxThreads && ps >= 1 && ps <= 5)) t = nw = 0, ps = 1;
  *ft = t, *fnw = nw, *fps = ps;
}

// source records per bank launch: a whole bank, or half a bank for small target sets (measured, r2z: 28 672 targets 1.667 ->
// 1.62 ms, 57 344 targets 3.18 -> 3.07 ms per evaluation; no difference from 114 688 targets up).  LPMX_CONST_BATCH overrides.
constexpr int kCsHalfBatchBelow = 100000;
int const_batch(int n_tgt) {
  int b = n_tgt < kCsHalfBatchBelow ? kCsBatch / 2 : kCsBatch;
  if (const char* e = getenv("LPMX_CONST_BATCH")) b = atoi(e) <= kCsBatch / 2 ? kCsBatch / 2 : kCsBatch;
  return b;
}

// LPMX_CONST_PDL=0 switches the pipelining of the bank launches off (then: whole waves + a ring remainder, see below)
bool const_pdl() {
  const char* e = getenv("LPMX_CONST_PDL");
  return !(e && atoi(e) == 0);
}

double pick_const_split(lpmx_handle_t h, int num_sms, int n_tgt, int n_src, int* T_out, int* nw_out, int* ctas_out, int* n_const_out,
                        double* ring_s_out) {
  const long n_batches = ((long)round_up_chunk(n_src) + kCsBatch - 1) / kCsBatch;  // (the whole-wave planner below)
  SumPlan ring;
  double ring_s = 0.0;
  if (make_plan(h, kVel, n_tgt, n_src, &ring, false) == LPMX_OK) ring_s = ring_plan_seconds(ring);
  if (ring_s_out) *ring_s_out = ring_s;
  int ft, fnw, fps;
  forced_shape(&ft, &fnw, &fps);
  if (const_pdl()) {
    // Pipelined launches (programmatic dependent launch): CTAs of launch b + 1 take the slots the CTAs of launch b leave, so
    // there are no waves to fill and no remainder -- every target goes through the banks.  Shape: T = 6 targets per thread,
    // 4 compute warps + the prefetch warp, three CTAs per SM (126 registers).  Measured on one B200 (r2y,
    // tools/rank_size_sweep.py: 28 672 .. 229 376 targets x 98 304 sources): t = pairs / 2.0e12 + 0.26 ms with 1 280 records
    // per launch -- the FP64 pipe is 97 % busy while the pipeline is full, and filling and draining it costs about the time a
    // CTA holds its slot (three CTAs share an SM: 3 x 768 x 1 280 pairs = 216 us) once per evaluation.  Enough launches have
    // to be in flight to fill the chip's 444 CTA slots: one per bank, so small launches (a rank's share of a small mesh: 38
    // CTAs at cubed-7 on eight GPUs) are limited by the number of banks.
    const int T = ft ? ft : 6, nw = ft ? fnw : 4;
    const int per_sm = ft ? fps : 3;
    const long tb = (long)T * nw * 32;
    const long ctas = (n_tgt + tb - 1) / tb;
    *T_out = T, *nw_out = nw, *ctas_out = (int)ctas, *n_const_out = n_tgt;
    const int batch = (T == 6 && nw == 4) ? const_batch(n_tgt) : kCsBatch;
    const double slots = (double)num_sms * per_sm;
    double fill = (double)ctas * kCsBanks / slots;
    if (fill > 1.0) fill = 1.0;
    const double sum_s = (double)ctas * tb * (double)round_up_chunk(n_src) / (2.0e12 * fill);
    const double drain_s = (double)per_sm * tb * batch * 9.0 / (64.0 * 0.97 * 1.965e9);
    return sum_s + drain_s + 30e-6;
  }
  double best = -1.0;
  // 8 warps = 2 per scheduler: with 9-11 warps two of the SM's four schedulers carry one warp more and the CTA waits for them
  // (r2b: T = 5 with 10 warps 66.4 ms at cubed-7), 12 warps and T = 8 are slower than the ring kernel (r2e)
  for (int T = ft ? ft : 5; T <= (ft ? ft : 7); ++T) {
    const int nw = ft ? fnw : 8;
    const long tb = (long)T * nw * 32;
    const long wave = (long)num_sms * (ft ? fps : 1) * tb;
    const long full = n_tgt / wave;
    // candidates: `full` whole waves + a ring remainder, or one more (partly empty) wave and no remainder
    for (int extra = 0; extra < 2; ++extra) {
      const long waves = full + extra;
      if (waves < 1) continue;
      const long n_const = extra ? n_tgt : full * wave;
      const long ctas = (n_const + tb - 1) / tb;
      double t = (double)n_batches * const_launch_seconds((int)waves, T, nw);
      const long n_rem = n_tgt - n_const;
      if (n_rem > 0) {
        SumPlan r;
        if (make_best_ring_plan(h, (int)n_rem, n_src, &r) != LPMX_OK) continue;
        t += ring_plan_seconds(r) + 4e-6;  // + the fold
      }
      if (best < 0 || t < best) {
        best = t;
        *T_out = T, *nw_out = nw, *ctas_out = (int)ctas, *n_const_out = (int)n_const;
      }
    }
  }
  return best;
}

bool make_const_plan(lpmx_handle_t h, int n_tgt, int n_src, SumPlan* p) {
  const int mode = const_stream_mode(h);
  if (mode == 0) return false;
  // auto: from one full wave of the smallest shape upwards, and only where the modelled time beats the ring kernel's;
  // forced (1 / 2): wherever one wave can be filled at all.  LPMX_CONST_MIN_TARGETS lowers the floor (parity tests on small
  // meshes: whatever does not fill a wave goes through an only partly filled one)
  // pipelined launches fill the chip with the CTAs of several launches, so the floor is low and the model decides
  long min_tgt = const_pdl() ? 16384 : (long)h->num_sms * 32 * 8 * 5;
  if (const char* e = getenv("LPMX_CONST_MIN_TARGETS")) min_tgt = atol(e);
  if ((long)n_tgt < min_tgt || n_tgt < 1 || n_src < 4 * kCsBatch) return false;
  int T = 0, nw = 0, ctas = 0, n_const = 0;
  double ring_s = 0.0;
  const double t = pick_const_split(h, h->num_sms, n_tgt, n_src, &T, &nw, &ctas, &n_const, &ring_s);
  if (t < 0) return false;
  if (mode < 0 && !(t < 0.98 * ring_s)) return false;
  if (!claim_bank(h)) return false;
  *p = SumPlan();
  p->kind = kVel;
  p->shape = kShapeConstStream;
  p->T = T;
  p->tb = T * nw * 32;
  p->n_tgt = n_tgt;
  p->cs_ctas = ctas;
  p->cs_n_const = n_const;
  p->cs_batch = (const_pdl() && T == 6 && nw == 4) ? const_batch(n_tgt) : kCsBatch;
  p->n_src_pad = round_up_chunk(n_src);
  p->n_sc = p->n_src_pad / kChunk;
  p->grid = 1;  // what the finalize kernels see: every target block was summed by "CTA 0", i.e. slot 0 only
  p->max_slots = 1;
  const long covered = (long)ctas * p->tb;  // >= n_const; == n_const when a remainder follows
  p->n_tgt_pad = covered > n_tgt ? covered : (long)n_tgt;
  p->n_tb = (int)((p->n_tgt_pad + p->tb - 1) / p->tb);
  p->smem_bytes = 0;
  if (n_const < n_tgt) {
    SumPlan r;
    if (make_best_ring_plan(h, n_tgt - n_const, n_src, &r) != LPMX_OK) return false;
    p->rem.shape = r.shape, p->rem.T = r.T, p->rem.tb = r.tb, p->rem.n_tgt = r.n_tgt, p->rem.n_tb = r.n_tb, p->rem.grid = r.grid;
    p->rem.max_slots = r.max_slots, p->rem.n_tgt_pad = r.n_tgt_pad, p->rem.smem_bytes = r.smem_bytes;
  }
  return true;
}

// the launch sequence of one evaluation, enqueued on the handle's two streams (eagerly, or into a stream capture)
static int enqueue_const_stream(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed,
                                double kappa, double* partials, const int* tgt_map, int mode, int pf_stride, void* stage_v,
                                long* n_launches, long* n_bank_launches) {
  const int batch = p.cs_batch;
  const long n_batches = ((long)p.n_src_pad + batch - 1) / batch;
  const long n_out = n_batches * batch;
  const double* stage = (const double*)stage_v;
  const long launches0 = h->launches, cs0 = h->cs_launches;
  cs_repack_kernel<<<(int)((n_out + 255) / 256), 256, 0, h->stream>>>(packed, p.n_src_pad, (double*)stage_v, n_out);
  ++h->launches;
  LPMX_CUDA(h, cudaGetLastError());
  if (h->cs_events.empty()) {
    h->cs_events.assign(1 + 2 * kCsBanks, nullptr);
    for (auto& e : h->cs_events) LPMX_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaEvent_t ev_repack = h->cs_events[0];
  cudaEvent_t* ev_filled = &h->cs_events[1];             // [bank]
  cudaEvent_t* ev_summed = &h->cs_events[1 + kCsBanks];  // [bank]
  // the prefetch warp (lpmx_const_bank.cuh) rides along on launches of a single wave, where every CTA starts on cold constant
  // caches; later waves of a longer launch find the bank cached, and without the extra warp two CTAs of T <= 6 fit an SM
  // (r2p: 88.6 % of the FP64 pipe at icos-8 that way, 78 % with the extra warp in every CTA, r2q)
  int fts, fnws, per_sm;
  forced_shape(&fts, &fnws, &per_sm);
  const bool single_wave = p.cs_ctas <= h->num_sms * (fts ? per_sm : 1);
  const bool pdl = const_pdl();  // bank launch b + 1 may start while launch b runs (programmatic dependent launch)
  // the pipelined shapes: their kernel instances carry the prefetch warp and exist in every bank
  const bool small_cta = p.T == 6 && p.tb == 6 * 4 * 32;
  const int pf = (single_wave || (pdl && small_cta)) ? pf_stride : 0;
  const int threads = p.tb / p.T + (pf > 0 ? 32 : 0);
  // banks in rotation: as many launches can be in flight at once.  Two for the large shapes (their launches fill the chip
  // by themselves); for the pipelined small-CTA shapes enough of them to fill every CTA slot of the chip with the CTAs of
  // consecutive launches (a rank's share of a small mesh is a few dozen CTAs per launch), LPMX_CONST_BANKS caps it.
  int nb = 2;
  if (pdl && small_cta && pf > 0 && mode == 1) {
    const int slots = h->num_sms * 3;
    nb = (slots + p.cs_ctas - 1) / p.cs_ctas + 2;
    if (const char* e = getenv("LPMX_CONST_BANKS")) nb = atoi(e);
    if (nb < 2) nb = 2;
    if (nb > kCsBanks) nb = kCsBanks;
  }
  cudaStream_t cps = mode == 1 ? h->cs_stream : h->stream;
  // refill of bank (b % nb) with batch b; in the overlapped mode it waits for the launch that last read that bank
  auto fill_batch = [&](long b) -> int {
    const int bank = (int)(b % nb);
    if (mode == 1 && b >= nb) LPMX_CUDA(h, cudaStreamWaitEvent(cps, ev_summed[bank], 0));
    LPMX_CUDA(h, kBanks[bank].fill(stage + (size_t)b * batch * kCsRec, batch, cps));
    if (mode == 1) LPMX_CUDA(h, cudaEventRecord(ev_filled[bank], cps));
    return LPMX_OK;
  };
  if (mode == 1) {
    // everything queued so far on the compute stream (the repack, and any earlier launch still reading the banks)
    LPMX_CUDA(h, cudaEventRecord(ev_repack, h->stream));
    LPMX_CUDA(h, cudaStreamWaitEvent(cps, ev_repack, 0));
    for (long b = 0; b < nb && b < n_batches; ++b) LPMX_TRY(fill_batch(b));
  }
  // the remainder first: the ring kernel runs while the first banks are being filled
  double* rem_partials = partials + 3 * (size_t)p.n_tgt_pad;
  if (p.rem.n_tgt > 0) LPMX_TRY(launch_ring_remainder(h, p, tgt, self_idx, packed, kappa, rem_partials, tgt_map));
  CsArgs a;
  a.tgt = tgt.p;
  a.tgt_si = tgt.si;
  a.tgt_sk = tgt.sk;
  a.tgt_map = tgt_map;
  a.self_idx = self_idx;
  a.acc = partials;
  a.n_tgt_pad = p.n_tgt_pad;
  a.n_tgt = p.cs_n_const;
  a.kappa = kappa;
  a.prefetch_stride = pf;
  for (long b = 0; b < n_batches; ++b) {
    const int bank = (int)(b % nb);
    if (mode == 1)
      LPMX_CUDA(h, cudaStreamWaitEvent(h->stream, ev_filled[bank], 0));
    else
      LPMX_TRY(fill_batch(b));
    a.j0 = (int)(b * batch);
    a.n_rec = batch;
    a.first = b == 0 ? 1 : 0;
    const cudaError_t e = kBanks[bank].launch(p.T, p.cs_ctas, threads, h->stream, a, (pdl && b > 0) ? 1 : 0);
    if (e == cudaErrorInvalidValue) return set_error(h, LPMX_ERR_STATE, "no constant-bank kernel for T = %d", p.T);
    ++h->launches;
    ++h->cs_launches;
    LPMX_CUDA(h, e);
    if (mode == 1) {
      LPMX_CUDA(h, cudaEventRecord(ev_summed[bank], h->stream));
      if (b + nb < n_batches) LPMX_TRY(fill_batch(b + nb));
    }
  }
  if (p.rem.n_tgt > 0) {
    const long rem_items = (long)p.rem.n_tb * p.n_sc;
    cs_fold_kernel<<<(p.rem.n_tgt + 255) / 256, 256, 0, h->stream>>>(rem_partials, p.rem.n_tgt_pad, p.rem.n_tgt, p.rem.tb, p.n_sc,
                                                                      p.rem.grid, rem_items, partials, p.n_tgt_pad, p.cs_n_const);
    ++h->launches;
    LPMX_CUDA(h, cudaGetLastError());
  }
  *n_launches = h->launches - launches0;
  *n_bank_launches = h->cs_launches - cs0;
  return LPMX_OK;
}

// One evaluation is ~80 kernel launches, as many bank refills and twice as many event operations: enqueued one by one they
// cost the host ~15 ms per BVERK4 step at cubed-7 (r2p: e2e 81 ms against 66 ms on the device).  The sequence depends only on
// the plan and the pointers, and a stepper repeats the same few of them (two source buffers x the work array), so the second
// time a sequence comes by it is captured into a CUDA graph -- both streams, the event edges, the refills as memcpy nodes --
// and from then on it is ONE cudaGraphLaunch.  LPMX_CONST_GRAPH=0 keeps the eager sequence.
namespace {
struct CsGraphKey {
  const void *tgt, *self_idx, *packed, *partials, *stage, *tgt_map;
  long si, sk, n_tgt_pad, rem_pad;
  int T, tb, ctas, n_const, n_tgt, n_src_pad, rem_n, rem_shape, rem_grid, mode, pf, pdl, banks, batch;
  double kappa;
  bool operator==(const CsGraphKey& o) const {
    return tgt == o.tgt && self_idx == o.self_idx && packed == o.packed && partials == o.partials && stage == o.stage && tgt_map == o.tgt_map && si == o.si &&
           sk == o.sk && n_tgt_pad == o.n_tgt_pad && rem_pad == o.rem_pad && T == o.T && tb == o.tb && ctas == o.ctas &&
           n_const == o.n_const && n_tgt == o.n_tgt && n_src_pad == o.n_src_pad && rem_n == o.rem_n && rem_shape == o.rem_shape &&
           rem_grid == o.rem_grid && mode == o.mode && pf == o.pf && pdl == o.pdl && banks == o.banks && batch == o.batch && kappa == o.kappa;
  }
};
struct CsGraph {
  CsGraphKey key;
  cudaGraphExec_t exec = nullptr;
  long n_launches = 0, n_bank_launches = 0;
  unsigned long tick = 0;
};
struct CsGraphCache {
  std::vector<CsGraph> entries;
  unsigned long tick = 0;
  bool broken = false;  // a capture failed on this handle: stay eager
};
constexpr size_t kCsMaxGraphs = 16;
}  // namespace

int launch_const_stream(lpmx_handle_t h, const SumPlan& p, Vec3View tgt, const int* self_idx, const double* packed, double kappa,
                        double* partials, const int* tgt_map) {
  const int mode = const_stream_mode(h) == 2 ? 2 : 1;
  if (mode == 1 && !h->cs_stream) LPMX_CUDA(h, cudaStreamCreateWithFlags(&h->cs_stream, cudaStreamNonBlocking));
  // read per call (two getenv per evaluation) so that tests can switch them within one process
  const int pf_stride = [] {  // LPMX_CONST_PREFETCH = bytes between the prefetch warp's loads (0: off); default one per 128 B (r2r: 64 B 51.9 ms, 128 B 51.4 ms per step at cubed-7)
    const char* e = getenv("LPMX_CONST_PREFETCH");
    const int v = e ? atoi(e) : 128;
    return v <= 0 ? 0 : (v < 8 ? 1 : v / 8);
  }();
  const bool use_graphs = [] {
    const char* e = getenv("LPMX_CONST_GRAPH");
    return !(e && atoi(e) == 0);
  }();
  const long n_batches = ((long)p.n_src_pad + p.cs_batch - 1) / p.cs_batch;
  void* stage_v = nullptr;
  LPMX_TRY(dev_buffer(h, "const_stage", sizeof(double) * kCsRec * (size_t)(n_batches * p.cs_batch), &stage_v));
  long nl = 0, nb = 0;
  if (!use_graphs) return enqueue_const_stream(h, p, tgt, self_idx, packed, kappa, partials, tgt_map, mode, pf_stride, stage_v, &nl, &nb);
  if (!h->cs_graph_cache) h->cs_graph_cache = new CsGraphCache();
  CsGraphCache* cache = (CsGraphCache*)h->cs_graph_cache;
  if (cache->broken) return enqueue_const_stream(h, p, tgt, self_idx, packed, kappa, partials, tgt_map, mode, pf_stride, stage_v, &nl, &nb);
  CsGraphKey key{tgt.p, self_idx, packed, partials, stage_v, tgt_map, tgt.si, tgt.sk, p.n_tgt_pad, p.rem.n_tgt_pad, p.T, p.tb, p.cs_ctas,
                 p.cs_n_const, p.n_tgt, p.n_src_pad, p.rem.n_tgt, p.rem.shape, p.rem.grid, mode, pf_stride, const_pdl() ? 1 : 0, getenv("LPMX_CONST_BANKS") ? atoi(getenv("LPMX_CONST_BANKS")) : 0, p.cs_batch, kappa};
  CsGraph* g = nullptr;
  for (auto& e : cache->entries)
    if (e.key == key) g = &e;
  if (!g) {
    // first sighting: run it eagerly (this also does the one-time work a capture must not contain: function attributes, event
    // creation, buffer growth) and remember the key
    if (cache->entries.size() >= kCsMaxGraphs) {
      size_t old = 0;
      for (size_t i = 1; i < cache->entries.size(); ++i)
        if (cache->entries[i].tick < cache->entries[old].tick) old = i;
      if (cache->entries[old].exec) cudaGraphExecDestroy(cache->entries[old].exec);
      cache->entries.erase(cache->entries.begin() + old);
    }
    CsGraph e;
    e.key = key;
    e.tick = ++cache->tick;
    cache->entries.push_back(e);
    return enqueue_const_stream(h, p, tgt, self_idx, packed, kappa, partials, tgt_map, mode, pf_stride, stage_v, &nl, &nb);
  }
  g->tick = ++cache->tick;
  if (!g->exec) {
    const long launches0 = h->launches, cs0 = h->cs_launches;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    int rc = LPMX_OK;
    if (ce == cudaSuccess) {
      rc = enqueue_const_stream(h, p, tgt, self_idx, packed, kappa, partials, tgt_map, mode, pf_stride, stage_v, &g->n_launches, &g->n_bank_launches);
      ce = cudaStreamEndCapture(h->stream, &graph);
    }
    h->launches = launches0, h->cs_launches = cs0;  // nothing has run yet
    if (ce == cudaSuccess && rc == LPMX_OK && graph) ce = cudaGraphInstantiate(&g->exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ce != cudaSuccess || rc != LPMX_OK || !g->exec) {
      // not capturable here: clear the error state and stay with the eager sequence on this handle
      cudaGetLastError();
      h->err.clear();
      g->exec = nullptr;
      cache->broken = true;
      return enqueue_const_stream(h, p, tgt, self_idx, packed, kappa, partials, tgt_map, mode, pf_stride, stage_v, &nl, &nb);
    }
  }
  LPMX_CUDA(h, cudaGraphLaunch(g->exec, h->stream));
  h->launches += g->n_launches;
  h->cs_launches += g->n_bank_launches;
  return LPMX_OK;
}

void const_stream_teardown(lpmx_handle_t h) {
  release_bank(h);
  if (h->cs_graph_cache) {
    CsGraphCache* cache = (CsGraphCache*)h->cs_graph_cache;
    for (auto& e : cache->entries)
      if (e.exec) cudaGraphExecDestroy(e.exec);
    delete cache;
    h->cs_graph_cache = nullptr;
  }
  if (h->cs_stream) {
    cudaStreamDestroy(h->cs_stream);
    h->cs_stream = nullptr;
  }
  for (auto& e : h->cs_events)
    if (e) cudaEventDestroy(e);
  h->cs_events.clear();
}

}  // namespace lpmx

extern "C" int lpmx_const_stream_split(int num_sms, int n_tgt, int n_src, int* T, int* n_warps, int* ctas, int* n_const,
                                       double* model_seconds, double* ring_seconds) {
  if (num_sms < 1 || n_tgt < 1 || n_src < 1 || !T || !n_warps || !ctas || !n_const) return LPMX_ERR_INVALID;
  lpmx_handle_s probe;  // planning only: the ring planner reads num_sms and nothing else
  probe.num_sms = num_sms;
  probe.device = -1;
  double ring_s = 0.0;
  const double t = lpmx::pick_const_split(&probe, num_sms, n_tgt, n_src, T, n_warps, ctas, n_const, &ring_s);
  if (t < 0) return LPMX_ERR_STATE;
  if (model_seconds) *model_seconds = t;
  if (ring_seconds) *ring_seconds = ring_s;
  return LPMX_OK;
}

extern "C" int lpmx_const_stream_launch_count(lpmx_handle_t h, long* n) {
  if (!h || !n) return LPMX_ERR_INVALID;
  *n = h->cs_launches;
  return LPMX_OK;
}

extern "C" int lpmx_pair_sum_const_stream(lpmx_handle_t h, int mode) {
  if (!h || mode < -1 || mode > 2) return LPMX_ERR_INVALID;
  h->const_stream = mode;
  if (mode == 0) lpmx::release_bank(h);
  return LPMX_OK;
}
