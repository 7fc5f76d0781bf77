// lpmx_const_bank.cuh -- ONE constant bank of the constant-bank velocity path and the kernels that read it.
// Compiled kCsBanks times (lpmx_const_bank.cu with -DLPMX_CS_BANK=0 .. kCsBanks-1, lpm_b200/build.py): without relocatable
// device code every translation unit is its own module with its own 64 KB user constant bank, so N translation units give N
// banks of cs::kBatch = 1 280 records.  A launch of bank b's kernel sums the whole bank into its targets while the other banks
// are refilled behind it or read by the launches pipelined around it (lpmx_const_stream.cu).  Banks 0 and 1 carry every kernel
// shape; the others only the pipelined small-CTA shapes, the only ones that keep more than two launches in flight.
#ifndef LPMX_CS_BANK
#error "compile lpmx_const_bank.cu with -DLPMX_CS_BANK=<n>"
#endif

#include "lpmx_const_stream_body.h"
#include "lpmx_internal.h"

namespace lpmx {

namespace {

__constant__ double c_src[cs::kBankDoubles];  // 61 488 B of this module's 64 KB bank

// the kernel body's platform on the GPU (lpmx_const_stream_body.h)
struct CsDevice {
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int bid() const { return blockIdx.x; }
  int lanes;  // compute threads of the CTA (the launch may carry one more warp: the prefetcher)
  __device__ __forceinline__ int n_threads() const { return lanes; }
  __device__ __forceinline__ bool any_sync(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
  __device__ __forceinline__ double src(int i) const { return c_src[i]; }  // warp-uniform index: LDCU, uniform-register operand
  __device__ __forceinline__ double rcp_seed(double d) const {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
  }
  // *p += v as a reduction that returns nothing (RED.E.ADD.F64): the CTA does not wait for the old value on its way out.  One
  // thread per address and launch, launches in stream order: the same IEEE sum as load-add-store, deterministic.
  // Programmatic dependent launch (LPMX_CONST_PDL): the next bank launch may start while this one runs -- its CTAs load their
  // targets and sum their bank beside ours -- and only its reduction into the accumulators waits for us (wait_prior), so the
  // per-target order of the additions stays the launch order.  Both are no-ops in a launch without the attribute.
  __device__ __forceinline__ void launch_dependents() const { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  __device__ __forceinline__ void wait_prior() const { asm volatile("griddepcontrol.wait;" ::: "memory"); }
  __device__ __forceinline__ void accumulate(double* p, double v, bool first) const {
    if (first)
      *p = v;
    else
      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  }
};

// The last warp of the CTA walks the bank ahead of the compute warps, one load per kCsLineBytes.  Why: after a launch the
// SM's constant caches are cold, and the compute warps read the bank in step -- every new line is a miss that all of them
// wait for, one miss at a time (r2p: 56 % of the FP64 pipe on the first wave of a launch against 88 % on the later waves
// of the same launch, whose lines are still cached; ~1.6 ns per byte = 96 us per 61 KB bank and SM).  The prefetch warp takes
// those misses instead, ahead of the compute warps, and exits.  (CsArgs::prefetch_stride = 0 switches it off: LPMX_CONST_PREFETCH.)
__device__ __forceinline__ void prefetch_bank(int stride, int n_rec, double* never) {  // stride in doubles: one load per cache line
  const int lane = threadIdx.x & 31;
  double sink = 0.0;
#pragma unroll 4
  for (int i = lane * stride; i < n_rec * cs::kRec; i += 32 * stride) sink += c_src[i];  // lane-varying index: LDC
  if (sink == -1.2345678901234567e300) *never = sink;  // keeps the loads alive (ptxas drops loads nobody consumes)
}

// compute warps the register budget of one CTA allows (+ the prefetch warp)
__host__ __device__ constexpr int cs_max_threads(int T) { return T >= 5 ? 288 : T == 4 ? 416 : 544; }

// PF: launched with one extra warp that prefetches (single-wave launches).  Without it (launches of several waves) T <= 6 is
// held to 128 registers so that two CTAs share an SM: 16 resident warps and no gap between waves (r2p: 88.6 % of the pipe).
// SMALL: the pipelined shape -- 4 compute warps + the prefetch warp, three CTAs per SM, so that CTAs of consecutive launches
// share an SM (programmatic dependent launch, lpmx_const_stream.cu).  (A 2-warp shape, five per SM, lost at every size: r2y.)
// NREC: source records per launch (half a bank for small target sets: the CTAs hold their slots half as long when the
// pipeline of an evaluation drains, r2z).
template <int T, bool PF, bool SMALL = false, int NREC = cs::kBatch>
__global__ void __launch_bounds__(SMALL ? 160 : (PF ? cs_max_threads(T) : cs_max_threads(T) - 32), SMALL ? 3 : ((!PF && T <= 6) ? 2 : 1))
    pair_sum_const_kernel(const cs::CsArgs a) {
  if (PF && threadIdx.x >= blockDim.x - 32) {
    prefetch_bank(a.prefetch_stride, a.n_rec, a.acc);
    return;
  }
  CsDevice pf;
  pf.lanes = PF ? blockDim.x - 32 : blockDim.x;
  cs::body<T, NREC>(pf, a);
}

typedef void (*cs_kernel_t)(const cs::CsArgs);
template <bool PF>
cs_kernel_t cs_kernel_for(int T) {
#if LPMX_CS_BANK < 2
  switch (T) {
    case 3: return pair_sum_const_kernel<3, PF>;
    case 4: return pair_sum_const_kernel<4, PF>;
    case 5: return pair_sum_const_kernel<5, PF>;
    case 6: return pair_sum_const_kernel<6, PF>;
    case 7: return pair_sum_const_kernel<7, PF>;
    case 8: return pair_sum_const_kernel<8, PF>;
    default: return nullptr;
  }
#else
  return nullptr;
#endif
}

}  // namespace

#define LPMX_CS_CAT2(a, b) a##b
#define LPMX_CS_CAT(a, b) LPMX_CS_CAT2(a, b)

// cudaErrorInvalidValue when there is no kernel for T
// pdl: launch with programmatic stream serialization (may start before the preceding kernel of the stream has completed)
cudaError_t LPMX_CS_CAT(cs_bank_launch_, LPMX_CS_BANK)(int T, int grid, int threads, cudaStream_t stream, const cs::CsArgs& a, int pdl) {
  const bool pf = a.prefetch_stride > 0;  // then `threads` includes the prefetch warp
  cs_kernel_t kern = pf ? cs_kernel_for<true>(T) : cs_kernel_for<false>(T);
  if (pf && T == 6 && threads == 160) kern = a.n_rec == cs::kBatch / 2 ? pair_sum_const_kernel<6, true, true, cs::kBatch / 2> : pair_sum_const_kernel<6, true, true>;
  else if (a.n_rec != cs::kBatch) return cudaErrorInvalidValue;
  if (!kern || threads > cs_max_threads(T) - (pf ? 0 : 32)) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

// device-to-device refill of the bank from `records` (n_rec <= cs::kBatch records of cs::kRec doubles)
cudaError_t LPMX_CS_CAT(cs_bank_fill_, LPMX_CS_BANK)(const double* records, int n_rec, cudaStream_t stream) {
  return cudaMemcpyToSymbolAsync(c_src, records, sizeof(double) * n_rec * cs::kRec, 0, cudaMemcpyDeviceToDevice, stream);  // not the read-ahead record
}

}  // namespace lpmx
