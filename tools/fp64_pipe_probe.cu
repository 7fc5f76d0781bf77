// fp64_pipe_probe.cu -- microbenchmarks of the sm_100a FP64 issue path (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp64_pipe_probe tools/fp64_pipe_probe.cu
// Reports, per variant, DFMA-equivalents per clock per SM (64 = nominal) so that kernel design decisions
// (operand reuse, MUFU / LDS / integer co-issue, DMMA as a second pipe) rest on measurements.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

// V0: 16 chains, v = fma(v, a, b) with a,b kernel constants (1 register operand)
__global__ void __launch_bounds__(256) k_const(double* out, double a, double b) {
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = fma(v[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}
// V1: 3 distinct register operands, no reuse possible: v[k] = fma(p[k], q[k], v[k])
__global__ void __launch_bounds__(256) k_reg3(double* out, const double* in) {
  double v[16], p[16], q[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { v[k] = in[k]; p[k] = in[16 + k + threadIdx.x % 3]; q[k] = in[40 + k + threadIdx.x % 5]; }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = fma(p[k], q[(k + r) & 15], v[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}
// V2: 2 distinct + one shared operand that can sit in the reuse cache: v[k] = fma(p, q[k], v[k])
__global__ void __launch_bounds__(256) k_reg2(double* out, const double* in) {
  double v[16], q[16];
  double p = in[threadIdx.x % 7];
#pragma unroll
  for (int k = 0; k < 16; ++k) { v[k] = in[k]; q[k] = in[40 + k + threadIdx.x % 5]; }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = fma(p, q[(k + r) & 15], v[k]);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}
// V3: const-operand DFMA with one MUFU.RCP64H per 9 DFMA
__global__ void __launch_bounds__(256) k_mufu(double* out, double a, double b) {
  double v[18], m[2];
#pragma unroll
  for (int k = 0; k < 18; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  m[0] = 1.5; m[1] = 2.5;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < 18; ++k) v[k] = fma(v[k], a, b);
#pragma unroll
      for (int j = 0; j < 2; ++j) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(m[j]) : "d"(m[j]));
    }
  }
  double s = m[0] + m[1];
#pragma unroll
  for (int k = 0; k < 18; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}
// V4: const-operand DFMA with one integer op per 9 DFMA
__global__ void __launch_bounds__(256) k_int(double* out, double a, double b, int c) {
  double v[18];
  int m[2] = {(int)threadIdx.x, c};
#pragma unroll
  for (int k = 0; k < 18; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < 18; ++k) v[k] = fma(v[k], a, b);
#pragma unroll
      for (int j = 0; j < 2; ++j) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(m[j]) : "r"(c));
    }
  }
  double s = m[0] + m[1];
#pragma unroll
  for (int k = 0; k < 18; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}
// V5: const-operand DFMA with one LDS.128 per 18 DFMA
__global__ void __launch_bounds__(256) k_lds(double* out, double a, double b) {
  __shared__ double2 sm[512];
  sm[threadIdx.x] = make_double2(a, b);
  sm[threadIdx.x + 256] = make_double2(b, a);
  __syncthreads();
  double v[18];
  double2 acc = make_double2(0, 0);
#pragma unroll
  for (int k = 0; k < 18; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < 18; ++k) v[k] = fma(v[k], a, b);
      double2 t;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"((unsigned)__cvta_generic_to_shared(&sm[(i + r) & 511])));
      if (t.x == 77.0) acc.y = t.y;
    }
  }
  double s = acc.x + acc.y;
#pragma unroll
  for (int k = 0; k < 18; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// V6: DMMA m8n8k4 only, 8 independent accumulator pairs. 256 FMA per instruction per warp = 8 DFMA-warp-equivalents.
__global__ void __launch_bounds__(256) k_dmma(double* out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k) dmma884(c[k][0], c[k][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  if (s == 123.456) out[0] = s;
}
// V7: DMMA interleaved with DFMA: per 1 DMMA (8 DFMA-equivalents of math), NF DFMAs
template <int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, double a, double b) {
  double c[4][2], v[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; }
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dmma884(c[k][0], c[k][1], a, b);
#pragma unroll
        for (int f = 0; f < NF; ++f) v[(k * NF + f) & 15] = fma(v[(k * NF + f) & 15], a, b);
      }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}

// accuracy of the MUFU.RCP64H seed: max over d of |1 - d * rcp.approx.ftz.f64(d)|
__global__ void k_rcp_acc(double* out, double lo, double ratio, int n) {
  double worst = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = lo * pow(ratio, (double)i);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fabs(fma(-d, r, 1.0));
    worst = fmax(worst, e);
  }
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(worst));
}

template <class F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 1e30;
  for (int rep = 0; rep < 8; ++rep) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 2 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int sms = p.multiProcessorCount;
  const double clk = clk_khz * 1e3;
  printf("%s: %d SMs, %.0f MHz\n", p.name, sms, clk * 1e-6);
  double *out, *in; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&in, 4096)); CK(cudaMemset(in, 0, 4096));
  {
    CK(cudaMemset(out, 0, 8));
    const int n = 1 << 24;
    k_rcp_acc<<<sms * 4, 256>>>(out, 1e-12, pow(4e12, 1.0 / n), n);
    double worst = 0;
    CK(cudaMemcpy(&worst, out, 8, cudaMemcpyDeviceToHost));
    printf("MUFU.RCP64H seed: max |1 - d*r0| over %d log-spaced d in [1e-12, 4] = %.3e = 2^%.2f\n", n, worst, log2(worst));
  }
  for (int bps : {1, 2, 4}) {   // blocks of 256 threads per SM -> 2, 4, 8 warps per scheduler
    const int blocks = sms * bps; const double warps = (double)blocks * 8;
    auto rep = [&](const char* name, double ms, double dfma_per_thread_iter, double other) {
      const double dfma_warp = warps * ITERS * 4 * dfma_per_thread_iter;            // warp-level DFMA(-equivalent) instructions
      const double per_clk_sm = dfma_warp * 32 / (ms * 1e-3 * clk) / sms;            // lanes per clock per SM
      printf("  [%d warps/SMSP] %-34s %8.3f ms  %6.2f DFMA-lanes/clk/SM (%.1f%% of 64)%s\n", bps * 2, name, ms, per_clk_sm,
             per_clk_sm / 64 * 100, other > 0 ? "  (+other)" : "");
    };
    rep("DFMA const operands", time_ms([&] { k_const<<<blocks, 256>>>(out, 0.999999, 1e-7); }), 16, 0);
    rep("DFMA 3 distinct regs", time_ms([&] { k_reg3<<<blocks, 256>>>(out, in); }), 16, 0);
    rep("DFMA 2 regs + 1 reusable", time_ms([&] { k_reg2<<<blocks, 256>>>(out, in); }), 16, 0);
    rep("DFMA + MUFU.RCP64H (9:1)", time_ms([&] { k_mufu<<<blocks, 256>>>(out, 0.999999, 1e-7); }), 18, 1);
    rep("DFMA + IMAD (9:1)", time_ms([&] { k_int<<<blocks, 256>>>(out, 0.999999, 1e-7, 3); }), 18, 1);
    rep("DFMA + LDS.128 (18:1)", time_ms([&] { k_lds<<<blocks, 256>>>(out, 0.999999, 1e-7); }), 18, 1);
    rep("DMMA m8n8k4 only (8 eq each)", time_ms([&] { k_dmma<<<blocks, 256>>>(out, 0.5, 0.25); }), 8 * 8, 0);
    rep("DMMA + 2 DFMA (math 8+2)", time_ms([&] { k_mix<2><<<blocks, 256>>>(out, 0.5, 0.25); }), 4 * (8 + 2), 0);
    rep("DMMA + 4 DFMA (math 8+4)", time_ms([&] { k_mix<4><<<blocks, 256>>>(out, 0.5, 0.25); }), 4 * (8 + 4), 0);
    rep("DMMA + 8 DFMA (math 8+8)", time_ms([&] { k_mix<8><<<blocks, 256>>>(out, 0.5, 0.25); }), 4 * (8 + 8), 0);
    rep("DMMA + 16 DFMA (math 8+16)", time_ms([&] { k_mix<16><<<blocks, 256>>>(out, 0.5, 0.25); }), 4 * (8 + 16), 0);
  }
  return 0;
}
