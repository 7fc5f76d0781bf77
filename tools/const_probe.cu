// const_probe.cu -- does streaming the (warp-uniform) source records through the constant bank, so that they reach the
// DFMAs as UNIFORM-register operands (LDCU -> UR), beat the shared-memory ring?  Development probe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/const_probe tools/const_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ double rcp_seed(double d) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
constexpr int NS = 1280;               // sources per batch: 1280 * 48 B = 61440 B of the 64 KB constant bank
__constant__ double c_src[NS * 6];

template <int T, int U, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) k(const double* __restrict__ tgt, double* __restrict__ out, int n_tgt, int n_src, double kappa) {
  const long base = (long)blockIdx.x * (T * NW * 32) + threadIdx.x;
  double x[T][3], acc[T][3];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base + (long)t * NW * 32;
#pragma unroll
    for (int k2 = 0; k2 < 3; ++k2) { x[t][k2] = tg < n_tgt ? tgt[(long)k2 * n_tgt + tg] : 0.0; acc[t][k2] = out[(long)k2 * n_tgt + (tg < n_tgt ? tg : 0)]; }
  }
#pragma unroll U
  for (int j = 0; j < n_src; ++j) {
    const double s0 = c_src[6 * j], s1 = c_src[6 * j + 1], s2 = c_src[6 * j + 2], s3 = c_src[6 * j + 3], s4 = c_src[6 * j + 4], s5 = c_src[6 * j + 5];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const double d = fma(-x[t][0], s0, fma(-x[t][1], s1, fma(-x[t][2], s2, kappa)));
      const double r0 = rcp_seed(d);
      const double e = fma(-d, r0, 1.0);
      const double p = fma(e, e, e);
      const double r = fma(r0, p, r0);
      acc[t][0] = fma(r, s3, acc[t][0]);
      acc[t][1] = fma(r, s4, acc[t][1]);
      acc[t][2] = fma(r, s5, acc[t][2]);
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base + (long)t * NW * 32;
    if (tg < n_tgt)
#pragma unroll
      for (int k2 = 0; k2 < 3; ++k2) out[(long)k2 * n_tgt + tg] = acc[t][k2];
  }
}

// duplicate-operand probes for the FP64 pipe
template <int MODE>
__global__ void __launch_bounds__(256) dup_probe(double* out, const double* in) {
  double v[16], w[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { v[k] = in[k + threadIdx.x % 3]; w[k] = in[32 + k + threadIdx.x % 5]; }
  for (int i = 0; i < 4096; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (MODE == 0) v[k] = fma(v[k], v[k], v[k]);        // e, e, e
        if (MODE == 1) v[k] = fma(v[k], w[k], v[k]);        // r0, p, r0
        if (MODE == 2) v[k] = fma(v[k], w[k], 1e-7);        // 2 distinct
      }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 123.456) out[0] = s;
}

template <class F> static float best_ms(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (r > 0 && ms < best) best = ms; }
  return best;
}

template <int T, int U, int NW, int MINB>
static void run(const double* d_tgt, double* d_out, const double* d_src, int n_tgt, int n_src_total, const char* label) {
  const int tb = T * NW * 32, grid = (n_tgt + tb - 1) / tb;
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k<T, U, NW, MINB>));
  CK(cudaMemset(d_out, 0, sizeof(double) * 3 * n_tgt));
  const int batches = n_src_total / NS;
  float ms = best_ms([&] {
    for (int b = 0; b < batches; ++b) {
      CK(cudaMemcpyToSymbolAsync(c_src, d_src + (size_t)b * NS * 6, sizeof(double) * NS * 6, 0, cudaMemcpyDeviceToDevice, 0));
      k<T, U, NW, MINB><<<grid, NW * 32>>>(d_tgt, d_out, n_tgt, NS, 1.0);
    }
  }, 3);
  CK(cudaGetLastError());
  const double pairs = (double)n_tgt * NS * batches;
  printf("%-22s regs %3d grid %4d batches %3d  %8.3f ms  %7.4f T-pairs/s  alg %6.2f TF/s\n", label, fa.numRegs, grid, batches, ms, pairs / ms * 1e-9, pairs * 24 / ms * 1e-9);
  fflush(stdout);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int sms = p.multiProcessorCount;
  double *d_in, *d_o; CK(cudaMalloc(&d_in, 4096)); CK(cudaMemset(d_in, 0, 4096)); CK(cudaMalloc(&d_o, 64));
  const char* names[3] = {"fma(v,v,v)", "fma(v,w,v)", "fma(v,w,imm)"};
  for (int m = 0; m < 3; ++m) {
    float ms = m == 0 ? best_ms([&] { dup_probe<0><<<sms * 4, 256>>>(d_o, d_in); }) : m == 1 ? best_ms([&] { dup_probe<1><<<sms * 4, 256>>>(d_o, d_in); }) : best_ms([&] { dup_probe<2><<<sms * 4, 256>>>(d_o, d_in); });
    const double lanes = (double)sms * 4 * 256 * 4096 * 64 / (ms * 1e-3 * clk_khz * 1e3) / sms;
    printf("dup probe %-14s %7.3f ms  %6.2f DFMA-lanes/clk/SM (%.1f%% of 64)\n", names[m], ms, lanes, lanes / 64 * 100);
  }
  const int n_tgt = 229376, n_src = 98304 / NS * NS;
  std::vector<double> ht(3 * (size_t)n_tgt), hs(6 * (size_t)n_src);
  for (int i = 0; i < n_tgt; ++i) { double z = 1 - (2.0 * i + 1) / n_tgt, r = sqrt(1 - z * z), ph = i * 2.399963 + 0.3; ht[i] = r * cos(ph); ht[n_tgt + i] = r * sin(ph); ht[2 * (size_t)n_tgt + i] = z; }
  for (int j = 0; j < n_src; ++j) { double z = 1 - (2.0 * j + 1) / n_src, r = sqrt(1 - z * z), ph = j * 2.399963; double g = -z / n_src; hs[6 * (size_t)j] = r * cos(ph); hs[6 * (size_t)j + 1] = r * sin(ph); hs[6 * (size_t)j + 2] = z; hs[6 * (size_t)j + 3] = g * r * cos(ph); hs[6 * (size_t)j + 4] = g * r * sin(ph); hs[6 * (size_t)j + 5] = g * z; }
  double *d_tgt, *d_out, *d_src;
  CK(cudaMalloc(&d_tgt, ht.size() * 8)); CK(cudaMalloc(&d_out, ht.size() * 8)); CK(cudaMalloc(&d_src, hs.size() * 8));
  CK(cudaMemcpy(d_tgt, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_src, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice));
  run<6, 1, 8, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T6 U1 NW8");
  run<6, 2, 8, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T6 U2 NW8");
  run<6, 4, 8, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T6 U4 NW8");
  run<8, 1, 8, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T8 U1 NW8");
  run<8, 2, 8, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T8 U2 NW8");
  run<4, 2, 8, 2>(d_tgt, d_out, d_src, n_tgt, n_src, "T4 U2 NW8 B2");
  run<4, 2, 16, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T4 U2 NW16");
  run<3, 2, 16, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T3 U2 NW16");
  run<6, 2, 4, 2>(d_tgt, d_out, d_src, n_tgt, n_src, "T6 U2 NW4 B2");
  run<12, 1, 4, 1>(d_tgt, d_out, d_src, n_tgt, n_src, "T12 U1 NW4");
  // checksum against a direct evaluation of target 0
  std::vector<double> ho(3 * (size_t)n_tgt); CK(cudaMemcpy(ho.data(), d_out, ho.size() * 8, cudaMemcpyDeviceToHost));
  double m[3] = {0, 0, 0};
  for (int j = 0; j < n_src; ++j) { double d = 1 - (ht[0] * hs[6 * (size_t)j] + ht[n_tgt] * hs[6 * (size_t)j + 1] + ht[2 * (size_t)n_tgt] * hs[6 * (size_t)j + 2]); for (int q = 0; q < 3; ++q) m[q] += hs[6 * (size_t)j + 3 + q] / d; }
  printf("check target 0: gpu (%.15e %.15e %.15e) host (%.15e %.15e %.15e)  [gpu accumulates over repeated timing launches: ratio %.6f]\n", ho[0], ho[n_tgt], ho[2 * (size_t)n_tgt], m[0], m[1], m[2], ho[0] / m[0]);
  return 0;
}
