// tune_pair_sum.cu -- times alternative shapes of the pair-sum kernel on the GPU box.
// Build (here):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//                     -Ilpm_b200/csrc -o tools/tune_pair_sum tools/tune_pair_sum.cu
// Run (gpurun):  ./tools/tune_pair_sum [n_tgt n_src]
// Development tool; not part of the product library.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lpmx_pair_kernel.cuh"

using namespace lpmx;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

static int g_sms = 148;
static double* d_tgt;   // SoA [3][n_tgt]
static double* d_src;   // packed
static double* d_part;
static size_t part_cap;
static int n_tgt, n_src;
static double ref_sum = 0;

static void fib(int n, std::vector<double>& x, double rot) {
  x.resize(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    double z = 1.0 - (2.0 * i + 1.0) / n;
    double r = sqrt(std::max(0.0, 1.0 - z * z));
    double phi = i * (M_PI * (3.0 - sqrt(5.0))) + rot;
    x[3 * (size_t)i] = r * cos(phi);
    x[3 * (size_t)i + 1] = r * sin(phi);
    x[3 * (size_t)i + 2] = z;
  }
}

template <class C>
static void run(const char* label) {
  const int tb = C::TB;
  const int n_tb = (n_tgt + tb - 1) / tb;
  const int n_sc = n_src / kChunk;
  const long n_items = (long)n_tb * n_sc;
  int grid = C::MINB * g_sms;
  if (grid > n_items) grid = (int)n_items;
  int max_slots = 1;
  for (int t = 0; t < n_tb; ++t) {
    int c0 = cta_of_item((long)t * n_sc, grid, n_items), c1 = cta_of_item((long)(t + 1) * n_sc - 1, grid, n_items);
    max_slots = std::max(max_slots, c1 - c0 + 1);
  }
  const long n_tgt_pad = (long)n_tb * tb;
  const size_t need = (size_t)max_slots * kind_nacc(C::KIND) * n_tgt_pad * sizeof(double);
  if (need > part_cap) {
    printf("%-28s skipped (partials %zu MB)\n", label, need >> 20);
    return;
  }
  SumArgs a;
  a.tgt.p = d_tgt;
  a.tgt.si = 1;
  a.tgt.sk = n_tgt;
  a.self_idx = nullptr;
  a.tgt_map = nullptr;
  a.packed = d_src;
  a.part = d_part;
  a.n_tgt = n_tgt;
  a.n_tb = n_tb;
  a.n_sc = n_sc;
  a.n_tgt_pad = n_tgt_pad;
  a.kappa = 1.0;
  a.aux = 0.0;
  const size_t smem = pair_smem_bytes(C::KIND, C::T, C::LANES);
  auto kern = pair_sum_kernel<C>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    kern<<<grid, C::THREADS, smem>>>(a);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  // checksum of slot-0 x-moment over the first 1000 targets (sanity: all shapes agree to rounding)
  std::vector<double> hp(1000);
  CK(cudaMemcpy(hp.data(), d_part, 1000 * sizeof(double), cudaMemcpyDeviceToHost));
  double cs = 0;
  for (double v : hp) cs += v;
  (void)cs;
  const double pairs = (double)n_tgt * n_src;
  printf("%-28s regs %3d occ %d/%d grid %4d  %8.3f ms  %7.4f T-pairs/s  alg %6.2f TF/s\n", label, fa.numRegs, occ,
         C::MINB, grid, best, pairs / best * 1e-9, pairs * 24 / best * 1e-9);
  fflush(stdout);
}


// ---- FP64 pipe probes: DFMA only, DMMA (m8n8k4) only, and both interleaved -------------------
template <int MODE>
__global__ void __launch_bounds__(256) fp64_mix_kernel(double* out, int iters) {
  double v[8], c0[4], c1[4];
  for (int k = 0; k < 8; ++k) v[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  for (int k = 0; k < 4; ++k) c0[k] = c1[k] = 0.0;
  const double a = 1.0 + 1e-7 * threadIdx.x, b = 0.999999;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fma(v[k], b, 1e-7);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[k]), "+d"(c1[k])
                     : "d"(a), "d"(b));
    }
  }
  double s = 0;
  for (int k = 0; k < 8; ++k) s += v[k];
  for (int k = 0; k < 4; ++k) s += c0[k] + c1[k];
  if (s == 123.456) out[0] = s;
}

template <int MODE>
static void mix(const char* label) {
  const int iters = 1 << 14, blocks = g_sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    fp64_mix_kernel<MODE><<<blocks, threads>>>(d_part, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  const double n = (double)iters * blocks * threads;
  const double dfma_flops = (MODE != 1) ? n * 8 * 2 : 0;
  const double dmma_flops = (MODE != 0) ? n / 32 * 4 * 512 : 0;
  printf("%-22s %8.3f ms   DFMA %6.2f TF/s   DMMA %6.2f TF/s\n", label, best, dfma_flops / best * 1e-9,
         dmma_flops / best * 1e-9);
}

// ---- register-file operand bandwidth probe: DFMA with 1, 2 or 3 distinct register operand pairs ----
template <int MODE>
__global__ void __launch_bounds__(256) rf_probe_kernel(double* out, int iters, double seed) {
  double v[8], x[8], y[8];
  for (int k = 0; k < 8; ++k) {
    v[k] = seed + 1e-9 * (threadIdx.x + k);
    x[k] = 0.999 + 1e-6 * (threadIdx.x + 3 * k) * seed;
    y[k] = 1e-7 * (threadIdx.x + 5 * k) * seed;
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) v[k] = fma(v[k], 0.999999, 1e-7);   // 1 register operand pair
      if (MODE == 1) v[k] = fma(v[k], x[0], 1e-7);       // 2, one shared (reuse cache)
      if (MODE == 2) v[k] = fma(v[k], x[k], 1e-7);       // 2 distinct
      if (MODE == 3) v[k] = fma(x[k], y[0], v[k]);       // 3, one shared
      if (MODE == 4) v[k] = fma(x[k], y[k], v[k]);       // 3 distinct
      if (MODE == 5) v[k] = fma(x[k], y[(k + i) & 7], v[k]);
    }
  }
  double s = 0;
  for (int k = 0; k < 8; ++k) s += v[k] + x[k] + y[k];
  if (s == 123.456) out[0] = s;
}
template <int MODE>
static void rfp(const char* label) {
  const int iters = 1 << 14, blocks = g_sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    rf_probe_kernel<MODE><<<blocks, threads>>>(d_part, iters, 1.0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  const double n = (double)iters * blocks * threads;
  printf("%-34s %8.3f ms   DFMA %6.2f TF/s\n", label, best, n * 16 / best * 1e-9);
}

// ---- constant-bank source streaming probe: sources read through LDCU into uniform registers ----
constexpr int kConstSources = 672;  // 672 * 48 B = 32256 B
__constant__ double c_src[kConstSources * 6];
template <int T, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) const_pair_kernel(const double* tgt, double* part, int n_tgt, int reps) {
  const long base = (long)blockIdx.x * (T * NW * 32) + threadIdx.x;
  double x[T][3], acc[T][3];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base + (long)t * NW * 32;
    const bool v = tg < n_tgt;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      x[t][k] = v ? tgt[(long)k * n_tgt + tg] : 0.0;
      acc[t][k] = 0.0;
    }
  }
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
    for (int j = 0; j < kConstSources; ++j) {
      const double y0 = c_src[6 * j], y1 = c_src[6 * j + 1], y2 = c_src[6 * j + 2];
      const double g0 = c_src[6 * j + 3], g1 = c_src[6 * j + 4], g2 = c_src[6 * j + 5];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const double d = fma(-x[t][0], y0, fma(-x[t][1], y1, fma(-x[t][2], y2, 1.0)));
        const double r0 = rcp_seed(d);
        const double e = fma(-d, r0, 1.0);
        const double p = fma(e, e, e);
        const double r = fma(r0, p, r0);
        acc[t][0] = fma(r, g0, acc[t][0]);
        acc[t][1] = fma(r, g1, acc[t][1]);
        acc[t][2] = fma(r, g2, acc[t][2]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long tg = base + (long)t * NW * 32;
    if (tg < n_tgt)
      for (int k = 0; k < 3; ++k) part[(long)k * n_tgt + tg] = acc[t][k];
  }
}
template <int T, int NW, int MINB>
static void cprobe(const char* label) {
  const int tb = T * NW * 32;
  const int grid = (n_tgt + tb - 1) / tb;
  const int reps = 64;
  auto kern = const_pair_kernel<T, NW, MINB>;
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    kern<<<grid, NW * 32>>>(d_tgt, d_part, n_tgt, reps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  const double pairs = (double)n_tgt * kConstSources * reps;
  printf("%-28s regs %3d grid %5d  %8.3f ms  %7.4f T-pairs/s  alg %6.2f TF/s\n", label, fa.numRegs, grid, best,
         pairs / best * 1e-9, pairs * 24 / best * 1e-9);
  fflush(stdout);
}

int main(int argc, char** argv) {
  n_tgt = argc > 1 ? atoi(argv[1]) : 229376;
  n_src = argc > 2 ? atoi(argv[2]) : 98304;
  n_src = (n_src / kChunk) * kChunk;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("device %s, %d SMs; n_tgt %d n_src %d\n", prop.name, g_sms, n_tgt, n_src);
  std::vector<double> ht, hs, soa(3 * (size_t)n_tgt), pk(8 * (size_t)n_src);
  fib(n_tgt, ht, 0.3);
  fib(n_src, hs, 0.0);
  for (int i = 0; i < n_tgt; ++i)
    for (int k = 0; k < 3; ++k) soa[(size_t)k * n_tgt + i] = ht[3 * (size_t)i + k];
  for (int j = 0; j < n_src; ++j) {
    const double gam = -(hs[3 * (size_t)j + 2] * 4 * M_PI / n_src) / (4 * M_PI);
    for (int k = 0; k < 3; ++k) {
      pk[8 * (size_t)j + k] = hs[3 * (size_t)j + k];
      pk[8 * (size_t)j + 3 + k] = gam * hs[3 * (size_t)j + k];
    }
    pk[8 * (size_t)j + 6] = gam;
    pk[8 * (size_t)j + 7] = 0;
  }
  CK(cudaMalloc(&d_tgt, soa.size() * sizeof(double)));
  CK(cudaMalloc(&d_src, pk.size() * sizeof(double)));
  part_cap = (size_t)1 << 30;
  CK(cudaMalloc(&d_part, part_cap));
  CK(cudaMemcpy(d_tgt, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_src, pk.data(), pk.size() * sizeof(double), cudaMemcpyHostToDevice));

#define RUNK(K, T, NW, MINB, UNR) run<PairCfg<K, T, NW, MINB, UNR>>(#K " T" #T " NW" #NW " B" #MINB " U" #UNR)
#define RUN(T, NW, MINB, UNR) run<PairCfg<kVel, T, NW, MINB, UNR>>("vel T" #T " NW" #NW " B" #MINB " U" #UNR)
  if (argc > 3 && !strcmp(argv[3], "r2")) {  // round 2: the product shapes and their neighbours 

    RUN(6, 8, 1, 2);
    RUN(6, 8, 1, 4);
    RUN(7, 8, 1, 2);
    RUN(8, 8, 1, 2);
    RUN(8, 8, 1, 4);
    RUN(4, 8, 2, 2);
    RUNK(kVelPsi, 4, 8, 1, 2);
    RUNK(kVelPsi, 4, 8, 1, 4);
    RUNK(kVelPsi, 6, 8, 1, 2);
    RUNK(kPsi, 8, 8, 1, 2);
    return 0;
  }
  if (argc > 5) {  // stream-function kinds only: more warps per SM against the two divergent table lookups per pair
    RUNK(kVelPsi, 4, 8, 1, 2);
    RUNK(kVelPsi, 4, 12, 1, 2);
    RUNK(kVelPsi, 4, 16, 1, 2);
    RUNK(kVelPsi, 3, 12, 1, 2);
    RUNK(kVelPsi, 3, 16, 1, 2);
    RUNK(kVelPsi, 2, 16, 1, 2);
    RUNK(kVelPsi, 2, 24, 1, 2);
    RUNK(kVelPsi, 2, 8, 2, 2);
    RUNK(kVelPsi, 2, 12, 2, 2);
    RUNK(kVelPsi, 2, 8, 3, 2);
    RUNK(kVelPsi, 4, 8, 1, 4);
    RUNK(kPsi, 8, 8, 1, 2);
    RUNK(kPsi, 6, 12, 1, 2);
    RUNK(kPsi, 4, 16, 1, 2);
    RUNK(kPsi, 4, 24, 1, 2);
    RUNK(kPsi, 4, 12, 2, 2);
    RUNK(kPsi, 2, 16, 2, 2);
    return 0;
  }
  if (argc > 3) goto probes;
  RUN(4, 8, 2, 2);
  RUN(4, 8, 2, 1);
  RUN(4, 8, 2, 4);
  RUN(4, 16, 1, 2);
  RUN(4, 4, 4, 2);
  RUN(4, 12, 1, 2);
  RUN(3, 8, 2, 2);
  RUN(3, 8, 3, 2);
  RUN(3, 8, 3, 1);
  RUN(3, 12, 2, 2);
  RUN(3, 6, 4, 2);
  RUN(2, 8, 2, 2);
  RUN(2, 8, 3, 2);
  RUN(2, 8, 4, 2);
  RUN(2, 8, 4, 4);
  RUN(2, 16, 2, 2);
  RUN(2, 12, 2, 2);
  RUN(2, 4, 8, 2);
  RUN(1, 8, 4, 4);
  RUN(1, 16, 2, 4);
  RUN(6, 8, 1, 2);
  RUN(6, 8, 1, 1);
  RUN(6, 16, 1, 1);
  RUN(8, 8, 1, 1);
  RUN(8, 4, 2, 1);
  if (argc <= 3) {
    RUNK(kVelPsi, 2, 8, 1, 2);
    RUNK(kVelPsi, 2, 8, 2, 2);
    RUNK(kVelPsi, 2, 16, 1, 2);
    RUNK(kVelPsi, 4, 8, 1, 1);
    RUNK(kVelPsi, 4, 8, 2, 1);
    RUNK(kPsi, 2, 8, 2, 2);
    RUNK(kPsi, 4, 8, 2, 2);
    RUNK(kPsi, 4, 16, 1, 2);
  }
probes:

  RUN(6, 8, 1, 2);
  RUN(6, 8, 1, 4);
  RUN(6, 12, 1, 2);
  RUN(6, 12, 1, 4);
  RUN(7, 8, 1, 2);
  RUN(8, 8, 1, 2);
  RUN(8, 8, 1, 4);
  RUN(8, 4, 1, 2);
  RUN(8, 4, 2, 2);
  RUN(5, 8, 1, 2);
  RUN(5, 12, 1, 2);
  RUN(4, 12, 1, 2);
  RUN(4, 16, 1, 2);
  RUN(4, 8, 2, 2);
  RUN(3, 8, 3, 2);
  RUNK(kVelPsi, 2, 16, 1, 2);
  RUNK(kVelPsi, 2, 8, 2, 2);
  RUNK(kVelPsi, 3, 8, 1, 2);
  RUNK(kVelPsi, 3, 12, 1, 2);
  RUNK(kVelPsi, 4, 8, 1, 2);
  RUNK(kVelPsi, 4, 8, 1, 1);
  RUNK(kVelPsi, 4, 12, 1, 2);
  RUNK(kVelPsi, 6, 8, 1, 1);
  RUNK(kVelPsi, 6, 8, 1, 2);
  RUNK(kPsi, 4, 16, 1, 2);
  RUNK(kPsi, 4, 8, 1, 2);
  RUNK(kPsi, 4, 8, 2, 2);
  RUNK(kPsi, 6, 8, 1, 2);
  RUNK(kPsi, 8, 8, 1, 2);
  RUNK(kPsi, 8, 8, 1, 1);
  RUNK(kPsi, 2, 8, 2, 2);
  RUNK(kSwe, 2, 8, 1, 2);
  RUNK(kSwe, 2, 12, 1, 2);
  RUNK(kSwe, 1, 16, 1, 2);
  RUNK(kSwe, 3, 8, 1, 1);
  if (argc > 4) return 0;
  {
    std::vector<double> cs(kConstSources * 6);
    for (int j = 0; j < kConstSources; ++j) {
      const double gam = -(hs[3 * (size_t)j + 2] * 4 * M_PI / n_src) / (4 * M_PI);
      for (int k = 0; k < 3; ++k) {
        cs[6 * j + k] = hs[3 * (size_t)j + k];
        cs[6 * j + 3 + k] = gam * hs[3 * (size_t)j + k];
      }
    }
    CK(cudaMemcpyToSymbol(c_src, cs.data(), cs.size() * sizeof(double)));
    cprobe<4, 8, 2>("const T4 NW8 B2");
    cprobe<4, 16, 1>("const T4 NW16 B1");
    cprobe<2, 8, 4>("const T2 NW8 B4");
    cprobe<2, 16, 2>("const T2 NW16 B2");
    cprobe<1, 16, 2>("const T1 NW16 B2");
    cprobe<6, 8, 1>("const T6 NW8 B1");
    cprobe<8, 8, 1>("const T8 NW8 B1");
    cprobe<3, 8, 3>("const T3 NW8 B3");
  }
  rfp<0>("rf: 1 reg operand");
  rfp<1>("rf: 2 reg operands, 1 shared");
  rfp<2>("rf: 2 distinct");
  rfp<3>("rf: 3 reg operands, 1 shared");
  rfp<4>("rf: 3 distinct");
  mix<0>("DFMA only");
  mix<1>("DMMA only");
  mix<2>("DFMA + DMMA");
  return 0;
}
