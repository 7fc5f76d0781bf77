// Host model of pair_sum_const_kernel (lpm_b200/csrc/lpmx_const_stream_body.h -- the body the CUDA kernel runs): every CUDA
// thread is a loop iteration, the two constant banks plain arrays, the launch sequence of launch_const_stream (one launch per
// 1 280-record batch, alternating banks, zero-padded last batch, `first` on batch 0) restated around it.  Checks the kernel's
// indexing -- T targets per thread at a stride of blockDim, padded tail threads, the [3][n_tgt_pad] accumulator layout that
// the stage kernels read as "slot 0", accumulation across launches, self-pair exclusion by compact index, both target
// layouts -- against a direct double loop.
//   g++ -O2 -std=c++17 -ffp-contract=off -I lpm_b200/csrc tests/cpp/const_stream_model.cpp -o model && ./model
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "lpmx_const_stream_body.h"

using namespace lpmx::cs;

static std::vector<double> g_banks[2] = {std::vector<double>(kBankDoubles, 1e300), std::vector<double>(kBankDoubles, 1e300)};  // the read-ahead record stays poisoned
static int g_bank = 0;  // the bank the running launch's module sees

struct HostPlatform {
  int tid_, bid_, nt_;
  int tid() const { return tid_; }
  int bid() const { return bid_; }
  int n_threads() const { return nt_; }
  bool any_sync(bool p) const { return p; }  // per-thread here: the checked loop is a superset of the unchecked one
  double src(int i) const { return g_banks[g_bank][(size_t)i]; }
  double rcp_seed(double d) const { return (double)(float)(1.0 / d); }  // 24 good bits; the cubic step must do the rest
  void accumulate(double* p, double v, bool first) const { *p = first ? v : *p + v; }
  void launch_dependents() const {}
  void wait_prior() const {}
};

template <int T>
static void launch(const CsArgs& a, int grid, int threads) {
  for (int b = 0; b < grid; ++b)
    for (int t = 0; t < threads; ++t) {
      HostPlatform pf{t, b, threads};
      body<T>(pf, a);
    }
}

template <int T>
static int run_case(int n_tgt, int n_src, int threads, bool collocated, bool soa, unsigned seed, bool mapped = false) {
  std::mt19937_64 rng(seed);
  std::normal_distribution<double> nd;
  auto unit = [&](double* p) {
    double n2;
    do {
      p[0] = nd(rng), p[1] = nd(rng), p[2] = nd(rng);
      n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    } while (n2 < 1e-3);
    const double s = 1.0 / std::sqrt(n2);
    p[0] *= s, p[1] *= s, p[2] *= s;
  };
  // sources (compact leaf records) and targets; collocated: target i IS source i for i < n_src, the rest are other points
  std::vector<double> src(6 * (size_t)n_src), tx(3 * (size_t)n_tgt);
  for (int j = 0; j < n_src; ++j) {
    double y[3];
    unit(y);
    const double g = 0.1 * nd(rng);
    for (int k = 0; k < 3; ++k) src[6 * (size_t)j + k] = y[k], src[6 * (size_t)j + 3 + k] = g * y[k];
  }
  std::vector<int> self(n_tgt, -1);
  for (int i = 0; i < n_tgt; ++i) {
    double x[3];
    if (collocated && i < n_src) {
      for (int k = 0; k < 3; ++k) x[k] = src[6 * (size_t)i + k];
      self[i] = i;
    } else {
      unit(x);
    }
    for (int k = 0; k < 3; ++k) tx[soa ? (size_t)k * n_tgt + i : 3 * (size_t)i + k] = x[k];
  }
  // mapped: the launch's targets are an index list (a sharded solver's): target tg of the launch is element map[tg] of the
  // views; the accumulators stay in launch order
  std::vector<int> map(n_tgt);
  for (int i = 0; i < n_tgt; ++i) map[i] = i;
  if (mapped) std::shuffle(map.begin(), map.end(), rng);
  const double kappa = collocated ? 1.0 : 1.0 + 1e-4;
  const int tb = T * threads, grid = (n_tgt + tb - 1) / tb;
  const long n_tgt_pad = (long)grid * tb;
  std::vector<double> acc(3 * (size_t)n_tgt_pad, 1e300);  // poisoned: `first` must overwrite
  // launch_const_stream: batches of kBatch records, zero-padded, alternating banks
  const int n_batches = (n_src + kBatch - 1) / kBatch;
  CsArgs a{};
  a.tgt = tx.data();
  a.tgt_si = soa ? 1 : 3;
  a.tgt_sk = soa ? n_tgt : 1;
  a.tgt_map = mapped ? map.data() : nullptr;
  a.self_idx = collocated ? self.data() : nullptr;
  a.acc = acc.data();
  a.n_tgt_pad = n_tgt_pad;
  a.n_tgt = n_tgt;
  a.kappa = kappa;
  for (int b = 0; b < n_batches; ++b) {
    g_bank = b & 1;
    for (int j = 0; j < kBatch; ++j)
      for (int k = 0; k < kRec; ++k) {
        const long js = (long)b * kBatch + j;
        g_banks[g_bank][(size_t)kRec * j + k] = js < n_src ? src[6 * (size_t)js + k] : 0.0;
      }
    a.j0 = b * kBatch, a.n_rec = kBatch, a.first = b == 0;
    launch<T>(a, grid, threads);
  }
  // direct evaluation: d formed as the kernel forms it (for random points the closest pairs have d ~ 1/N^2, and ANY double
  // evaluation of 1 - x.y carries a relative error of 2^-53 / d there -- DESIGN.md section 7 -- so a long-double d would
  // measure that conditioning, not the kernel), the reciprocal exact, the terms added in source order
  double worst = 0, scale = 0;
  for (int li = 0; li < n_tgt; ++li) {
    const int i = map[li];
    double m[3] = {0, 0, 0};
    const double x[3] = {tx[soa ? i : 3 * (size_t)i], tx[soa ? (size_t)n_tgt + i : 3 * (size_t)i + 1],
                         tx[soa ? 2 * (size_t)n_tgt + i : 3 * (size_t)i + 2]};
    for (int j = 0; j < n_src; ++j) {
      if (j == self[i]) continue;
      const double* y = &src[6 * (size_t)j];
      const double d = std::fma(-x[0], y[0], std::fma(-x[1], y[1], std::fma(-x[2], y[2], kappa)));
      const double r = 1.0 / d;
      for (int k = 0; k < 3; ++k) m[k] = std::fma(r, y[3 + k], m[k]);
    }
    for (int k = 0; k < 3; ++k) {
      worst = std::fmax(worst, std::fabs(acc[(size_t)k * n_tgt_pad + li] - m[k]));
      scale = std::fmax(scale, std::fabs(m[k]));
    }
  }
  int bad = !(worst <= 1e-13 * scale);
  for (long i = n_tgt; i < n_tgt_pad; ++i)  // padded targets: finite, never poison
    for (int k = 0; k < 3; ++k) bad += !std::isfinite(acc[(size_t)k * n_tgt_pad + i]) || acc[(size_t)k * n_tgt_pad + i] == 1e300;
  std::printf("T=%d n_tgt=%d n_src=%d threads=%d colloc=%d soa=%d mapped=%d: rel err %.2e %s\n", T, n_tgt, n_src, threads,
              (int)collocated, (int)soa, (int)mapped, worst / scale, bad ? "FAILED" : "ok");
  return bad;
}

int main() {
  int bad = 0;
  bad += run_case<6>(1000, 3000, 64, false, true, 1);    // 3 batches, last one padded; tail threads
  bad += run_case<5>(3000, 2800, 96, true, true, 2);     // collocated: self pairs in batches 0..2
  bad += run_case<7>(777, 1280, 32, true, false, 3);     // exactly one batch, AoS targets
  bad += run_case<6>(2000, 5200, 128, true, false, 4);   // 5 batches: both banks reused
  bad += run_case<4>(129, 1400, 32, false, false, 5);
  bad += run_case<8>(300, 1281, 32, true, true, 6);      // second batch holds one record
  bad += run_case<6>(1500, 2700, 64, true, true, 7, true);   // an index list of targets, collocated
  bad += run_case<5>(900, 1300, 32, false, false, 8, true);
  return bad ? 1 : 0;
}
