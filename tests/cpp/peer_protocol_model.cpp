// Host model of the peer exchange (lpm_b200/csrc/lpmx_peer_protocol.h): the protocol code the CUDA kernel runs, with
// every CUDA thread a std::thread, every "GPU" a set of heap arrays, peer mappings plain pointers.  It checks what can
// be checked without a GPU -- flag indexing, epochs, rotated peer order, the last-CTA ticket, the ready handshake when
// the same buffer is exchanged again and again while a slow rank is still reading it, ragged / empty / odd segments,
// and that a missing rank produces a timeout report instead of a hang or a silent "done".  It says nothing about the
// GPU memory model (the real kernel's fences are in lpmx_peer.cu).
//
//   g++ -O1 -std=c++20 -pthread -I lpm_b200/csrc tests/cpp/peer_protocol_model.cpp -o model && ./model
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "lpmx_peer_protocol.h"

using namespace lpmx::peer;

struct HostPlatform {
  int tid_, bid_, nt_, nb_;
  std::barrier<>* cta;
  int* sh;
  int tid() const { return tid_; }
  int bid() const { return bid_; }
  int n_threads() const { return nt_; }
  int n_blocks() const { return nb_; }
  unsigned long long now_ns() const {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  void backoff() const { std::this_thread::yield(); }
  unsigned long long ld_acquire_sys(const unsigned long long* p) const {
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
  }
  void st_release_sys(unsigned long long* p, unsigned long long v) const { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
  void fence_system() const { std::atomic_thread_fence(std::memory_order_seq_cst); }
  void sync_threads() const { cta->arrive_and_wait(); }
  unsigned long long atomic_add(unsigned long long* p, unsigned long long v) const {
    return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL);
  }
  void report(int* host_err, int v) const { __atomic_store_n(host_err, v, __ATOMIC_RELAXED); }
  int& s_ok() const { return sh[0]; }
  int& s_last() const { return sh[1]; }
};

struct Gpu {
  std::vector<unsigned long long> flags = std::vector<unsigned long long>((size_t)kFlagSlots * kFlagStride, 0ull);
  std::vector<double> buf[2];
  int host_err = 0;
};

// one "kernel launch" of rank r: n_blocks CTAs of n_threads threads, joined before returning (stream order)
template <int VEC>
static void launch(const PushArgs& a, int n_blocks, int n_threads) {
  std::vector<std::thread> th;
  std::vector<std::unique_ptr<std::barrier<>>> bars;
  std::vector<std::unique_ptr<int[]>> shared;
  for (int b = 0; b < n_blocks; ++b) {
    bars.emplace_back(new std::barrier<>(n_threads));
    shared.emplace_back(new int[2]{0, 0});
  }
  for (int b = 0; b < n_blocks; ++b)
    for (int t = 0; t < n_threads; ++t)
      th.emplace_back([&, b, t] {
        HostPlatform pf{t, b, n_threads, n_blocks, bars[b].get(), shared[b].get()};
        push_body<VEC>(pf, a);
      });
  for (auto& x : th) x.join();
}

static double value_of(int rank, int exchange, long i) { return rank * 1.0e6 + exchange * 1.0e3 + (double)(i % 997) + 0.25; }

// world ranks exchange `n_exchanges` times; `pingpong` alternates two buffers like the steppers, otherwise the same
// buffer is reused every time (what lpmx_bve_solver_init_velocity followed by stream_fn does)
static int run_case(int world, const std::vector<long>& seg, int n_exchanges, bool pingpong, int n_blocks, int n_threads,
                    unsigned seed) {
  std::vector<long> off(world + 1, 0);
  for (int r = 0; r < world; ++r) off[r + 1] = off[r] + seg[r];
  std::vector<Gpu> gpu(world);
  for (auto& g : gpu)
    for (auto& b : g.buf) b.assign((size_t)off[world] + 2, -1.0);
  std::atomic<int> bad{0};
  std::vector<std::thread> ranks;
  for (int r = 0; r < world; ++r)
    ranks.emplace_back([&, r] {
      std::mt19937 rng(seed * 131 + r);
      for (int e = 1; e <= n_exchanges; ++e) {
        const int which = pingpong ? (e & 1) : 0;
        // the stage kernel: this rank's own records
        for (long i = off[r]; i < off[r + 1]; ++i) gpu[r].buf[which][i] = value_of(r, e, i);
        PushArgs a{};
        a.rank = r, a.world = world, a.epoch = (unsigned long long)e, a.timeout_ns = 20000000000ull;
        a.src = gpu[r].buf[which].data() + off[r];
        a.n = seg[r];
        for (int q = 0; q < world; ++q) {
          a.dst[q] = q == r ? nullptr : gpu[q].buf[which].data() + off[r];
          a.flags_peer[q] = q == r ? nullptr : gpu[q].flags.data();
        }
        a.flags_local = gpu[r].flags.data();
        a.host_err = &gpu[r].host_err;
        const bool vec2 = (off[r] % 2 == 0) && (seg[r] % 2 == 0);
        if (vec2)
          launch<2>(a, n_blocks, n_threads);
        else
          launch<1>(a, n_blocks, n_threads);
        // the pair sum: reads every record, slowly on some ranks -- nobody may overwrite them meanwhile
        if (rng() % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 300));
        for (int q = 0; q < world; ++q)
          for (long i = off[q]; i < off[q + 1]; ++i)
            if (gpu[r].buf[which][i] != value_of(q, e, i)) bad.fetch_add(1);
        if (gpu[r].buf[which][off[world]] != -1.0) bad.fetch_add(1);  // nothing written past the end
      }
      if (gpu[r].host_err != 0 || gpu[r].flags[kFail * kFlagStride] != 0) bad.fetch_add(1);
      if (gpu[r].flags[kTicket * kFlagStride] != 0) bad.fetch_add(1);
    });
  for (auto& t : ranks) t.join();
  return bad.load();
}

// rank `missing` never shows up: everyone else must report a timeout naming a rank, within the deadline, and must not
// have told anybody "done"
static int run_missing(int world, int missing) {
  std::vector<Gpu> gpu(world);
  const long n = 64;
  for (auto& g : gpu) g.buf[0].assign((size_t)n * world, 0.0);
  std::atomic<int> bad{0};
  std::vector<std::thread> ranks;
  for (int r = 0; r < world; ++r) {
    if (r == missing) continue;
    ranks.emplace_back([&, r] {
      PushArgs a{};
      a.rank = r, a.world = world, a.epoch = 1, a.timeout_ns = 200000000ull;  // 0.2 s
      a.src = gpu[r].buf[0].data() + n * r, a.n = n;
      for (int q = 0; q < world; ++q) {
        a.dst[q] = q == r ? nullptr : gpu[q].buf[0].data() + n * r;
        a.flags_peer[q] = q == r ? nullptr : gpu[q].flags.data();
      }
      a.flags_local = gpu[r].flags.data();
      a.host_err = &gpu[r].host_err;
      const auto t0 = std::chrono::steady_clock::now();
      launch<2>(a, 2, 4);
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (s > 5.0) bad.fetch_add(1);
      if (gpu[r].host_err != 1 + missing) bad.fetch_add(1);
      if (gpu[r].flags[kFail * kFlagStride] != 1) bad.fetch_add(1);
    });
  }
  for (auto& t : ranks) t.join();
  for (int r = 0; r < world; ++r)
    for (int q = 0; q < world; ++q)
      if (gpu[r].flags[(kDone + q) * kFlagStride] != 0) bad.fetch_add(1);  // no rank claimed delivery
  return bad.load();
}

int main() {
  int bad = 0;
  // peer_of visits every other rank exactly once
  for (int world = 2; world <= kMaxRanks; ++world)
    for (int rank = 0; rank < world; ++rank)
      for (int bid = 0; bid < 70; ++bid) {
        unsigned seen = 0;
        for (int k = 0; k < world - 1; ++k) {
          const int p = peer_of(rank, world, bid, k);
          if (p < 0 || p >= world || p == rank || (seen >> p & 1u)) ++bad;
          seen |= 1u << p;
        }
      }
  std::printf("peer_of %d\n", bad);
  int c = 0;
  c += run_case(2, {96, 160}, 40, true, 1, 4, 1);
  c += run_case(2, {96, 160}, 40, false, 3, 4, 2);          // same buffer every time
  c += run_case(3, {0, 30, 50}, 30, false, 2, 3, 3);        // a rank that owns no leaf, even offsets
  c += run_case(3, {7, 0, 33}, 30, true, 2, 3, 4);          // odd offsets -> scalar copies
  c += run_case(4, {64, 64, 64, 64}, 25, false, 3, 4, 5);
  c += run_case(8, {0, 0, 0, 40, 48, 48, 48, 48}, 12, true, 2, 8, 6);   // vertices first: early ranks own no sources
  c += run_case(8, {16, 16, 16, 16, 16, 16, 16, 16}, 12, false, 1, 8, 7);
  std::printf("exchange %d\n", c);
  int m = run_missing(3, 1) + run_missing(2, 0);
  std::printf("missing %d\n", m);
  return (bad || c || m) ? 1 : 0;
}
