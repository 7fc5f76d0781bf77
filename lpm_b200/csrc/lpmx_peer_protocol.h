// lpmx_peer_protocol.h -- the exchange protocol of lpmx_peer.cu, written once over a small platform interface so
// that the very same code runs as the CUDA kernel (DevicePlatform, lpmx_peer.cu) and as a host model in which every
// CUDA thread is a std::thread (tests/cpp/peer_protocol_model.cpp, `-m "not gpu"`).  The model cannot say anything
// about the GPU memory model; it pins the flag indexing, the epochs, the rotated peer order, the last-CTA ticket, the
// ready handshake (a buffer exchanged twice in a row) and the deadline handling.
//
// Platform P provides:
//   int tid(), bid(), n_threads(), n_blocks();          position in the launch
//   unsigned long long now_ns();  void backoff();        deadline clock, polite spinning
//   unsigned long long ld_acquire_sys(const unsigned long long*);  void st_release_sys(unsigned long long*, v);
//   void fence_system();  void sync_threads();           __threadfence_system / __syncthreads
//   unsigned long long atomic_add(unsigned long long*, v);
//   int& s_ok(), int& s_last();                          two per-CTA shared words
//   void report(int* host_err, int v);                   error word (several threads may report at once)
#ifndef LPMX_PEER_PROTOCOL_H
#define LPMX_PEER_PROTOCOL_H

#if defined(__CUDACC__)
#define LPMX_PEER_HD __host__ __device__ __forceinline__
#else
#define LPMX_PEER_HD inline
#endif

namespace lpmx {
namespace peer {

constexpr int kMaxRanks = 8;
constexpr int kFlagStride = 16;            // 128 bytes between flags
constexpr int kReady = 0;                  // ready[q]: rank q's copy of the buffer may be overwritten      (epoch)
constexpr int kDone = kMaxRanks;           // done[q]:  rank q's segment has landed here                    (epoch)
constexpr int kBye = 2 * kMaxRanks;        // bye[q]:   rank q has closed its mappings                      (teardown)
constexpr int kTicket = 3 * kMaxRanks;     // CTA ticket counter of the running launch
constexpr int kMagic = 3 * kMaxRanks + 1;  // mapping validation word
constexpr int kFail = 3 * kMaxRanks + 2;   // set, never cleared, when a wait of this rank expired
constexpr int kFlagSlots = 3 * kMaxRanks + 3;

struct PushArgs {
  int rank, world;
  unsigned long long epoch;       // number of this exchange, 1-based, identical on every rank
  unsigned long long timeout_ns;  // deadline of every wait, from the start of the launch
  const double* src;              // this rank's segment in its own slab
  long n;                         // doubles in the segment
  double* dst[kMaxRanks];         // the same segment in rank q's slab (dst[rank] unused)
  unsigned long long* flags_local;
  unsigned long long* flags_peer[kMaxRanks];
  int* host_err;                  // 1 + rank a wait gave up on
};

// k-th peer visited by CTA `bid` of rank `rank`: every rank but `rank`, each exactly once for k = 0 .. world-2
LPMX_PEER_HD int peer_of(int rank, int world, int bid, int k) {
  const int np = world - 1;
  return (rank + 1 + (k + bid) % np) % world;
}

// spin until *p >= want or the deadline passes; false on timeout
template <class P>
LPMX_PEER_HD bool wait_flag(P& pf, const unsigned long long* p, unsigned long long want, unsigned long long deadline) {
  while (pf.ld_acquire_sys(p) < want) {
    if (pf.now_ns() > deadline) return false;
    pf.backoff();
  }
  return true;
}

template <int VEC, class P>
LPMX_PEER_HD void push_body(P& pf, const PushArgs& a) {
  const int tid = pf.tid(), bid = pf.bid();
  const unsigned long long deadline = pf.now_ns() + a.timeout_ns;
  // phase 0: every earlier reader of this rank's buffer is ordered before this launch
  if (bid == 0 && tid < a.world && tid != a.rank) pf.st_release_sys(a.flags_peer[tid] + (kReady + a.rank) * kFlagStride, a.epoch);
  // phase 1: this rank's segment into every peer that is ready for it
  for (int k = 0; k < a.world - 1; ++k) {
    const int p = peer_of(a.rank, a.world, bid, k);
    if (tid == 0) {
      pf.s_ok() = wait_flag(pf, a.flags_local + (kReady + p) * kFlagStride, a.epoch, deadline) ? 1 : 0;
      if (!pf.s_ok()) {
        pf.report(a.host_err, 1 + p);
        pf.st_release_sys(a.flags_local + kFail * kFlagStride, 1ull);
      }
    }
    pf.sync_threads();
    const bool ok = pf.s_ok() != 0;
    pf.sync_threads();
    if (!ok) continue;
    const long first = (long)bid * pf.n_threads() + tid, stride = (long)pf.n_blocks() * pf.n_threads();
    if (VEC == 2) {
      struct alignas(16) D2 {
        double x, y;
      };
      const D2* s2 = reinterpret_cast<const D2*>(a.src);
      D2* d2 = reinterpret_cast<D2*>(a.dst[p]);
      for (long i = first; i < a.n / 2; i += stride) d2[i] = s2[i];
    } else {
      for (long i = first; i < a.n; i += stride) a.dst[p][i] = a.src[i];
    }
  }
  // phase 2: every store of this CTA is ordered before its ticket, every ticket before the last CTA's flags
  pf.fence_system();
  pf.sync_threads();
  if (tid == 0) {
    const unsigned long long t = pf.atomic_add(a.flags_local + kTicket * kFlagStride, 1ull);
    pf.s_last() = (t == (unsigned long long)pf.n_blocks() - 1) ? 1 : 0;
  }
  pf.sync_threads();
  if (!pf.s_last()) return;
  if (tid == 0) a.flags_local[kTicket * kFlagStride] = 0;  // the next launch on this stream starts from zero
  pf.fence_system();
  // a rank that could not deliver says nothing, so that its peers run into their own deadline instead of computing
  // on records that never arrived (the fail word was written before its CTA's ticket)
  if (pf.ld_acquire_sys(a.flags_local + kFail * kFlagStride) != 0) return;
  if (tid < a.world && tid != a.rank) {
    pf.st_release_sys(a.flags_peer[tid] + (kDone + a.rank) * kFlagStride, a.epoch);
    if (!wait_flag(pf, a.flags_local + (kDone + tid) * kFlagStride, a.epoch, deadline)) {
      pf.report(a.host_err, 1 + tid);
      pf.st_release_sys(a.flags_local + kFail * kFlagStride, 1ull);
    }
  }
}

// teardown barrier (one CTA): bye[rank] := 1 on every peer, wait -- briefly -- for theirs
template <class P>
LPMX_PEER_HD void bye_body(P& pf, int rank, int world, unsigned long long timeout_ns, unsigned long long* flags_local,
                           unsigned long long* const* flags_peer) {
  const int tid = pf.tid();
  const unsigned long long deadline = pf.now_ns() + timeout_ns;
  if (tid < world && tid != rank) {
    pf.st_release_sys(flags_peer[tid] + (kBye + rank) * kFlagStride, 1ull);
    wait_flag(pf, flags_local + (kBye + tid) * kFlagStride, 1ull, deadline);
  }
}

}  // namespace peer
}  // namespace lpmx

#endif
