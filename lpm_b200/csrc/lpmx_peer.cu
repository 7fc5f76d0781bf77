// lpmx_peer.cu -- the per-stage exchange of packed source records as ONE kernel over NVLink peer memory
// ("peer exchange").  Opt-in: lpmx_comm_enable_peer_exchange() or LPMX_PEER_EXCHANGE=1.
//
// The reference has no multi-device path at all; the default exchange here is a
// group of ncclBroadcast calls (lpmx_core.cu: comm_allgatherv).  At cubed-7 on eight GPUs the pair sums are
// down to 1.8 ms per evaluation and the NCCL group launch is what is left of the step, so this file replaces
// it, for buffers that were allocated through slab_alloc(), by peer_push_kernel:
//
//   phase 0  every rank tells every peer "my copy of the buffer may be overwritten" (all earlier readers of it
//            are ordered before this kernel on the stream)                      ready[rank] := epoch   on peer
//   phase 1  each CTA waits for ready[p] and stores this rank's segment into peer p's mapping of the slab
//            (16-byte stores over NVLink; peers are visited in a per-CTA rotated order)
//   phase 2  the last CTA to finish (system-scope fence + atomic ticket) tells every peer "my segment has
//            landed"  done[rank] := epoch on peer,  and waits until every peer said the same.
//
// The ready handshake makes the kernel as safe as the NCCL rendezvous for any call sequence (the same buffer
// may be exchanged twice in a row); the ping-pong of the steppers is not relied upon.  Every spin has a
// deadline (LPMX_PEER_TIMEOUT_S, default 600 s: it has to outlast any legitimate skew between ranks, e.g. one rank writing
// output -- r2m: a 10 s deadline fired because rank 0 alone ran a host-side check between two exchanges): on expiry the kernel records the peer in a host-mapped error
// word and returns, and the next lpmx_sync() / exchange reports LPMX_ERR_COMM -- it never hangs the GPU.
//
// Slabs are mapped with CUDA IPC (one process per GPU).  A mapping is validated by reading a magic word the
// owner wrote, and the outcome is all-gathered so that every rank takes the same decision (peer path, or the
// NCCL path when any mapping failed).
#include <dlfcn.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>

#include "lpmx_internal.h"
#include "lpmx_peer_protocol.h"

namespace lpmx {

namespace {

using namespace lpmx::peer;
static_assert(kMaxRanks == kMaxPeers, "lpmx_internal.h and lpmx_peer_protocol.h disagree on the rank limit");
constexpr size_t kFlagWords = (size_t)kFlagSlots * kFlagStride;
constexpr size_t kMinIpcBytes = (size_t)2 << 20;     // allocations below this may share a block with others
constexpr unsigned long long kMagicBase = 0x6c706d785f706565ull;  // "lpmx_pee"

struct Blob {  // what the ranks all-gather when a buffer is registered
  cudaIpcMemHandle_t handle;
  unsigned long long bytes;
  int device;
  int pid;
};

// the protocol's platform on the GPU (lpmx_peer_protocol.h)
struct DevicePlatform {
  int* sh;  // two shared words of the CTA
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int bid() const { return blockIdx.x; }
  __device__ __forceinline__ int n_threads() const { return blockDim.x; }
  __device__ __forceinline__ int n_blocks() const { return gridDim.x; }
  __device__ __forceinline__ unsigned long long now_ns() const {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
  }
  __device__ __forceinline__ void backoff() const { __nanosleep(200); }
  __device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) const {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) const {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  }
  __device__ __forceinline__ void fence_system() const { __threadfence_system(); }
  __device__ __forceinline__ void sync_threads() const { __syncthreads(); }
  __device__ __forceinline__ unsigned long long atomic_add(unsigned long long* p, unsigned long long v) const { return atomicAdd(p, v); }
  __device__ __forceinline__ void report(int* host_err, int v) const { *(volatile int*)host_err = v; }
  __device__ __forceinline__ int& s_ok() const { return sh[0]; }
  __device__ __forceinline__ int& s_last() const { return sh[1]; }
};

template <int VEC>
__global__ void __launch_bounds__(256) peer_push_kernel(const PushArgs a) {
  __shared__ int sh[2];
  DevicePlatform pf{sh};
  push_body<VEC>(pf, a);
}

struct ByeArgs {
  unsigned long long* flags_peer[kMaxRanks];
};
__global__ void peer_bye_kernel(int rank, int world, unsigned long long timeout_ns, unsigned long long* flags_local, ByeArgs b) {
  __shared__ int sh[2];
  DevicePlatform pf{sh};
  bye_body(pf, rank, world, timeout_ns, flags_local, b.flags_peer);
}

typedef int (*nccl_allgather_t)(const void*, void*, size_t, int, void*, cudaStream_t);

// all-gather `bytes` bytes per rank between host buffers over the handle's NCCL communicator (setup path only)
int allgather_host(lpmx_handle_t h, const void* in, void* out, size_t bytes) {
  static nccl_allgather_t ag = nullptr;
  if (!ag) ag = (nccl_allgather_t)dlsym(h->nccl_lib, "ncclAllGather");
  if (!ag) return set_error(h, LPMX_ERR_COMM, "ncclAllGather not found");
  void *s = nullptr, *r = nullptr;
  LPMX_TRY(dev_buffer(h, "peer_ag_send", bytes, &s));
  LPMX_TRY(dev_buffer(h, "peer_ag_recv", bytes * h->world, &r));
  LPMX_CUDA(h, cudaMemcpyAsync(s, in, bytes, cudaMemcpyHostToDevice, h->stream));
  const int kNcclInt8 = 0;
  if (ag(s, r, bytes, kNcclInt8, h->nccl_comm, h->stream) != 0) return set_error(h, LPMX_ERR_COMM, "ncclAllGather failed");
  LPMX_CUDA(h, cudaMemcpyAsync(out, r, bytes * h->world, cudaMemcpyDeviceToHost, h->stream));
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

// Map `local` (the base of a cudaMalloc allocation of `bytes` bytes; the 8 bytes at `magic_off` may be scribbled
// on and are zero afterwards) into every peer and vice versa.  On return *ok says whether EVERY rank validated
// EVERY mapping; peer[] holds the mappings (closed again when !*ok).  Collective.
int map_everywhere(lpmx_handle_t h, void* local, size_t bytes, size_t magic_off, void** peer, bool* ok) {
  const int world = h->world, rank = h->rank;
  *ok = false;
  for (int q = 0; q < kMaxPeers; ++q) peer[q] = nullptr;
  const unsigned long long magic = kMagicBase ^ (unsigned long long)(rank + 1);
  LPMX_CUDA(h, cudaMemcpyAsync((char*)local + magic_off, &magic, sizeof(magic), cudaMemcpyHostToDevice, h->stream));
  Blob mine;
  memset(&mine, 0, sizeof(mine));
  int good = 1;
  if (cudaIpcGetMemHandle(&mine.handle, local) != cudaSuccess) {
    cudaGetLastError();
    good = 0;
  }
  mine.bytes = bytes;
  mine.device = h->device;
  mine.pid = (int)getpid();
  std::vector<Blob> all(world);
  LPMX_TRY(allgather_host(h, &mine, all.data(), sizeof(Blob)));  // also orders the magic write before any peer read
  for (int q = 0; q < world && good; ++q) {
    if (q == rank) continue;
    if (all[q].bytes != bytes) {
      good = 0;
      break;
    }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      good = 0;
      break;
    }
    peer[q] = p;
    unsigned long long seen = 0;
    if (cudaMemcpy(&seen, (const char*)p + magic_off, sizeof(seen), cudaMemcpyDeviceToHost) != cudaSuccess) {
      cudaGetLastError();
      good = 0;
      break;
    }
    if (seen != (kMagicBase ^ (unsigned long long)(q + 1))) good = 0;
  }
  std::vector<int> verdict(world);
  LPMX_TRY(allgather_host(h, &good, verdict.data(), sizeof(int)));  // doubles as the barrier before the word is zeroed
  bool all_good = true;
  for (int q = 0; q < world; ++q) all_good = all_good && verdict[q] != 0;
  LPMX_CUDA(h, cudaMemsetAsync((char*)local + magic_off, 0, sizeof(magic), h->stream));
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (!all_good) {
    for (int q = 0; q < world; ++q)
      if (peer[q]) {
        cudaIpcCloseMemHandle(peer[q]);
        peer[q] = nullptr;
      }
    cudaGetLastError();
  }
  *ok = all_good;
  return LPMX_OK;
}

const PeerRegion* find_region(const PeerState* ps, const void* p) {
  for (const PeerRegion& r : ps->regions) {
    const char* b = (const char*)r.local;
    if ((const char*)p >= b && (const char*)p < b + r.bytes) return &r;
  }
  return nullptr;
}

}  // namespace

int peer_check_error(lpmx_handle_t h) {
  PeerState* ps = h->peer;
  if (!ps || !ps->host_err) return LPMX_OK;
  if (*ps->host_err != 0) {
    ps->dead_peer = *ps->host_err;
    *ps->host_err = 0;
  }
  if (ps->dead_peer == 0) return LPMX_OK;
  // fatal for this handle's exchanges: the ranks no longer agree on what has been delivered
  return set_error(h, LPMX_ERR_COMM, "peer exchange timed out waiting for rank %d (at or before exchange %llu); destroy the handle",
                   ps->dead_peer - 1, ps->epoch);
}

int slab_alloc(lpmx_handle_t h, void** out, size_t bytes) {
  PeerState* ps = h->peer;
  if (!ps || !ps->enabled || h->world == 1) {
    LPMX_CUDA(h, cudaMalloc(out, bytes));
    return LPMX_OK;
  }
  // own mapping granule, so the IPC handle names this allocation and nothing else
  const size_t cap = bytes < kMinIpcBytes ? kMinIpcBytes : bytes;
  LPMX_CUDA(h, cudaMalloc(out, cap));
  PeerRegion r;
  r.local = *out;
  r.bytes = cap;
  bool ok = false;
  LPMX_TRY(map_everywhere(h, r.local, cap, 0, r.peer, &ok));
  if (ok) ps->regions.push_back(r);  // otherwise the buffer is exchanged over NCCL like any other
  return LPMX_OK;
}

void slab_free(lpmx_handle_t h, void* p) {
  if (!p) return;
  PeerState* ps = h->peer;
  if (ps) {
    for (size_t i = 0; i < ps->regions.size(); ++i) {
      if (ps->regions[i].local != p) continue;
      // Peers may still hold a mapping of this slab, and freeing exported memory before they close it is
      // undefined: close ours of theirs, keep the memory until lpmx_destroy (after the teardown barrier).
      for (int q = 0; q < kMaxPeers; ++q)
        if (ps->regions[i].peer[q]) cudaIpcCloseMemHandle(ps->regions[i].peer[q]);
      cudaGetLastError();
      ps->regions.erase(ps->regions.begin() + i);
      ps->graveyard.push_back(p);
      return;
    }
  }
  cudaFree(p);
}

bool peer_can_exchange(lpmx_handle_t h, const double* base) {
  PeerState* ps = h->peer;
  return ps && ps->enabled && h->world > 1 && find_region(ps, base) != nullptr;
}

int peer_allgatherv(lpmx_handle_t h, double* base, const long* offsets, cudaStream_t stream) {
  if (!stream) stream = h->stream;
  PeerState* ps = h->peer;
  const PeerRegion* r = find_region(ps, base);
  if (!r) return set_error(h, LPMX_ERR_STATE, "peer exchange on an unregistered buffer");
  LPMX_TRY(peer_check_error(h));
  PushArgs a;
  memset(&a, 0, sizeof(a));
  a.rank = h->rank;
  a.world = h->world;
  a.epoch = ++ps->epoch;
  a.timeout_ns = ps->timeout_ns;
  a.src = base + offsets[h->rank];
  a.n = offsets[h->rank + 1] - offsets[h->rank];
  if (a.n < 0) return set_error(h, LPMX_ERR_INVALID, "decreasing offsets");
  const size_t rel = (size_t)((const char*)a.src - (const char*)r->local);
  if (rel + (size_t)a.n * sizeof(double) > r->bytes) return set_error(h, LPMX_ERR_INVALID, "segment leaves the registered slab");
  for (int q = 0; q < h->world; ++q) {
    a.dst[q] = q == h->rank ? nullptr : reinterpret_cast<double*>((char*)r->peer[q] + rel);
    a.flags_peer[q] = q == h->rank ? nullptr : ps->flags_peer[q];
  }
  a.flags_local = ps->flags_local;
  a.host_err = ps->host_err_dev;
  const bool vec2 = (rel % 16 == 0) && (a.n % 2 == 0);
  const size_t out_bytes = (size_t)a.n * sizeof(double) * (size_t)(h->world - 1);
  int grid = (int)(out_bytes / 32768);
  grid = grid < 1 ? 1 : grid > 64 ? 64 : grid;
  if (vec2)
    peer_push_kernel<2><<<grid, 256, 0, stream>>>(a);
  else
    peer_push_kernel<1><<<grid, 256, 0, stream>>>(a);
  ++h->launches;
  LPMX_CUDA(h, cudaGetLastError());
  return LPMX_OK;
}

void peer_teardown(lpmx_handle_t h) {
  PeerState* ps = h->peer;
  if (!ps) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (PeerRegion& r : ps->regions) {
    for (int q = 0; q < kMaxPeers; ++q)
      if (r.peer[q]) cudaIpcCloseMemHandle(r.peer[q]);
    ps->graveyard.push_back(r.local);
  }
  ps->regions.clear();
  if (ps->flags_local) {
    // nobody frees exported memory before every rank has closed its mappings (or 2 s have passed)
    ByeArgs b;
    memset(&b, 0, sizeof(b));
    for (int q = 0; q < h->world; ++q) b.flags_peer[q] = q == h->rank ? nullptr : ps->flags_peer[q];
    peer_bye_kernel<<<1, 32, 0, h->stream>>>(h->rank, h->world, 2000000000ull, ps->flags_local, b);
    cudaStreamSynchronize(h->stream);
    for (int q = 0; q < h->world; ++q)
      if (q != h->rank && ps->flags_peer[q]) cudaIpcCloseMemHandle(ps->flags_peer[q]);
  }
  for (void* p : ps->graveyard) cudaFree(p);
  if (ps->flags_local) cudaFree(ps->flags_local);
  if (ps->host_err) cudaFreeHost(ps->host_err);
  cudaGetLastError();
  delete ps;
  h->peer = nullptr;
}

int peer_enable(lpmx_handle_t h, int enable) {
  if (!enable) {
    if (h->peer) h->peer->enabled = false;  // mappings stay; exchanges go back to NCCL
    return LPMX_OK;
  }
  if (h->world == 1) return LPMX_OK;
  if (!h->nccl_comm) return set_error(h, LPMX_ERR_STATE, "lpmx_comm_enable_peer_exchange before lpmx_comm_init");
  if (h->world > kMaxPeers) return set_error(h, LPMX_ERR_UNSUPPORTED, "peer exchange supports at most %d ranks", kMaxPeers);
  if (h->peer) {
    h->peer->enabled = h->peer->flags_local != nullptr;
    return h->peer->enabled ? LPMX_OK : set_error(h, LPMX_ERR_UNSUPPORTED, "peer memory is not available between these GPUs");
  }
  LPMX_CUDA(h, cudaSetDevice(h->device));
  PeerState* ps = new (std::nothrow) PeerState;
  if (!ps) return LPMX_ERR_NOMEM;
  h->peer = ps;
  const char* t = getenv("LPMX_PEER_TIMEOUT_S");
  const double ts = t ? atof(t) : 30.0;
  ps->timeout_ns = (unsigned long long)((ts > 0 ? ts : 600.0) * 1e9);
  LPMX_CUDA(h, cudaHostAlloc((void**)&ps->host_err, sizeof(int), cudaHostAllocMapped));
  *ps->host_err = 0;
  LPMX_CUDA(h, cudaHostGetDevicePointer((void**)&ps->host_err_dev, ps->host_err, 0));
  void* f = nullptr;
  LPMX_CUDA(h, cudaMalloc(&f, kMinIpcBytes));
  LPMX_CUDA(h, cudaMemsetAsync(f, 0, kMinIpcBytes, h->stream));
  static_assert(kFlagWords * sizeof(unsigned long long) <= kMinIpcBytes, "flag block too small");
  // the magic word lives in its own slot so that validating the mapping never touches a live flag
  void* peer[kMaxPeers];
  bool ok = false;
  LPMX_TRY(map_everywhere(h, f, kMinIpcBytes, (size_t)kMagic * kFlagStride * sizeof(unsigned long long), peer, &ok));
  if (!ok) {
    ps->graveyard.push_back(f);  // freed at teardown
    ps->enabled = false;
    return set_error(h, LPMX_ERR_UNSUPPORTED, "peer memory is not available between these GPUs (CUDA IPC mapping failed)");
  }
  for (int q = 0; q < kMaxPeers; ++q) ps->flags_peer[q] = (unsigned long long*)peer[q];
  ps->flags_local = (unsigned long long*)f;
  ps->enabled = true;
  return LPMX_OK;
}

}  // namespace lpmx

extern "C" int lpmx_comm_enable_peer_exchange(lpmx_handle_t h, int enable) {
  if (!h) return LPMX_ERR_INVALID;
  return lpmx::peer_enable(h, enable);
}

extern "C" int lpmx_comm_peer_exchange_enabled(lpmx_handle_t h, int* enabled, int* n_regions) {
  if (!h) return LPMX_ERR_INVALID;
  if (enabled) *enabled = (h->peer && h->peer->enabled) ? 1 : 0;
  if (n_regions) *n_regions = h->peer ? (int)h->peer->regions.size() : 0;
  return LPMX_OK;
}
