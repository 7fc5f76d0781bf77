// lpmx_core.cu -- handle, error reporting, scratch/staging buffers, NCCL bootstrap.
#include <dlfcn.h>

#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "lpmx_internal.h"

namespace lpmx {

int set_error(lpmx_handle_t h, int code, const char* fmt, ...) {
  if (h) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    h->err = std::string(lpmx_error_name(code)) + ": " + buf;
  }
  return code;
}

int check_cuda(lpmx_handle_t h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return LPMX_OK;
  const int code = (e == cudaErrorMemoryAllocation) ? LPMX_ERR_NOMEM : LPMX_ERR_CUDA;
  return set_error(h, code, "%s -> %s", what, cudaGetErrorString(e));
}

int dev_buffer(lpmx_handle_t h, const char* name, size_t bytes, void** out) {
  DevBuf& b = h->bufs[name];
  if (b.cap < bytes) {
    if (b.p) {
      LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
      LPMX_CUDA(h, cudaFree(b.p));
      b.p = nullptr;
      b.cap = 0;
    }
    const size_t cap = bytes + bytes / 8 + 256;
    LPMX_CUDA(h, cudaMalloc(&b.p, cap));
    b.cap = cap;
  }
  *out = b.p;
  return LPMX_OK;
}

int pinned_buffer(lpmx_handle_t h, const char* name, size_t bytes, void** out) {
  DevBuf& b = h->pinned[name];
  if (b.cap < bytes) {
    if (b.p) {
      LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
      LPMX_CUDA(h, cudaFreeHost(b.p));
      b.p = nullptr;
      b.cap = 0;
    }
    const size_t cap = bytes + bytes / 8 + 256;
    LPMX_CUDA(h, cudaMallocHost(&b.p, cap));
    b.cap = cap;
  }
  *out = b.p;
  return LPMX_OK;
}

bool is_device_pointer(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int stage_in(lpmx_handle_t h, const char* name, const void* user, size_t bytes, const void** dev) {
  if (!user || is_device_pointer(user)) {
    *dev = user;
    return LPMX_OK;
  }
  void* d = nullptr;
  LPMX_TRY(dev_buffer(h, name, bytes, &d));
  LPMX_CUDA(h, cudaMemcpyAsync(d, user, bytes, cudaMemcpyHostToDevice, h->stream));
  *dev = d;
  return LPMX_OK;
}

int stage_out_begin(lpmx_handle_t h, const char* name, void* user, size_t bytes, void** dev) {
  if (!user || is_device_pointer(user)) {
    *dev = user;
    return LPMX_OK;
  }
  return dev_buffer(h, name, bytes, dev);
}

int stage_out_end(lpmx_handle_t h, void* user, const void* dev, size_t bytes) {
  if (!user || user == dev) return LPMX_OK;
  LPMX_CUDA(h, cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
  return LPMX_OK;
}

int comm_allgatherv(lpmx_handle_t h, double* base, const long* offsets, cudaStream_t stream) {
  if (h->world == 1) return LPMX_OK;
  if (!stream) stream = h->stream;
  if (!h->nccl_comm || !h->nccl_lib) return set_error(h, LPMX_ERR_COMM, "world > 1 but lpmx_comm_init was not called");
  if (peer_can_exchange(h, base)) return peer_allgatherv(h, base, offsets, stream);  // one kernel over NVLink peer memory
  typedef int (*group_t)(void);
  typedef int (*bcast_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static group_t gstart = nullptr, gend = nullptr;
  static bcast_t bcast = nullptr;
  if (!bcast) {
    gstart = (group_t)dlsym(h->nccl_lib, "ncclGroupStart");
    gend = (group_t)dlsym(h->nccl_lib, "ncclGroupEnd");
    bcast = (bcast_t)dlsym(h->nccl_lib, "ncclBroadcast");
    if (!gstart || !gend || !bcast) return set_error(h, LPMX_ERR_COMM, "NCCL symbols not found");
  }
  const int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64
  int rc = gstart();
  for (int r = 0; r < h->world && rc == 0; ++r) {
    const long n = offsets[r + 1] - offsets[r];
    if (n <= 0) continue;
    double* seg = base + offsets[r];
    rc = bcast(seg, seg, (size_t)n, kNcclFloat64, r, h->nccl_comm, stream);
  }
  const int rc2 = gend();
  if (rc != 0 || rc2 != 0) return set_error(h, LPMX_ERR_COMM, "ncclBroadcast group failed (%d/%d)", rc, rc2);
  return LPMX_OK;
}

}  // namespace lpmx

using namespace lpmx;

extern "C" {

const char* lpmx_version_string(void) { return "lpmx 0.1 (sm_100a)"; }

const char* lpmx_error_name(int code) {
  switch (code) {
    case LPMX_OK: return "LPMX_OK";
    case LPMX_ERR_INVALID: return "LPMX_ERR_INVALID";
    case LPMX_ERR_CUDA: return "LPMX_ERR_CUDA";
    case LPMX_ERR_NOMEM: return "LPMX_ERR_NOMEM";
    case LPMX_ERR_NO_DEVICE: return "LPMX_ERR_NO_DEVICE";
    case LPMX_ERR_COMM: return "LPMX_ERR_COMM";
    case LPMX_ERR_UNSUPPORTED: return "LPMX_ERR_UNSUPPORTED";
    case LPMX_ERR_STATE: return "LPMX_ERR_STATE";
    default: return "LPMX_ERR_UNKNOWN";
  }
}

int lpmx_create(lpmx_handle_t* out, int device_id) {
  if (!out) return LPMX_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return LPMX_ERR_NO_DEVICE;  // no CPU fallback, by design
  }
  if (device_id < 0 || device_id >= n) return LPMX_ERR_INVALID;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return LPMX_ERR_CUDA;
  if (prop.major < 10) return LPMX_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  if (cudaSetDevice(device_id) != cudaSuccess) return LPMX_ERR_CUDA;
  lpmx_handle_s* h = new (std::nothrow) lpmx_handle_s;
  if (!h) return LPMX_ERR_NOMEM;
  h->device = device_id;
  h->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return LPMX_ERR_CUDA;
  }
  *out = h;
  return LPMX_OK;
}

int lpmx_bve_solver_destroy(lpmx_bve_solver_t s);
int lpmx_ic2d_solver_destroy(lpmx_ic2d_solver_t s);
int lpmx_swe_solver_destroy(lpmx_swe_solver_t s);

int lpmx_destroy(lpmx_handle_t h) {
  if (!h) return LPMX_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->cached_bve) lpmx_bve_solver_destroy(h->cached_bve);
  if (h->cached_ic2d) lpmx_ic2d_solver_destroy(h->cached_ic2d);
  if (h->cached_swe) lpmx_swe_solver_destroy(h->cached_swe);
  if (h->cached_plane) lpmx_plane_swe_solver_destroy(h->cached_plane);
  const_stream_teardown(h);
  peer_teardown(h);  // before the NCCL communicator goes: closes the IPC mappings, then frees what was exported
  for (auto& kv : h->bufs)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : h->pinned)
    if (kv.second.p) cudaFreeHost(kv.second.p);
  for (int i = 0; i < 2; ++i)
    if (h->xchg_ev[i]) cudaEventDestroy(h->xchg_ev[i]);
  for (auto& ev : h->prof_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  if (h->nccl_comm && h->nccl_lib) {
    typedef int (*destroy_t)(void*);
    destroy_t f = (destroy_t)dlsym(h->nccl_lib, "ncclCommDestroy");
    if (f) f(h->nccl_comm);
  }
  cudaStreamDestroy(h->stream);
  cudaStreamDestroy(h->copy_stream);
  delete h;
  return LPMX_OK;
}

int lpmx_sync(lpmx_handle_t h) {
  if (!h) return LPMX_ERR_INVALID;
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return peer_check_error(h);
}

const char* lpmx_last_error_string(lpmx_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int lpmx_stream(lpmx_handle_t h, void** s) {
  if (!h || !s) return LPMX_ERR_INVALID;
  *s = (void*)h->stream;
  return LPMX_OK;
}

int lpmx_launch_count(lpmx_handle_t h, long* n) {
  if (!h || !n) return LPMX_ERR_INVALID;
  *n = h->launches;
  return LPMX_OK;
}

int lpmx_copy(lpmx_handle_t h, void* dst, const void* src, long bytes) {
  if (!h || bytes < 0 || (bytes > 0 && (!dst || !src))) return h ? set_error(h, LPMX_ERR_INVALID, "bad copy arguments") : LPMX_ERR_INVALID;
  if (bytes == 0) return LPMX_OK;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  LPMX_CUDA(h, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, h->stream));
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return LPMX_OK;
}

int lpmx_profile_enable(lpmx_handle_t h, int enable) {
  if (!h) return LPMX_ERR_INVALID;
  h->profile = enable != 0;
  return LPMX_OK;
}

int lpmx_profile_read(lpmx_handle_t h, long* n_launches, double* total_ms, double* pair_visits) {
  if (!h) return LPMX_ERR_INVALID;
  LPMX_CUDA(h, cudaStreamSynchronize(h->stream));
  double ms = 0;
  const bool dump = getenv("LPMX_PROFILE_DUMP") != nullptr;  // per-launch durations and the gaps between them, to stderr
  for (size_t i = 0; i < h->prof_used; ++i) {
    float t = 0;
    LPMX_CUDA(h, cudaEventElapsedTime(&t, h->prof_events[i].first, h->prof_events[i].second));
    ms += t;
    if (dump) {
      float gap = 0;
      if (i > 0) cudaEventElapsedTime(&gap, h->prof_events[i - 1].second, h->prof_events[i].first);
      fprintf(stderr, "[lpmx profile] rank %d launch %zu: %.3f ms, gap before %.3f ms\n", h->rank, i, t, gap);
    }
  }
  if (n_launches) *n_launches = (long)h->prof_used;
  if (total_ms) *total_ms = ms;
  if (pair_visits) *pair_visits = h->prof_pairs;
  h->prof_used = 0;
  h->prof_pairs = 0;
  return LPMX_OK;
}

int lpmx_set_partition(lpmx_handle_t h, int rank, int world) {
  if (!h || world < 1 || rank < 0 || rank >= world) return LPMX_ERR_INVALID;
  h->rank = rank;
  h->world = world;
  return LPMX_OK;
}

int lpmx_set_io_sharded(lpmx_handle_t h, int on) {
  if (!h) return LPMX_ERR_INVALID;
  h->io_sharded = on != 0;
  return LPMX_OK;
}

int lpmx_fp64_peak_tflops(lpmx_handle_t h, double* tflops, double* ms) {
  if (!h) return LPMX_ERR_INVALID;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  return fp64_probe(h, tflops, ms);
}

// ---- NCCL bootstrap.  The library binds NCCL at run time (dlopen) so that a process which has
// already loaded an NCCL (torch's bundled one) shares it, and single-GPU users need none. ----
static void* open_nccl() {
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  return lib;
}

int lpmx_comm_unique_id(void* id128) {
  if (!id128) return LPMX_ERR_INVALID;
  void* lib = open_nccl();
  if (!lib) return LPMX_ERR_COMM;
  typedef int (*getid_t)(void*);
  getid_t f = (getid_t)dlsym(lib, "ncclGetUniqueId");
  if (!f) return LPMX_ERR_COMM;
  return f(id128) == 0 ? LPMX_OK : LPMX_ERR_COMM;
}

int lpmx_comm_init(lpmx_handle_t h, const void* id128, int rank, int world) {
  if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return LPMX_ERR_INVALID;
  LPMX_CUDA(h, cudaSetDevice(h->device));
  h->nccl_lib = open_nccl();
  if (!h->nccl_lib) return set_error(h, LPMX_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
  struct Id {
    char b[128];
  } id;
  memcpy(id.b, id128, 128);
  typedef int (*init_t)(void**, int, Id, int);
  init_t f = (init_t)dlsym(h->nccl_lib, "ncclCommInitRank");
  if (!f) return set_error(h, LPMX_ERR_COMM, "ncclCommInitRank not found");
  void* comm = nullptr;
  const int rc = f(&comm, world, id, rank);
  if (rc != 0) return set_error(h, LPMX_ERR_COMM, "ncclCommInitRank failed (%d)", rc);
  h->nccl_comm = comm;
  h->rank = rank;
  h->world = world;
  // The peer-memory exchange is on by default on one box of <= 8 GPUs (measured bit-identical to the NCCL exchange on 2 and 8
  // GPUs, r2h / r2m; it is what lets the sharded steppers overlap the exchange with the pair sum); LPMX_PEER_EXCHANGE=0 keeps the
  // NCCL exchange.  When the GPUs cannot map each other's memory the NCCL exchange stays and the call still succeeds.
  const char* e = getenv("LPMX_PEER_EXCHANGE");
  if (!(e && atoi(e) == 0) && world > 1 && world <= kMaxPeers) {
    if (peer_enable(h, 1) != LPMX_OK) fprintf(stderr, "lpmx: %s; keeping the NCCL exchange\n", h->err.c_str());
  }
  return LPMX_OK;
}

}  // extern "C"
