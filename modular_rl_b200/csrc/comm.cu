// NCCL binding without a link-time dependency: libnccl.so.2 is resolved with dlopen at the first
// mrl_comm_* call (inside a torch process it is already mapped, so the very same library serves
// torch.distributed and this communicator).  Only sum all-reduces of small fp64 vectors are needed:
// the parameter-sized gradient / Fisher-vector product and a handful of loss scalars (DESIGN.md
// "Multi-GPU").  fp64 on the wire keeps every rank's replicated CG bit-identical and costs
// <= 356 KB per message for the largest network.
#include "comm.h"
#include "../../include/mrl_b200.h"
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
static NcclApi g_nccl;

static thread_local std::string g_comm_err;
extern "C" const char* mrl_last_error(void);
int mrl_set_error(const char* msg);   // api.cu

static int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return mrl_set_error("mrl_comm: cannot dlopen libnccl.so.2 (import torch first, or add nvidia/nccl/lib to LD_LIBRARY_PATH)");
  g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
    return mrl_set_error("mrl_comm: libnccl is missing a required symbol");
  g_nccl.handle = h;
  return 0;
}
static int nccl_fail(const char* what, int rc) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
  return mrl_set_error(buf);
}

// exported buffer of one rank: [flags 2 (parity) x 8, raised by the peers][this rank's vector, 2 (parity) x cap]
struct P2pState {
  bool on = false;
  long long cap = 0;                                  // doubles per slot
  unsigned char* local = nullptr;                     // this rank's buffer (cudaMalloc)
  unsigned char* peer[MRL_P2P_MAX_WORLD] = {nullptr}; // every rank's buffer as mapped here (peer[rank] == local)
  unsigned int* counter = nullptr;
  unsigned long long seq = 0;
  int* h_err = nullptr;                               // mapped pinned host word: set by a gather that timed out
  int* d_err = nullptr;                               // its device alias
  unsigned long long timeout_ns = 600ull * 1000000000ull;
};
#define P2P_HEADER 256
static inline unsigned long long* p2p_flags(unsigned char* base, int parity) {
  return reinterpret_cast<unsigned long long*>(base) + parity * MRL_P2P_MAX_WORLD;
}
static inline double* p2p_vector(unsigned char* base, long long cap, int parity) {
  return reinterpret_cast<double*>(base + P2P_HEADER) + (size_t)parity * cap;
}

struct mrl_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  P2pState p2p;
};
int mrl_comm_world(const mrl_comm* c) { return c ? c->world : 1; }
int mrl_comm_rank(const mrl_comm* c) { return c ? c->rank : 0; }

extern "C" int mrl_comm_unique_id(char id_out[128]) {
  if (load_nccl()) return 1;
  ncclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != NCCL_SUCCESS) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id_out, id.internal, 128);
  return 0;
}
extern "C" int mrl_comm_create(mrl_comm** out, const char id[128], int rank, int world, int device) {
  if (!out || !id || world < 1 || rank < 0 || rank >= world) return mrl_set_error("mrl_comm_create: bad arguments");
  if (load_nccl()) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return mrl_set_error("mrl_comm_create: cudaSetDevice failed");
  mrl_comm* c = new mrl_comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  ncclUniqueId uid;
  memcpy(uid.internal, id, 128);
  int rc = g_nccl.CommInitRank(&c->comm, world, uid, rank);
  if (rc != NCCL_SUCCESS) {
    delete c;
    return nccl_fail("ncclCommInitRank", rc);
  }
  *out = c;
  return 0;
}
extern "C" int mrl_comm_destroy(mrl_comm* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->p2p.local) {
    cudaDeviceSynchronize();
    for (int q = 0; q < c->world; ++q)
      if (q != c->rank && c->p2p.peer[q]) cudaIpcCloseMemHandle(c->p2p.peer[q]);
    cudaFree(c->p2p.local);
    cudaFree(c->p2p.counter);
    if (c->p2p.h_err) cudaFreeHost(c->p2p.h_err);
  }
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

// ---------------------------------------------------------------------------------- peer-memory transport
// Step 1 (every rank): allocate the exported buffer, return its CUDA IPC handle (64 bytes).
extern "C" int mrl_comm_p2p_export(mrl_comm* c, long long max_doubles, char handle_out[64]) {
  if (!c || !handle_out || max_doubles <= 0) return mrl_set_error("mrl_comm_p2p_export: bad arguments");
  if (c->world > MRL_P2P_MAX_WORLD) return mrl_set_error("mrl_comm_p2p_export: world > 8");
  if (cudaSetDevice(c->device) != cudaSuccess) return mrl_set_error("mrl_comm_p2p_export: cudaSetDevice failed");
  P2pState& p = c->p2p;
  if (p.local) return mrl_set_error("mrl_comm_p2p_export: already exported");
  p.cap = (max_doubles + 31) / 32 * 32;
  const size_t bytes = P2P_HEADER + (size_t)2 * p.cap * 8;
  if (cudaMalloc(&p.local, bytes) != cudaSuccess || cudaMalloc(&p.counter, 4) != cudaSuccess)
    return mrl_set_error("mrl_comm_p2p_export: out of device memory");
  cudaMemset(p.local, 0, bytes);
  cudaMemset(p.counter, 0, 4);
  if (cudaHostAlloc((void**)&p.h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&p.d_err, p.h_err, 0) != cudaSuccess)
    return mrl_set_error("mrl_comm_p2p_export: cannot allocate the mapped error word");
  *p.h_err = 0;
  if (const char* t = getenv("MRL_P2P_TIMEOUT_S")) {
    const double sec = atof(t);
    if (sec > 0) p.timeout_ns = (unsigned long long)(sec * 1e9);
  }
  cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p.local);
  if (e != cudaSuccess) return mrl_set_error(cudaGetErrorString(e));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  memcpy(handle_out, &h, 64);
  return 0;
}
// Step 2 (every rank, after an all-gather of the handles by the caller): map the peers' buffers.
extern "C" int mrl_comm_p2p_connect(mrl_comm* c, const char* handles /* [world][64] */) {
  if (!c || !handles || !c->p2p.local) return mrl_set_error("mrl_comm_p2p_connect: export first");
  if (cudaSetDevice(c->device) != cudaSuccess) return mrl_set_error("mrl_comm_p2p_connect: cudaSetDevice failed");
  P2pState& p = c->p2p;
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) { p.peer[q] = p.local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)q * 64, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      char buf[200];
      snprintf(buf, sizeof(buf), "mrl_comm_p2p_connect: cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e));
      return mrl_set_error(buf);
    }
    p.peer[q] = (unsigned char*)ptr;
  }
  return 0;
}
// Step 3 (every rank, once all ranks have connected): switch the transport on (or back off).
extern "C" int mrl_comm_p2p_enable(mrl_comm* c, int on) {
  if (!c) return mrl_set_error("mrl_comm_p2p_enable: null communicator");
  if (on) {
    for (int q = 0; q < c->world; ++q)
      if (!c->p2p.peer[q]) return mrl_set_error("mrl_comm_p2p_enable: connect first");
  }
  c->p2p.on = on != 0;
  return 0;
}
bool mrl_comm_p2p_ready(const mrl_comm* c, long long n) { return c && c->world > 1 && c->p2p.on && n <= c->p2p.cap; }

int mrl_comm_p2p_begin(mrl_comm* c, long long n, P2pPush* push) {
  if (!mrl_comm_p2p_ready(c, n)) return mrl_set_error("mrl_comm_p2p_begin: transport not ready");
  P2pState& p = c->p2p;
  p.seq += 1;
  const int parity = (int)(p.seq & 1);
  for (int q = 0; q < MRL_P2P_MAX_WORLD; ++q) push->flag[q] = nullptr;
  for (int q = 0; q < c->world; ++q) push->flag[q] = p2p_flags(p.peer[q], parity) + c->rank;
  push->own = p2p_vector(p.local, p.cap, parity);
  push->counter = p.counter;
  push->seq = p.seq;
  push->world = c->world;
  return 0;
}

// Wait until every rank's flag of this parity shows `seq`, then out[i] = sum_r vector_r[i] in rank order.
// Double buffering by parity is enough: rank r can only overwrite its vector of operation seq (in seq+2) after it has
// seen every rank's flag for seq+1, which a rank raises after it finished reading the vectors of seq.
__global__ void p2p_gather_kernel(P2pGather ga, long long n, double* __restrict__ out64, float* __restrict__ out32) {
  p2p_wait_flags(ga, threadIdx.x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double s = p2p_gather_sum(ga, i);
    if (out64) out64[i] = s;
    if (out32) out32[i] = (float)s;
  }
}
int mrl_comm_p2p_pending(mrl_comm* c, P2pGather* out) {
  if (!c || !c->p2p.on) return mrl_set_error("mrl_comm_p2p_pending: transport not ready");
  P2pState& p = c->p2p;
  const int parity = (int)(p.seq & 1);
  out->flags = p2p_flags(p.local, parity);
  for (int q = 0; q < MRL_P2P_MAX_WORLD; ++q) out->src[q] = q < c->world ? p2p_vector(p.peer[q], p.cap, parity) : nullptr;
  out->seq = p.seq;
  out->timeout_ns = p.timeout_ns;
  out->err = p.d_err;
  out->world = c->world;
  return 0;
}
int mrl_comm_p2p_finish(mrl_comm* c, long long n, double* out64, float* out32, cudaStream_t st) {
  P2pGather ga;
  if (mrl_comm_p2p_pending(c, &ga)) return 1;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 64) blocks = 64;     // all CTAs are resident: they poll the flags
  p2p_gather_kernel<<<blocks, 256, 0, st>>>(ga, n, out64, out32);
  if (cudaGetLastError() != cudaSuccess) return mrl_set_error("p2p_gather_kernel launch failed");
  return 0;
}
int mrl_comm_p2p_error(const mrl_comm* c) {
  if (!c || !c->p2p.h_err || *c->p2p.h_err == 0) return 0;
  return mrl_set_error("peer-memory all-reduce: a rank did not deliver within MRL_P2P_TIMEOUT_S; the results of this update are invalid");
}
__global__ void p2p_push_kernel(const double* __restrict__ in, long long n, P2pPush push) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p2p_push_value(push, i, in[i]);
  p2p_push_done(push);
}

extern "C" int mrl_comm_allreduce_f64(mrl_comm* c, double* buf, long long n, void* stream) {
  if (!c || c->world == 1) return 0;
  if (mrl_comm_p2p_ready(c, n)) {
    P2pPush push;
    if (mrl_comm_p2p_begin(c, n, &push)) return 1;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 64) blocks = 64;
    p2p_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(buf, n, push);
    if (cudaGetLastError() != cudaSuccess) return mrl_set_error("p2p_push_kernel launch failed");
    return mrl_comm_p2p_finish(c, n, buf, nullptr, (cudaStream_t)stream);
  }
  int rc = g_nccl.AllReduce(buf, buf, (size_t)n, NCCL_FLOAT64, NCCL_SUM, c->comm, (cudaStream_t)stream);
  if (rc != NCCL_SUCCESS) return nccl_fail("ncclAllReduce", rc);
  return 0;
}
