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

struct mrl_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
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
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}
extern "C" int mrl_comm_allreduce_f64(mrl_comm* c, double* buf, long long n, void* stream) {
  if (!c || c->world == 1) return 0;
  int rc = g_nccl.AllReduce(buf, buf, (size_t)n, NCCL_FLOAT64, NCCL_SUM, c->comm, (cudaStream_t)stream);
  if (rc != NCCL_SUCCESS) return nccl_fail("ncclAllReduce", rc);
  return 0;
}
