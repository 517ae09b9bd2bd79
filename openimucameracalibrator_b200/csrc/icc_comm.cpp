// In-library NCCL communicator of the sharded solve ("handle owns ... NCCL communicator").
//
// Residual blocks shard over ranks by time slice; what crosses NVLink per LM iteration is ONE ncclAllReduce(sum, fp64) of the packed
// normal-equation buffer (band | border coupling | border block | gradient | cost: 4.1 MB for BASELINE config 4) issued on the
// solver's own stream right behind the evaluation kernels, plus one 8-byte all-reduce per candidate-cost evaluation.  No Python,
// no host callback in the collective path (round 1 went through a ctypes callback into torch.distributed).
//
// NCCL is bound with dlopen at first use instead of at link time: a process that already carries an NCCL (PyTorch bundles its own
// 2.28 next to the system's 2.27) must not end up with two copies, so the already-loaded library is preferred (RTLD_NOLOAD) and the
// system one is the fallback for plain C++ hosts (the drop-in CLI's --gpus mode: one PROCESS per GPU, like every NCCL
// deployment this library was validated in; communicators of several devices inside one process are deliberately not offered --
// host threads that allocate or synchronise while a peer's all-reduce kernel spins can deadlock the process).
#include "../../include/icc_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>

#include <mutex>
#include <string>

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
enum { kNcclSuccess = 0, kNcclDouble = 8, kNcclSum = 0 };   // ncclFloat64 = 8, ncclSum = 0 (nccl.h; stable since NCCL 2.0)

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.lib) break; }     // the process' own NCCL first
    if (!api.lib) for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) { api.error = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed"); return; }
    auto sym = [&](const char* s) { void* p = dlsym(api.lib, s); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s; return p; };
    api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclAllReduce");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  });
  return api;
}

thread_local std::string g_comm_error;

}  // namespace

struct icc_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

extern "C" {

const char* icc_comm_last_error(void) { return g_comm_error.c_str(); }

icc_status icc_comm_unique_id(unsigned char id[ICC_COMM_ID_BYTES]) {
  NcclApi& n = nccl();
  if (!id) return ICC_ERR_INVALID_ARGUMENT;
  if (!n.error.empty()) { g_comm_error = n.error; return ICC_ERR_UNSUPPORTED; }
  NcclUniqueId u;
  const int e = n.GetUniqueId(&u);
  if (e != kNcclSuccess) { g_comm_error = std::string("ncclGetUniqueId: ") + n.GetErrorString(e); return ICC_ERR_CUDA; }
  static_assert(sizeof(NcclUniqueId) == ICC_COMM_ID_BYTES, "unique id size");
  memcpy(id, &u, sizeof u);
  return ICC_OK;
}

icc_status icc_comm_create(icc_comm** out, const unsigned char id[ICC_COMM_ID_BYTES], int rank, int world, int device_ordinal) {
  if (!out || !id || world < 1 || rank < 0 || rank >= world) return ICC_ERR_INVALID_ARGUMENT;
  NcclApi& n = nccl();
  if (!n.error.empty()) { g_comm_error = n.error; return ICC_ERR_UNSUPPORTED; }
  if (cudaSetDevice(device_ordinal) != cudaSuccess) { g_comm_error = "cudaSetDevice failed"; return ICC_ERR_NO_DEVICE; }
  NcclUniqueId u; memcpy(&u, id, sizeof u);
  icc_comm* c = new icc_comm;
  c->rank = rank; c->world = world; c->device = device_ordinal;
  const int e = n.CommInitRank(&c->comm, world, u, rank);
  if (e != kNcclSuccess) { g_comm_error = std::string("ncclCommInitRank: ") + n.GetErrorString(e); delete c; return ICC_ERR_CUDA; }
  *out = c;
  return ICC_OK;
}

void icc_comm_destroy(icc_comm* c) {
  if (!c) return;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

int icc_comm_rank(const icc_comm* c) { return c ? c->rank : 0; }
int icc_comm_world(const icc_comm* c) { return c ? c->world : 1; }
int icc_comm_nccl_version(void) { NcclApi& n = nccl(); int v = 0; if (n.error.empty() && n.GetVersion) n.GetVersion(&v); return v; }

// in-place sum of n doubles on `stream` (internal to libicc_b200.so; used by icc_api.cu)
int icc_comm_allreduce_sum(icc_comm* c, double* dev, int64_t n, void* stream) {
  if (!c || !c->comm) return 1;
  const int e = nccl().AllReduce(dev, dev, (size_t)n, kNcclDouble, kNcclSum, c->comm, (cudaStream_t)stream);
  if (e != kNcclSuccess) { g_comm_error = std::string("ncclAllReduce: ") + nccl().GetErrorString(e); return 1; }
  return 0;
}

}  // extern "C"
