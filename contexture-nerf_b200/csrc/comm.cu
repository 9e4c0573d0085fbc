// Gradient all-reduce of the ray-sharded training step behind the C-ABI.
//
// The path has exactly one exchange step: the sum of the flat fp32 gradient bucket over ranks (it replaces the
// reduce-add of nn.DataParallel, /root/reference/src/training/trainer.py:134-135).  NCCL does it over NVLink; this file
// is the thin binding that lets a non-torch host (and the captured step graph) issue it on the step's own stream:
// the NCCL shared object is opened at run time (dlopen -- no link-time dependency, so libctxnerf.so loads on a box
// without NCCL and every other entry point keeps working), the communicator is an opaque handle owned by the caller,
// and the library keeps no state besides the resolved function table.
//
// The NCCL types are restated from its public header (nccl.h of NCCL 2.x: ncclUniqueId = 128 opaque bytes,
// ncclFloat32 = 7, ncclSum = 0); ctx_comm_load checks the major version.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "ctx_common.cuh"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_get_version)(int*);
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*fn_error_string)(int);

struct NcclApi {
  void* handle = nullptr;
  int version = 0;
  fn_get_version get_version = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_error_string error_string = nullptr;
  char last_error[256] = {0};
};
NcclApi g_nccl;

constexpr int kNcclFloat32 = 7;
constexpr int kNcclSum = 0;

int nccl_fail(int r) {
  if (r == 0) return 0;
  const char* s = g_nccl.error_string ? g_nccl.error_string(r) : "unknown NCCL error";
  snprintf(g_nccl.last_error, sizeof(g_nccl.last_error), "NCCL error %d: %s", r, s);
  return CTX_ERR_NO_NCCL;
}

}  // namespace

// Resolve the NCCL entry points.  `path` nullable: then libnccl.so.2 by soname (the copy torch has already mapped
// into the process, if any).  Returns 0, or CTX_ERR_NO_NCCL when the library or a symbol is missing.
extern "C" int ctx_comm_load(const char* path) {
  if (g_nccl.handle) return 0;
  void* h = dlopen(path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    snprintf(g_nccl.last_error, sizeof(g_nccl.last_error), "dlopen failed: %s", dlerror());
    return CTX_ERR_NO_NCCL;
  }
  NcclApi a;
  a.get_version = (fn_get_version)dlsym(h, "ncclGetVersion");
  a.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
  a.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
  a.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
  a.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
  a.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
  if (!a.get_version || !a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce) {
    snprintf(g_nccl.last_error, sizeof(g_nccl.last_error), "NCCL library lacks a required symbol");
    dlclose(h);
    return CTX_ERR_NO_NCCL;
  }
  int v = 0;
  if (a.get_version(&v) != 0 || v < 20000 || v >= 30000) {   // the restated ABI is that of NCCL 2.x
    snprintf(g_nccl.last_error, sizeof(g_nccl.last_error), "unsupported NCCL version code %d (need 2.x)", v);
    dlclose(h);
    return CTX_ERR_NO_NCCL;
  }
  a.handle = h;
  a.version = v;
  g_nccl = a;
  return 0;
}

// NCCL version code (e.g. 22809) once loaded, else 0.
extern "C" int ctx_comm_version(void) { return g_nccl.handle ? g_nccl.version : 0; }

// Text of the last communicator failure (host string, never NULL).
extern "C" const char* ctx_comm_last_error(void) { return g_nccl.last_error; }

// 128-byte rendezvous token: made on one rank, handed to the others by the host (any out-of-band channel).
extern "C" int ctx_comm_unique_id(void* id_out_host) {
  if (!id_out_host) return CTX_ERR_BAD_ARG;
  if (!g_nccl.handle) return CTX_ERR_NO_NCCL;
  NcclUniqueId id;
  memset(&id, 0, sizeof(id));
  if (int r = nccl_fail(g_nccl.get_unique_id(&id))) return r;
  memcpy(id_out_host, &id, sizeof(id));
  return 0;
}

// Collective over all ranks: creates this rank's communicator on the CURRENT device.  *comm_out = opaque handle.
extern "C" int ctx_comm_init(void** comm_out, int n_ranks, const void* id_host, int rank) {
  if (!comm_out || !id_host || n_ranks < 1 || rank < 0 || rank >= n_ranks) return CTX_ERR_BAD_ARG;
  if (!g_nccl.handle) return CTX_ERR_NO_NCCL;
  NcclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  NcclComm c = nullptr;
  if (int r = nccl_fail(g_nccl.comm_init_rank(&c, n_ranks, id, rank))) return r;
  *comm_out = c;
  return 0;
}

extern "C" int ctx_comm_destroy(void* comm) {
  if (!comm) return 0;
  if (!g_nccl.handle) return CTX_ERR_NO_NCCL;
  return nccl_fail(g_nccl.comm_destroy((NcclComm)comm));
}

// In-place sum of bucket[0..n) over the ranks of `comm`, enqueued on `stream` (capturable into a CUDA graph: every
// rank captures the same sequence).  The 1/N of the mean is folded into ctx_adam_step's grad_scale.
extern "C" int ctx_allreduce(void* comm, float* bucket, int64_t n, void* stream) {
  if (!comm || n < 0 || (n > 0 && !bucket)) return CTX_ERR_BAD_ARG;
  if (!g_nccl.handle) return CTX_ERR_NO_NCCL;
  if (n == 0) return 0;
  return nccl_fail(g_nccl.all_reduce(bucket, bucket, (size_t)n, kNcclFloat32, kNcclSum, (NcclComm)comm,
                                     (cudaStream_t)stream));
}
