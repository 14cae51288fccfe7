// Gradient all-reduce of the data-parallel path (the reference is single-process:
// no counterpart) through NCCL, bound at RUN TIME: the library is dlopen'ed -- preferably the
// copy the host process has already loaded (PyTorch bundles one), so that no second NCCL
// instance is created -- and only four of its entry points are used.  Everything a maintainer
// needs to drive the collective from another binding is in these exports:
//
//   rank 0:   sp_nccl_unique_id(id)            -> ship the 128 bytes to every rank (any transport)
//   all:      sp_nccl_init(id, nranks, rank, &comm)
//   per step: sp_allreduce_sum_f32(comm, bucket, count, stream)     (in place, stream-ordered,
//                                                                    capturable in a CUDA graph)
//   at exit:  sp_nccl_destroy(comm)
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace {
struct NcclUid {
  char internal[128];
};
using GetUniqueIdFn = int (*)(NcclUid*);
using CommInitRankFn = int (*)(void**, int, NcclUid, int);
using AllReduceFn = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using CommDestroyFn = int (*)(void*);
using GetErrorStringFn = const char* (*)(int);

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  bool ok = false;
};
Nccl g_nccl;
std::mutex g_nccl_mutex;

bool bind(void* h) {
  Nccl n;
  n.handle = h;
  n.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
  n.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
  n.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
  n.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
  n.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
  n.ok = n.get_unique_id && n.comm_init_rank && n.all_reduce && n.comm_destroy;
  if (n.ok) g_nccl = n;
  return n.ok;
}

int nccl_fail(const char* what, int rc) {
  const char* text = g_nccl.error_string ? g_nccl.error_string(rc) : "?";
  return sp::set_error(SP_ERR_DRIVER, "%s failed: NCCL error %d (%s)", what, rc, text);
}
}  // namespace

extern "C" {

int sp_nccl_load(const char* path) {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.ok) return 0;
  const char* env = getenv("SP_NCCL_LIB");
  // an already-loaded copy first (RTLD_NOLOAD): never two NCCL instances in one process
  const char* cands[4] = {env, path, "libnccl.so.2", "libnccl.so"};
  for (int pass = 0; pass < 2; ++pass) {
    for (const char* c : cands) {
      if (!c || !*c) continue;
      void* h = dlopen(c, RTLD_NOW | RTLD_GLOBAL | (pass == 0 ? RTLD_NOLOAD : 0));
      if (h && bind(h)) return 0;
    }
  }
  return sp::set_error(SP_ERR_UNSUPPORTED, "NCCL library not found (set SP_NCCL_LIB): %s",
                       dlerror() ? dlerror() : "no candidate could be opened");
}

int sp_nccl_unique_id(void* out128) {
  SP_REQUIRE(out128 != nullptr, "null output");
  if (!g_nccl.ok)
    if (int rc = sp_nccl_load(nullptr)) return rc;
  NcclUid id;
  if (int rc = g_nccl.get_unique_id(&id)) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int sp_nccl_init(const void* unique_id128, int nranks, int rank, void** comm_out) {
  SP_REQUIRE(unique_id128 && comm_out, "null argument");
  SP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / world size");
  if (!g_nccl.ok)
    if (int rc = sp_nccl_load(nullptr)) return rc;
  NcclUid id;
  memcpy(&id, unique_id128, sizeof(id));
  void* comm = nullptr;
  if (int rc = g_nccl.comm_init_rank(&comm, nranks, id, rank)) return nccl_fail("ncclCommInitRank", rc);
  *comm_out = comm;
  return 0;
}

int sp_allreduce_sum_f32(void* comm, float* buf, int64_t count, sp_stream_t stream) {
  SP_REQUIRE(comm != nullptr && g_nccl.ok, "communicator not initialised (sp_nccl_init)");
  SP_REQUIRE(count >= 0, "negative count");
  if (count == 0) return 0;
  // ncclFloat32 = 7, ncclSum = 0
  if (int rc = g_nccl.all_reduce(buf, buf, (size_t)count, 7, 0, comm, sp::as_stream(stream)))
    return nccl_fail("ncclAllReduce", rc);
  return 0;
}

int sp_nccl_destroy(void* comm) {
  if (!comm || !g_nccl.ok) return 0;
  if (int rc = g_nccl.comm_destroy(comm)) return nccl_fail("ncclCommDestroy", rc);
  return 0;
}

}  // extern "C"
