// NCCL plumbing of the chi-sharded local solve (BASELINE configs[4]): one communicator per process
// (one process per GPU), created from a unique id the host side passes around (torch.distributed broadcast in
// tnpy_b200/parallel.py).  The library does not link NCCL: libnccl.so.2 is resolved at run time -- the copy PyTorch
// already loaded when there is one, the system's otherwise -- so single-GPU users never need it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <new>

#include "comm.cuh"

namespace tnpy {

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy already in the process (PyTorch's)
    if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!api.lib) return;
    auto sym = [&](const char* name) { return dlsym(api.lib, name); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.GetErrorString;
  });
  return api;
}

int nccl_fail(const char* what, ncclResult_t rc) {
  set_error("%s failed: %s", what, nccl_api().GetErrorString ? nccl_api().GetErrorString(rc) : "NCCL error");
  return TNPY_ECUDA;
}
}  // namespace

int comm_allgather(const tnpy_comm* c, const double* send, double* recv, size_t count, cudaStream_t stream) {
  const ncclResult_t rc = nccl_api().AllGather(send, recv, count, ncclDouble, static_cast<ncclComm_t>(c->nccl), stream);
  return rc == ncclSuccess ? TNPY_OK : nccl_fail("ncclAllGather", rc);
}

int comm_allreduce_sum(const tnpy_comm* c, double* buf, size_t count, cudaStream_t stream) {
  const ncclResult_t rc =
      nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->nccl), stream);
  return rc == ncclSuccess ? TNPY_OK : nccl_fail("ncclAllReduce", rc);
}

}  // namespace tnpy

using namespace tnpy;

extern "C" int tnpy_comm_unique_id(char* id_out) {
  TNPY_CHECK_ARG(id_out != nullptr, "null pointer");
  NcclApi& api = nccl_api();
  if (!api.ok) {
    set_error("tnpy_comm_unique_id: libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    return TNPY_ECUDA;
  }
  ncclUniqueId id;
  const ncclResult_t rc = api.GetUniqueId(&id);
  if (rc != ncclSuccess) return nccl_fail("ncclGetUniqueId", rc);
  static_assert(sizeof(id.internal) == TNPY_COMM_ID_BYTES, "unique id size");
  memcpy(id_out, id.internal, TNPY_COMM_ID_BYTES);
  return TNPY_OK;
}

extern "C" int tnpy_comm_create(tnpy_comm** comm, const char* id, int world, int rank) {
  TNPY_CHECK_ARG(comm && id && world >= 1 && rank >= 0 && rank < world, "bad argument");
  NcclApi& api = nccl_api();
  if (!api.ok) {
    set_error("tnpy_comm_create: libnccl.so.2 could not be loaded");
    return TNPY_ECUDA;
  }
  ncclUniqueId uid;
  memcpy(uid.internal, id, TNPY_COMM_ID_BYTES);
  ncclComm_t nc = nullptr;
  const ncclResult_t rc = api.CommInitRank(&nc, world, uid, rank);  // collective over the ranks, on the current device
  if (rc != ncclSuccess) return nccl_fail("ncclCommInitRank", rc);
  tnpy_comm* c = new (std::nothrow) tnpy_comm;
  if (!c) {
    api.CommDestroy(nc);
    set_error("tnpy_comm_create: out of host memory");
    return TNPY_EINVAL;
  }
  c->nccl = nc;
  c->world = world;
  c->rank = rank;
  *comm = c;
  return TNPY_OK;
}

extern "C" int tnpy_comm_destroy(tnpy_comm* comm) {
  if (comm) {
    if (comm->nccl) nccl_api().CommDestroy(static_cast<ncclComm_t>(comm->nccl));
    delete comm;
  }
  return TNPY_OK;
}

extern "C" int tnpy_comm_world(const tnpy_comm* comm) { return comm ? comm->world : TNPY_EINVAL; }
extern "C" int tnpy_comm_rank(const tnpy_comm* comm) { return comm ? comm->rank : TNPY_EINVAL; }

// Collectives on device buffers, exposed for the host-side tests of the plumbing and for callers that shard more
// than the local solve: all-gather `count` doubles per rank (recv holds world * count), in-place sum all-reduce.
extern "C" int tnpy_comm_allgather(const tnpy_comm* comm, const double* send, double* recv, int64_t count, void* stream) {
  TNPY_CHECK_ARG(comm && send && recv && count > 0, "bad argument");
  return comm_allgather(comm, send, recv, (size_t)count, static_cast<cudaStream_t>(stream));
}
extern "C" int tnpy_comm_allreduce_sum(const tnpy_comm* comm, double* buf, int64_t count, void* stream) {
  TNPY_CHECK_ARG(comm && buf && count > 0, "bad argument");
  return comm_allreduce_sum(comm, buf, (size_t)count, static_cast<cudaStream_t>(stream));
}
