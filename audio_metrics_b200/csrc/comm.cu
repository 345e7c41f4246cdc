// amb_comm_*: NCCL communicators for a C caller's sharded path.
//
// The sweeps shard by rows with small exchanges in between (allgather of radii slices, allreduce of
// per-candidate counts, allreduce of fp64 moments); the Python mirror does those with
// torch.distributed.  A C caller — and amb_host_evaluate on several devices — gets the same
// collectives here: one communicator per device of THIS process (ncclCommInitAll; the reference's
// only multi-GPU model is threads of one process, util/gpu_parallel.py:20-76), asynchronous on the
// caller's streams, device pointers in and out.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library keeps loading on machines
// without NCCL, a process that already holds an NCCL (PyTorch's) shares it, and the first
// amb_comm_init fails loudly where there is none.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "internal.cuh"

// The slice of nccl.h this file needs (NCCL 2.x ABI: opaque communicator, C enums), declared here so
// that building the library does not need NCCL's headers either.
typedef struct ncclComm* ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;

namespace amb {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static const NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    auto sym = [&](const char* s) { return dlsym(api.lib, s); };
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    if (!api.CommInitAll || !api.CommDestroy || !api.AllReduce || !api.AllGather || !api.GroupStart || !api.GroupEnd ||
        !api.GetErrorString) {
      dlclose(api.lib);
      api.lib = nullptr;
    }
  });
  return api.lib ? &api : nullptr;
}

static int check_nccl(const NcclApi* a, ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return AMB_OK;
  return set_error(AMB_ERR_CUDA, "%s: NCCL error %d (%s)", what, static_cast<int>(r), a->GetErrorString(r));
}

static bool nccl_type(int dtype, ncclDataType_t* t) {
  switch (dtype) {
    case AMB_F32: *t = ncclFloat32; return true;
    case AMB_F64: *t = ncclFloat64; return true;
    case AMB_I32: *t = ncclInt32; return true;
    case AMB_I64: *t = ncclInt64; return true;
    case AMB_U8: *t = ncclUint8; return true;
    default: return false;
  }
}

}  // namespace amb

using namespace amb;

struct amb_comm {
  std::vector<int> devs;
  std::vector<ncclComm_t> comms;
};

extern "C" {

int amb_comm_init(const int* devs, int n_dev, amb_comm_t** out) {
  if (!devs || n_dev < 1 || n_dev > 64 || !out) return set_error(AMB_ERR_ARG, "amb_comm_init: bad argument");
  const NcclApi* a = nccl_api();
  if (!a) return set_error(AMB_ERR_CUDA, "amb_comm_init: NCCL is not available (dlopen of libnccl.so.2 / libnccl.so failed, or it lacks a symbol)");
  for (int i = 0; i < n_dev; ++i)
    for (int j = 0; j < i; ++j)
      if (devs[i] == devs[j]) return set_error(AMB_ERR_ARG, "amb_comm_init: device %d listed twice", devs[i]);
  amb_comm* c = new amb_comm;
  c->devs.assign(devs, devs + n_dev);
  c->comms.assign(n_dev, nullptr);
  const int rc = check_nccl(a, a->CommInitAll(c->comms.data(), n_dev, devs), "ncclCommInitAll");
  if (rc) {
    delete c;
    return rc;
  }
  *out = c;
  return AMB_OK;
}

int amb_comm_size(const amb_comm_t* c) { return c ? static_cast<int>(c->comms.size()) : 0; }

int amb_comm_device(const amb_comm_t* c, int rank) {
  return (c && rank >= 0 && rank < static_cast<int>(c->devs.size())) ? c->devs[rank] : -1;
}

int amb_comm_group_begin(void) {
  const NcclApi* a = nccl_api();
  if (!a) return set_error(AMB_ERR_CUDA, "amb_comm_group_begin: NCCL not loaded");
  return check_nccl(a, a->GroupStart(), "ncclGroupStart");
}

int amb_comm_group_end(void) {
  const NcclApi* a = nccl_api();
  if (!a) return set_error(AMB_ERR_CUDA, "amb_comm_group_end: NCCL not loaded");
  return check_nccl(a, a->GroupEnd(), "ncclGroupEnd");
}

int amb_comm_allreduce(amb_comm_t* c, int rank, const void* send, void* recv, long long count, int dtype, int op,
                       amb_stream_t stream) {
  ncclDataType_t t;
  if (!c || rank < 0 || rank >= static_cast<int>(c->comms.size()) || !send || !recv || count < 0 || !nccl_type(dtype, &t) ||
      (op != AMB_SUM && op != AMB_MAX))
    return set_error(AMB_ERR_ARG, "amb_comm_allreduce: bad argument");
  const NcclApi* a = nccl_api();
  return check_nccl(a, a->AllReduce(send, recv, static_cast<size_t>(count), t, op == AMB_SUM ? ncclSum : ncclMax,
                                    c->comms[rank], static_cast<cudaStream_t>(stream)), "ncclAllReduce");
}

int amb_comm_allgather(amb_comm_t* c, int rank, const void* send, void* recv, long long count_per_rank, int dtype,
                       amb_stream_t stream) {
  ncclDataType_t t;
  if (!c || rank < 0 || rank >= static_cast<int>(c->comms.size()) || !send || !recv || count_per_rank < 0 ||
      !nccl_type(dtype, &t))
    return set_error(AMB_ERR_ARG, "amb_comm_allgather: bad argument");
  const NcclApi* a = nccl_api();
  return check_nccl(a, a->AllGather(send, recv, static_cast<size_t>(count_per_rank), t, c->comms[rank],
                                    static_cast<cudaStream_t>(stream)), "ncclAllGather");
}

int amb_comm_destroy(amb_comm_t* c) {
  if (!c) return AMB_OK;
  const NcclApi* a = nccl_api();
  int rc = AMB_OK;
  if (a)
    for (ncclComm_t q : c->comms)
      if (q) {
        const int r = check_nccl(a, a->CommDestroy(q), "ncclCommDestroy");
        if (!rc) rc = r;
      }
  delete c;
  return rc;
}

}  // extern "C"
