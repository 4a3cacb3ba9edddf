// Multi-GPU plumbing of the C ABI: every batch item is independent, so the only exchange on this
// path is ONE all-gather of the per-device result slices.  For hosts that do not use torch.distributed (the Python
// package does: forest_benchmarking_b200/sharding.py) this file wraps NCCL's single-process API:
//     ncclCommInitAll over the box's devices, ncclAllGather inside a group call, ncclCommDestroy.
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy already loaded by the host process, e.g. torch's, or the
// system one), so libqtomo.so itself has no link-time dependency on it.
#include "qt_common.cuh"
#include "../../include/qtomo.h"

#include <dlfcn.h>
#include <vector>

namespace {
typedef void* nccl_comm_t;
struct NcclApi {
  void* handle = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

int load_nccl(NcclApi& api) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    qt_set_error("qt_comm_init_all: cannot load libnccl.so.2 (%s)", dlerror());
    return QT_ERR_UNSUPPORTED;
  }
#define BIND(field, sym)                                                       \
  *(void**)(&api.field) = dlsym(api.handle, sym);                              \
  if (!api.field) {                                                            \
    qt_set_error("qt_comm_init_all: libnccl has no symbol %s", sym);           \
    return QT_ERR_UNSUPPORTED;                                                 \
  }
  BIND(CommInitAll, "ncclCommInitAll")
  BIND(AllGather, "ncclAllGather")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  return QT_OK;
}
}  // namespace

struct qt_comm {
  NcclApi api;
  std::vector<int> devices;
  std::vector<nccl_comm_t> comms;
};

#define QT_NCCL(c, call)                                                                   \
  do {                                                                                     \
    const int r_ = (call);                                                                 \
    if (r_ != 0) {                                                                         \
      qt_set_error("%s failed: %s", #call, (c)->api.GetErrorString(r_));                   \
      return QT_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

extern "C" int qt_comm_init_all(int ndev, const int32_t* devices, qt_comm** comm_out) {
  QT_REQUIRE(ndev >= 1 && devices && comm_out, "qt_comm_init_all: bad arguments");
  qt_comm* c = new qt_comm();
  int rc = load_nccl(c->api);
  if (rc != QT_OK) {
    delete c;
    return rc;
  }
  c->devices.assign(devices, devices + ndev);
  c->comms.assign(ndev, nullptr);
  const int r = c->api.CommInitAll(c->comms.data(), ndev, c->devices.data());
  if (r != 0) {
    qt_set_error("ncclCommInitAll failed: %s", c->api.GetErrorString(r));
    delete c;
    return QT_ERR_CUDA;
  }
  *comm_out = c;
  return QT_OK;
}

extern "C" int qt_allgather_bytes(qt_comm* c, const void* const* sendbufs, void* const* recvbufs,
                                  int64_t nbytes_per_rank, void* const* streams) {
  QT_REQUIRE(c && sendbufs && recvbufs && nbytes_per_rank >= 0, "qt_allgather_bytes: bad arguments");
  if (nbytes_per_rank == 0) return QT_OK;
  int prev = 0;
  QT_CUDA(cudaGetDevice(&prev));
  QT_NCCL(c, c->api.GroupStart());
  for (size_t r = 0; r < c->comms.size(); ++r) {
    QT_CUDA(cudaSetDevice(c->devices[r]));
    const int rc = c->api.AllGather(sendbufs[r], recvbufs[r], (size_t)nbytes_per_rank, /*ncclInt8*/ 0, c->comms[r],
                                    streams ? (cudaStream_t)streams[r] : (cudaStream_t)0);
    if (rc != 0) {
      c->api.GroupEnd();
      cudaSetDevice(prev);
      qt_set_error("ncclAllGather failed on rank %d: %s", (int)r, c->api.GetErrorString(rc));
      return QT_ERR_CUDA;
    }
  }
  QT_NCCL(c, c->api.GroupEnd());
  QT_CUDA(cudaSetDevice(prev));
  return QT_OK;
}

extern "C" int qt_comm_destroy(qt_comm* c) {
  if (!c) return QT_OK;
  for (nccl_comm_t k : c->comms)
    if (k) c->api.CommDestroy(k);
  delete c;
  return QT_OK;
}
