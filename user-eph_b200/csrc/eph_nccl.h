// The slice of NCCL the engine uses for its multi-rank data plane, bound at run time.
//
// libeph_b200.so has no link-time dependency on NCCL: a single-rank run needs none, and inside a process that has
// already loaded a copy (PyTorch bundles one) the loader hands back that very copy for the soname, so one communicator
// library serves the whole process.  Only eph_b200_comm_* touches this; a missing library is an error there, not a
// silent fall-back.  Types and enumerator values are NCCL's public ABI (nccl.h: ncclUniqueId is 128 opaque bytes,
// ncclFloat64 = 8, ncclSum = 0, ncclSuccess = 0).
#pragma once

#include <cuda_runtime.h>

#ifndef EPHA_HOST_EMULATION
#include <dlfcn.h>
#endif

#include <string>

namespace ephb {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm *NcclComm;
enum { kNcclSuccess = 0, kNcclInt8 = 0, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string error;   // why loading failed
  bool ok = false;
};

#ifdef EPHA_HOST_EMULATION
// tests/emul: the host build links a stand-in that moves the bytes through the test's own transport (fake_nccl.cpp)
extern "C" {
int ncclGetUniqueId(NcclUniqueId *);
int ncclCommInitRank(NcclComm *, int, NcclUniqueId, int);
int ncclCommDestroy(NcclComm);
int ncclGroupStart();
int ncclGroupEnd();
int ncclSend(const void *, size_t, int, int, NcclComm, cudaStream_t);
int ncclRecv(void *, size_t, int, int, NcclComm, cudaStream_t);
int ncclAllReduce(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
int ncclAllGather(const void *, void *, size_t, int, NcclComm, cudaStream_t);
int ncclReduceScatter(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
const char *ncclGetErrorString(int);
}
#endif

inline NcclApi &nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
#ifdef EPHA_HOST_EMULATION
  api.GetUniqueId = ncclGetUniqueId; api.CommInitRank = ncclCommInitRank; api.CommDestroy = ncclCommDestroy;
  api.GroupStart = ncclGroupStart; api.GroupEnd = ncclGroupEnd; api.Send = ncclSend; api.Recv = ncclRecv;
  api.AllReduce = ncclAllReduce; api.AllGather = ncclAllGather; api.ReduceScatter = ncclReduceScatter; api.GetErrorString = ncclGetErrorString;
  api.ok = true;
#else
  void *lib = nullptr;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    const char *e = dlerror();
    api.error = std::string("cannot load libnccl.so.2: ") + (e ? e : "not found");
    return api;
  }
  bool all = true;
  auto bind = [&](auto &fn, const char *sym) {
    fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(lib, sym));
    if (!fn) { all = false; api.error = std::string("libnccl lacks ") + sym; }
  };
  bind(api.GetUniqueId, "ncclGetUniqueId"); bind(api.CommInitRank, "ncclCommInitRank"); bind(api.CommDestroy, "ncclCommDestroy");
  bind(api.GroupStart, "ncclGroupStart"); bind(api.GroupEnd, "ncclGroupEnd"); bind(api.Send, "ncclSend"); bind(api.Recv, "ncclRecv");
  bind(api.AllReduce, "ncclAllReduce"); bind(api.AllGather, "ncclAllGather"); bind(api.ReduceScatter, "ncclReduceScatter");
  bind(api.GetErrorString, "ncclGetErrorString");
  api.ok = all;
#endif
  return api;
}

}  // namespace ephb
