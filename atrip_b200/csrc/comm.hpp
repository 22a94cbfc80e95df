// NCCL binding of the engine: point-to-point slice exchange and the final energy all-reduce.
//
// Replaces the reference's MPI layer on the hot path: MPI_Isend / MPI_Irecv / MPI_Wait per slice
// (SliceUnion.cxx:456-462, 491-503; Slice.cxx:167) and MPI_Reduce of the energies
// (Atrip.cxx:1094-1107).  libnccl.so.2 is resolved with dlopen at the first use, so the C-ABI
// library itself loads on a machine without NCCL (and inside a process that already loaded
// PyTorch's bundled NCCL the same copy is reused: dlopen matches the SONAME).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace ab {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string error;
  bool ok = false;
};

inline NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
    return api;
  }
  bool all = true;
  auto sym = [&](const char *name) {
    void *p = dlsym(h, name);
    if (!p) {
      all = false;
      api.error = std::string("libnccl lacks ") + name;
    }
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  api.ok = all;
  return api;
}

}  // namespace ab
