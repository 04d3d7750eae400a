// NCCL communicator for the batch-sharded CTRL-SAC update (one process per GPU).
//
// libnccl is resolved at run time with dlopen so that the library has no link-time dependency on it: single-GPU users
// never touch NCCL, and inside a PyTorch process the already loaded libnccl.so.2 (torch's bundled copy) is the one that
// answers.  Only the seven entry points below are used; their prototypes are the stable NCCL 2.x C API (nccl.h).
#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace rlrep {

namespace {

using ncclComm_t = void*;
struct ncclUniqueId {
  char internal[128];
};
constexpr int kNcclSuccess = 0;
constexpr int kNcclFloat32 = 7;  // ncclDataType_t::ncclFloat32
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      err = std::string("libnccl.so.2 not found: ") + dlerror();
      return;
    }
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p && err.empty()) err = std::string("NCCL symbol missing: ") + n;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  });
  if (!err.empty()) throw Error(err);
  return api;
}

void check(int rc, const char* what) {
  if (rc != kNcclSuccess) throw Error(std::string(what) + " failed: " + nccl().GetErrorString(rc));
}

}  // namespace

void Comm::unique_id(unsigned char out[kUniqueIdBytes]) {
  static_assert(sizeof(ncclUniqueId) == kUniqueIdBytes, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out, id.internal, kUniqueIdBytes);
}

Comm::Comm(const unsigned char id_bytes[kUniqueIdBytes], int rank_, int world_) : rank(rank_), world(world_) {
  RLREP_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank / world size");
  ncclUniqueId id;
  std::memcpy(id.internal, id_bytes, kUniqueIdBytes);
  ncclComm_t c = nullptr;
  check(nccl().CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  comm_ = c;
}

Comm::~Comm() {
  if (comm_) nccl().CommDestroy(comm_);
}

int Comm::version() {
  int v = 0;
  check(nccl().GetVersion(&v), "ncclGetVersion");
  return v;
}

void Comm::all_gather(const float* send, float* recv, size_t count_per_rank, cudaStream_t s) {
  check(nccl().AllGather(send, recv, count_per_rank, kNcclFloat32, comm_, s), "ncclAllGather");
  ++collectives;
  // profile entries named nccl_* are NCCL's kernels, not this library's; bytes = what this rank receives over NVLink
  note_launch("nccl_all_gather", s, 4.0 * (double)count_per_rank * (world - 1), 0.0);
}
void Comm::reduce_scatter(const float* send, float* recv, size_t recv_count, cudaStream_t s) {
  check(nccl().ReduceScatter(send, recv, recv_count, kNcclFloat32, kNcclSum, comm_, s), "ncclReduceScatter");
  ++collectives;
  note_launch("nccl_reduce_scatter", s, 4.0 * (double)recv_count * (world - 1), 0.0);
}
void Comm::all_reduce(float* buf, size_t count, cudaStream_t s) {
  check(nccl().AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, comm_, s), "ncclAllReduce");
  ++collectives;
  note_launch("nccl_all_reduce", s, 2.0 * 4.0 * (double)count * (world - 1) / world, 0.0);
}

}  // namespace rlrep
