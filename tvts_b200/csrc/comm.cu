// tvts_comm_*: the two exchange steps of the data-parallel path as C-ABI entry points over NCCL --
//   all-gather of the [B_local, 2E] embeddings   (AllGather_multi.forward, v2/trainer/trainer.py:41-57, :481-482)
//   all-reduce (average) of the gradient arena   (DistributedDataParallel's gradient averaging, v2/base/base_trainer.py:23-25)
// One communicator per process (one process per GPU).  NCCL is resolved at run time: the library the process has already loaded (torch
// ships its own libnccl.so.2) is reused, otherwise the system one is opened -- no link-time dependency, and nothing happens unless
// tvts_comm_* is called.  Every collective is enqueued on the caller's stream (capturable in a CUDA graph), nothing synchronises.
#include <dlfcn.h>
#include <nccl.h>
#include "common.cuh"
#include "../../include/tvts_b200.h"

struct tvts_comm {
  ncclComm_t comm;
  int rank, world;
};

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.handle) return TVTS_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy this process already uses (torch's), if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return tvts_set_error(TVTS_ERR_UNSUPPORTED, "tvts_comm: libnccl.so.2 not found (%s)", dlerror());
#define TVTS_NCCL_SYM(field, name)                                                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                                     \
  if (!g_nccl.field) return tvts_set_error(TVTS_ERR_UNSUPPORTED, "tvts_comm: symbol %s missing in libnccl", name)
  TVTS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  TVTS_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  TVTS_NCCL_SYM(AllGather, "ncclAllGather");
  TVTS_NCCL_SYM(AllReduce, "ncclAllReduce");
  TVTS_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  TVTS_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef TVTS_NCCL_SYM
  g_nccl.handle = h;
  return TVTS_OK;
}

#define TVTS_CHECK_NCCL(expr)                                                                                   \
  do {                                                                                                          \
    const ncclResult_t r_ = (expr);                                                                             \
    if (r_ != ncclSuccess) return tvts_set_error(TVTS_ERR_CUDA, "%s -> %s", #expr, g_nccl.GetErrorString(r_)); \
  } while (0)

}  // namespace

static_assert(sizeof(ncclUniqueId) == TVTS_COMM_ID_BYTES, "tvts_comm: NCCL unique id size");

extern "C" int tvts_comm_unique_id(void* id_out) {
  TVTS_REQUIRE(id_out != nullptr, "tvts_comm_unique_id: null pointer");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  TVTS_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return TVTS_OK;
}

extern "C" int tvts_comm_init(tvts_comm** out, const void* id_bytes, int64_t rank, int64_t world) {
  TVTS_REQUIRE(out != nullptr && id_bytes != nullptr, "tvts_comm_init: null pointer");
  TVTS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "tvts_comm_init: rank %lld of %lld", (long long)rank, (long long)world);
  *out = nullptr;
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t c;
  TVTS_CHECK_NCCL(g_nccl.CommInitRank(&c, (int)world, id, (int)rank));       // binds to the calling thread's current CUDA device
  tvts_comm* h = new tvts_comm{c, (int)rank, (int)world};
  *out = h;
  return TVTS_OK;
}

extern "C" int tvts_comm_allgather(tvts_comm* c, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
  TVTS_REQUIRE(c != nullptr && send != nullptr && recv != nullptr, "tvts_comm_allgather: null pointer");
  TVTS_REQUIRE(bytes_per_rank > 0, "tvts_comm_allgather: bytes_per_rank=%lld", (long long)bytes_per_rank);
  TVTS_CHECK_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, ncclChar, c->comm, reinterpret_cast<cudaStream_t>(stream)));
  return TVTS_OK;
}

extern "C" int tvts_comm_allreduce(tvts_comm* c, float* buf, int64_t n, int64_t average, void* stream) {
  TVTS_REQUIRE(c != nullptr && buf != nullptr, "tvts_comm_allreduce: null pointer");
  TVTS_REQUIRE(n > 0, "tvts_comm_allreduce: n=%lld", (long long)n);
  TVTS_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, average ? ncclAvg : ncclSum, c->comm, reinterpret_cast<cudaStream_t>(stream)));
  return TVTS_OK;
}

extern "C" int tvts_comm_destroy(tvts_comm* c) {
  if (c == nullptr) return TVTS_OK;
  const ncclResult_t r = g_nccl.handle ? g_nccl.CommDestroy(c->comm) : ncclSuccess;
  delete c;
  if (r != ncclSuccess) return tvts_set_error(TVTS_ERR_CUDA, "ncclCommDestroy -> %s", g_nccl.GetErrorString(r));
  return TVTS_OK;
}
