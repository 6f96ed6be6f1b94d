// Data-parallel gradient exchange for the MLP training step (BASELINE config 4).  The reference
// has no distributed code at all; this is new: one process per GPU, NCCL over
// NVLink 5 / NVSwitch, all-reduce issued on a dedicated comm stream that is ordered after the
// compute stream by an event so the exchange of late-layer gradients overlaps the remaining
// backward GEMMs.  libnccl is dlopen()ed (path supplied by the host shim: the torch-bundled
// libnccl.so.2) so the library itself loads on machines without NCCL.
#include <dlfcn.h>
#include <string.h>

#include "mdb_common.cuh"

namespace mdb {

struct NcclUniqueId { char internal[128]; };
typedef void* ncclComm_t;
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_Group)(void);
typedef const char* (*fn_GetErrorString)(int);

static void* g_nccl = nullptr;
static fn_GetUniqueId p_GetUniqueId;
static fn_CommInitRank p_CommInitRank;
static fn_AllReduce p_AllReduce;
static fn_CommDestroy p_CommDestroy;
static fn_Group p_GroupStart, p_GroupEnd;
static fn_GetErrorString p_GetErrorString;
static ncclComm_t g_comm = nullptr;
static cudaStream_t g_comm_stream = nullptr;
static cudaEvent_t g_ev_compute = nullptr, g_ev_comm = nullptr;
// one completion event per all-reduce (ring): the optimiser can wait for ONE parameter's exchange and
// update it while the exchange of later gradients is still in flight
constexpr int kSeqRing = 64;
static cudaEvent_t g_ev_seq[kSeqRing] = {};
static uint64_t g_seq = 0;
static int g_world = 1;

static int load_nccl(const char* path) {
  if (g_nccl) return 0;
  const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* c : cands) {
    if (!c || !*c) continue;
    g_nccl = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl) break;
  }
  if (!g_nccl) return set_error(MDB_ECOMM, "cannot dlopen libnccl (%s)", dlerror());
  p_GetUniqueId = (fn_GetUniqueId)dlsym(g_nccl, "ncclGetUniqueId");
  p_CommInitRank = (fn_CommInitRank)dlsym(g_nccl, "ncclCommInitRank");
  p_AllReduce = (fn_AllReduce)dlsym(g_nccl, "ncclAllReduce");
  p_CommDestroy = (fn_CommDestroy)dlsym(g_nccl, "ncclCommDestroy");
  p_GetErrorString = (fn_GetErrorString)dlsym(g_nccl, "ncclGetErrorString");
  p_GroupStart = (fn_Group)dlsym(g_nccl, "ncclGroupStart");
  p_GroupEnd = (fn_Group)dlsym(g_nccl, "ncclGroupEnd");
  if (!p_GroupStart || !p_GroupEnd) return set_error(MDB_ECOMM, "libnccl is missing ncclGroupStart/End");
  if (!p_GetUniqueId || !p_CommInitRank || !p_AllReduce || !p_CommDestroy || !p_GetErrorString)
    return set_error(MDB_ECOMM, "libnccl is missing required symbols");
  return 0;
}
#define MDB_NCCL(call)                                                                          \
  do {                                                                                          \
    int r__ = (call);                                                                           \
    if (r__ != 0) return set_error(MDB_ECOMM, "%s -> %s", #call, p_GetErrorString(r__));        \
  } while (0)

}  // namespace mdb

using namespace mdb;

extern "C" {

int mdb_comm_unique_id(void* id128, const char* nccl_lib_path) {
  MDB_TRY(load_nccl(nccl_lib_path));
  NcclUniqueId id;
  MDB_NCCL(p_GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int mdb_comm_init(int rank, int world, const void* id128, const char* nccl_lib_path) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
  MDB_REQUIRE(g_comm == nullptr, "communicator already initialised");
  MDB_TRY(load_nccl(nccl_lib_path));
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  MDB_CUDA(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
  MDB_CUDA(cudaEventCreateWithFlags(&g_ev_compute, cudaEventDisableTiming));
  MDB_CUDA(cudaEventCreateWithFlags(&g_ev_comm, cudaEventDisableTiming));
  for (int i = 0; i < kSeqRing; ++i) MDB_CUDA(cudaEventCreateWithFlags(&g_ev_seq[i], cudaEventDisableTiming));
  g_seq = 0;
  MDB_NCCL(p_CommInitRank(&g_comm, world, id, rank));
  g_world = world;
  return 0;
}

int mdb_comm_allreduce_f32(void* ptr, size_t count, int average) {
  MDB_REQUIRE(g_comm != nullptr, "communicator not initialised");
  // comm stream waits for everything enqueued on the compute stream so far (the producer of ptr)
  MDB_CUDA(cudaEventRecord(g_ev_compute, g_stream));
  MDB_CUDA(cudaStreamWaitEvent(g_comm_stream, g_ev_compute, 0));
  const int ncclFloat32 = 7, ncclSum = 0, ncclAvg = 4;
  MDB_NCCL(p_AllReduce(ptr, ptr, count, ncclFloat32, average ? ncclAvg : ncclSum, g_comm,
                       g_comm_stream));
  MDB_CUDA(cudaEventRecord(g_ev_comm, g_comm_stream));
  ++g_seq;
  MDB_CUDA(cudaEventRecord(g_ev_seq[g_seq % kSeqRing], g_comm_stream));
  return 0;
}

// Several gradient buffers in ONE NCCL launch (ncclGroupStart/End fuses the calls into one kernel):
// a bias gradient of a few KB rides with its layer's weight gradient instead of paying its own launch.
int mdb_comm_allreduce_multi_f32(void* const* ptrs, const size_t* counts, int n, int average) {
  MDB_REQUIRE(g_comm != nullptr, "communicator not initialised");
  MDB_REQUIRE(n >= 1 && n <= 64 && ptrs && counts, "allreduce_multi: 1..64 buffers");
  MDB_CUDA(cudaEventRecord(g_ev_compute, g_stream));
  MDB_CUDA(cudaStreamWaitEvent(g_comm_stream, g_ev_compute, 0));
  const int ncclFloat32 = 7, ncclSum = 0, ncclAvg = 4;
  MDB_NCCL(p_GroupStart());
  for (int i = 0; i < n; ++i) {
    int r = p_AllReduce(ptrs[i], ptrs[i], counts[i], ncclFloat32, average ? ncclAvg : ncclSum, g_comm, g_comm_stream);
    if (r != 0) {
      p_GroupEnd();
      return set_error(MDB_ECOMM, "ncclAllReduce (grouped) -> %s", p_GetErrorString(r));
    }
  }
  MDB_NCCL(p_GroupEnd());
  MDB_CUDA(cudaEventRecord(g_ev_comm, g_comm_stream));
  ++g_seq;
  MDB_CUDA(cudaEventRecord(g_ev_seq[g_seq % kSeqRing], g_comm_stream));
  return 0;
}

uint64_t mdb_comm_last_seq(void) { return g_seq; }

int mdb_comm_wait_seq(uint64_t seq) {
  if (g_comm == nullptr || seq == 0) return 0;
  // an entry that has left the ring was followed by >= 64 later all-reduces on the same in-order
  // stream: waiting for the newest one covers it
  const uint64_t s = (seq + kSeqRing <= g_seq) ? g_seq : seq;
  MDB_CUDA(cudaStreamWaitEvent(g_stream, g_ev_seq[s % kSeqRing], 0));
  return 0;
}

int mdb_comm_wait(void) {
  if (g_comm == nullptr) return 0;
  MDB_CUDA(cudaStreamWaitEvent(g_stream, g_ev_comm, 0));
  return 0;
}

int mdb_comm_destroy(void) {
  if (g_comm) {
    cudaStreamSynchronize(g_comm_stream);
    p_CommDestroy(g_comm);
    cudaStreamDestroy(g_comm_stream);
    cudaEventDestroy(g_ev_compute);
    cudaEventDestroy(g_ev_comm);
    for (int i = 0; i < kSeqRing; ++i) if (g_ev_seq[i]) { cudaEventDestroy(g_ev_seq[i]); g_ev_seq[i] = nullptr; }
    g_comm = nullptr;
  }
  return 0;
}

}  // extern "C"
