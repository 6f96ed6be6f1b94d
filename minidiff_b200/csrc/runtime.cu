// Runtime of libminidiff_b200: device/stream ownership, caching allocator, copies, events.
// Stands behind array creation / as_numpy / finalizers of the backend boundary.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <set>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <atomic>

#include "mdb_common.cuh"

namespace mdb {
cudaStream_t g_stream = nullptr;
int g_sm_count = 148;
int g_device = -1;
static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};
static std::mutex g_mu;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int ensure_init() {
  if (g_device < 0) return mdb_init(0);
  // the CUDA current device is per THREAD: a call from another Python thread must land on the
  // library's device (LOCAL_RANK != 0 under torchrun), not on device 0
  static thread_local int bound = -1;
  if (bound != g_device) {
    if (cudaSetDevice(g_device) != cudaSuccess) {
      cudaGetLastError();
      return set_error(MDB_ECUDA, "cudaSetDevice(%d) failed on this thread", g_device);
    }
    bound = g_device;
  }
  return 0;
}

// ---- caching allocator ---------------------------------------------------------------------
// Small requests (< 1 MiB): exact 512-B size classes on per-class free lists, one cudaMalloc each.
// Large requests (2 MiB granules): blocks live inside cudaMalloc'ed SEGMENTS and are split and
// coalesced -- best fit over all cached blocks; a block more than 1.5x the request is split and the
// remainder stays cached; freed neighbours merge.  Workloads that change shape (the bench runs a
// 65536-row MLP, then 8192^3 GEMMs, then an 8192-row HVP) therefore reuse the same segments instead
// of paying a device-wide cudaMalloc synchronisation per new size (the first version kept whole
// blocks per size class: 66 cudaMallocs and +3 GB of cache when the HVP followed the MLP).
// Reuse is safe without events because every consumer runs on the single compute stream (stream
// order == program order); the comm stream only touches buffers between mdb_comm_allreduce_* and
// mdb_comm_wait, during which the frontend keeps them alive.
//
// CUDA-graph capture (mdb_graph_*): a captured graph replays the SAME addresses, so every block that
// is handed out while a capture is open is tagged with that graph's private pool: when released --
// during the capture or any time later -- it returns to the pool, never to the general cache, until
// the graph is destroyed.  Temporaries are therefore recycled inside one capture exactly as they
// are in eager mode, and nothing outside the graph can be given memory the graph writes on replay.
// Pool blocks are never split or merged.
struct GraphPool;
struct Block {
  char* ptr;
  size_t size;
  bool is_free;
  bool small;                 // stand-alone cudaMalloc of a small size class (never split / merged)
  Block *prev, *next;         // address-ordered neighbours inside the same segment
  GraphPool* owner;           // non-null: belongs to a captured graph
};
struct GraphPool {
  std::map<size_t, std::vector<Block*>> free_lists;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t launches = 0;     // kernel launches recorded while capturing (added per replay)
  size_t bytes = 0;          // device memory pinned by this graph
};

struct Allocator {
  static constexpr size_t kSmall = size_t(1) << 20, kGranule = size_t(2) << 20;
  std::map<size_t, std::vector<Block*>> small_free;          // size class -> blocks
  std::set<std::pair<size_t, Block*>> large_free;            // (size, block): best fit = lower_bound
  std::unordered_map<void*, Block*> live;
  GraphPool* capturing = nullptr;
  size_t in_use = 0, cached = 0, peak = 0;
  uint64_t n_device_allocs = 0;

  static size_t round(size_t b) {
    if (b == 0) b = 1;
    if (b <= kSmall) return (b + 511) & ~size_t(511);
    return (b + kGranule - 1) & ~(kGranule - 1);
  }
  void account_out(Block* b, void** out) {
    b->is_free = false;
    live[b->ptr] = b;
    in_use += b->size;
    if (in_use > peak) peak = in_use;
    if (capturing) { b->owner = capturing; capturing->bytes += b->size; }
    *out = b->ptr;
  }
  int device_alloc(size_t r, bool small, Block** out) {
    static const bool trace = getenv("MDB_ALLOC_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[mdb alloc miss] %.1f MB, cached %.1f MB\n", r / 1e6, cached / 1e6);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, r);
    if (e != cudaSuccess) {
      cudaGetLastError();
      // flushing the cache synchronises the stream, which would invalidate an open capture
      if (capturing)
        return set_error(MDB_ENOMEM, "out of device memory allocating %zu bytes while capturing a CUDA graph "
                         "(the cache cannot be flushed inside a capture)", r);
      release_cached();
      e = cudaMalloc(&p, r);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(MDB_ENOMEM, "out of device memory allocating %zu bytes (%s)", r, cudaGetErrorString(e));
      }
    }
    ++n_device_allocs;
    *out = new Block{(char*)p, r, false, small, nullptr, nullptr, nullptr};
    return 0;
  }
  int alloc(size_t bytes, void** out) {
    const size_t r = round(bytes);
    const bool small = r <= kSmall;
    if (capturing) {                                          // recycled inside this graph first
      auto it = capturing->free_lists.lower_bound(r);
      while (it != capturing->free_lists.end() && it->second.empty()) ++it;
      if (it != capturing->free_lists.end() && (it->first == r || (!small && it->first <= r + r / 2))) {
        Block* b = it->second.back();
        it->second.pop_back();
        b->is_free = false;
        live[b->ptr] = b;
        in_use += b->size;
        if (in_use > peak) peak = in_use;
        *out = b->ptr;
        return 0;
      }
    }
    Block* b = nullptr;
    if (small) {
      auto it = small_free.find(r);
      if (it != small_free.end() && !it->second.empty()) {
        b = it->second.back();
        it->second.pop_back();
        cached -= b->size;
      } else {
        MDB_TRY(device_alloc(r, true, &b));
      }
    } else {
      auto it = large_free.lower_bound({r, nullptr});
      if (it != large_free.end()) {
        b = it->second;
        large_free.erase(it);
        cached -= b->size;
        if (b->size > r + r / 2 && b->size - r >= kGranule) {   // split: the tail stays cached
          Block* tail = new Block{b->ptr + r, b->size - r, true, false, b, b->next, nullptr};
          if (b->next) b->next->prev = tail;
          b->next = tail;
          b->size = r;
          large_free.insert({tail->size, tail});
          cached += tail->size;
        }
      } else {
        MDB_TRY(device_alloc(r, false, &b));
      }
    }
    account_out(b, out);
    return 0;
  }
  void insert_free_large(Block* b) {       // merge with free, unowned neighbours of the same segment
    while (b->next && b->next->is_free && !b->next->owner) {
      Block* n = b->next;
      large_free.erase({n->size, n});
      cached -= n->size;
      b->size += n->size;
      b->next = n->next;
      if (n->next) n->next->prev = b;
      delete n;
    }
    while (b->prev && b->prev->is_free && !b->prev->owner) {
      Block* q = b->prev;
      large_free.erase({q->size, q});
      cached -= q->size;
      q->size += b->size;
      q->next = b->next;
      if (b->next) b->next->prev = q;
      delete b;
      b = q;
    }
    b->is_free = true;
    large_free.insert({b->size, b});
    cached += b->size;
  }
  int free(void* p) {
    auto it = live.find(p);
    if (it == live.end()) return set_error(MDB_EINVAL, "mdb_free: unknown pointer %p", p);
    Block* b = it->second;
    live.erase(it);
    in_use -= b->size;
    if (b->owner) {                          // graph memory goes back to its graph only
      b->is_free = true;
      b->owner->free_lists[b->size].push_back(b);
      return 0;
    }
    if (b->small) {
      b->is_free = true;
      small_free[b->size].push_back(b);
      cached += b->size;
    } else {
      insert_free_large(b);
    }
    return 0;
  }
  // the graph is gone: its pooled blocks join the general cache, its live blocks become ordinary
  void dissolve(GraphPool* g) {
    for (auto& kv : g->free_lists)
      for (Block* b : kv.second) {
        b->owner = nullptr;
        if (b->small) { small_free[b->size].push_back(b); cached += b->size; }
        else { b->is_free = false; insert_free_large(b); }
      }
    g->free_lists.clear();
    for (auto& kv : live)
      if (kv.second->owner == g) kv.second->owner = nullptr;
  }
  void release_cached() {
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (auto& kv : small_free)
      for (Block* b : kv.second) { cudaFree(b->ptr); cached -= b->size; delete b; }
    small_free.clear();
    // only whole segments can go back to the driver: a free block with no neighbours left
    for (auto it = large_free.begin(); it != large_free.end();) {
      Block* b = it->second;
      if (!b->prev && !b->next) {
        cudaFree(b->ptr);
        cached -= b->size;
        delete b;
        it = large_free.erase(it);
      } else {
        ++it;
      }
    }
  }
};
static Allocator g_alloc;

// ---- profiler ---------------------------------------------------------------------------------
struct ProfRecord { cudaEvent_t a, b; int cls; double work; };
static std::vector<ProfRecord> g_prof;
static size_t g_prof_used = 0;
static bool g_prof_on = false;
static int g_prof_depth = 0;

ProfScope::ProfScope(int cls, double work) : slot(-1) {
  if (!g_prof_on || g_prof_depth++ > 0) return;        // only the outermost entry point records
  if (g_prof_used == g_prof.size()) {
    ProfRecord r;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    g_prof.push_back(r);
  }
  slot = (int)g_prof_used++;
  g_prof[slot].cls = cls;
  g_prof[slot].work = work;
  cudaEventRecord(g_prof[slot].a, g_stream);
}
ProfScope::~ProfScope() {
  if (!g_prof_on) return;
  --g_prof_depth;
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, g_stream);
}

double algorithmic_bytes(const mdb_array* out, int n_in, const mdb_array* in) {
  auto unique_bytes = [](const mdb_array* a) {
    if (!a || !a->ptr) return 0.0;
    double n = 1.0;
    for (int d = 0; d < a->ndim; ++d)
      if (a->strides[d] != 0 || a->shape[d] == 1) n *= (double)a->shape[d];
    return n * dtype_size(a->dtype);
  };
  double b = unique_bytes(out);
  for (int k = 0; k < n_in; ++k) {
    bool alias = false;   // the same view passed twice (x*x) is read once
    for (int j = 0; j < k; ++j)
      alias = alias || (in[j].ptr == in[k].ptr && in[j].ndim == in[k].ndim &&
                        in[j].strides[0] == in[k].strides[0]);
    if (!alias) b += unique_bytes(&in[k]);
  }
  return b;
}
}  // namespace mdb

using namespace mdb;

extern "C" {

int mdb_abi_version(void) { return MDB_ABI_VERSION; }
const char* mdb_last_error(void) { return g_err; }

int mdb_device_count(int* count) {
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    return set_error(MDB_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return 0;
}

int mdb_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device >= 0) {
    if (device != g_device)
      return set_error(MDB_EINVAL, "already initialised on device %d (one backend per process)",
                       g_device);
    return 0;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(MDB_ECUDA, "no CUDA device available (%s); minidiff_b200 has no CPU path",
                     e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  }
  MDB_REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
  MDB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MDB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10 || prop.minor != 0)     // the .so carries sm_100a SASS only (no PTX): sm_103 cannot run it
    return set_error(MDB_ENOTSUP, "device %d is sm_%d%d; this library is built for sm_100a only",
                     device, prop.major, prop.minor);
  g_sm_count = prop.multiProcessorCount;
  MDB_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  g_device = device;
  return 0;
}

int mdb_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device < 0) return 0;
  if (g_alloc.capturing) return set_error(MDB_EINVAL, "cannot shut down while a CUDA-graph capture is open");
  g_alloc.release_cached();
  // forget every remaining block (cached fragments and blocks still referenced by live arrays, whose
  // later mdb_free then reports an unknown pointer instead of corrupting a re-initialised allocator)
  for (auto& kv : g_alloc.live) { if (!kv.second->prev && !kv.second->next) cudaFree(kv.second->ptr); delete kv.second; }
  g_alloc.live.clear();
  for (auto& kv : g_alloc.large_free) delete kv.second;
  g_alloc.large_free.clear();
  g_alloc.in_use = g_alloc.cached = 0;
  cudaStreamDestroy(g_stream);
  g_stream = nullptr;
  g_device = -1;
  return 0;
}

int mdb_device_info(int* sm_count, size_t* total_bytes, int* cc_major, int* cc_minor) {
  MDB_TRY(ensure_init());
  cudaDeviceProp prop;
  MDB_CUDA(cudaGetDeviceProperties(&prop, g_device));
  *sm_count = prop.multiProcessorCount;
  *total_bytes = prop.totalGlobalMem;
  *cc_major = prop.major;
  *cc_minor = prop.minor;
  return 0;
}

void* mdb_stream(void) { return (void*)g_stream; }

int mdb_sync(void) {
  MDB_TRY(ensure_init());
  MDB_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

int mdb_alloc(size_t bytes, void** out) {
  MDB_TRY(ensure_init());
  std::lock_guard<std::mutex> lk(g_mu);
  return g_alloc.alloc(bytes, out);
}
int mdb_free(void* ptr) {
  if (!ptr) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device < 0) return 0;  // after shutdown: memory already returned to the driver
  return g_alloc.free(ptr);
}
int mdb_empty_cache(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device >= 0) g_alloc.release_cached();
  return 0;
}
int mdb_mem_stats(size_t* in_use, size_t* cached, size_t* peak, uint64_t* n_device_allocs) {
  std::lock_guard<std::mutex> lk(g_mu);
  *in_use = g_alloc.in_use;
  *cached = g_alloc.cached;
  *peak = g_alloc.peak;
  *n_device_allocs = g_alloc.n_device_allocs;
  return 0;
}
int mdb_host_alloc(size_t bytes, void** out) {
  MDB_TRY(ensure_init());
  MDB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
int mdb_host_free(void* ptr) {
  if (ptr) MDB_CUDA(cudaFreeHost(ptr));
  return 0;
}

int mdb_h2d(void* dst, const void* src, size_t bytes) {
  MDB_TRY(ensure_init());
  if (bytes) MDB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
  return 0;
}
int mdb_d2h(void* dst, const void* src, size_t bytes) {
  MDB_TRY(ensure_init());
  if (bytes) MDB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
  MDB_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}
// Input prefetch: H2D on a dedicated copy stream so the next batch uploads while the current step
// computes.  The copy first waits for the compute work enqueued so far (the previous user of `dst`),
// mdb_prefetch_wait() then orders the compute stream after the copy.
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_ev_compute_done = nullptr, g_ev_copy_done = nullptr;

int mdb_prefetch_h2d(void* dst, const void* src, size_t bytes) {
  MDB_TRY(ensure_init());
  if (!g_copy_stream) {
    MDB_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    MDB_CUDA(cudaEventCreateWithFlags(&g_ev_compute_done, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&g_ev_copy_done, cudaEventDisableTiming));
  }
  MDB_CUDA(cudaEventRecord(g_ev_compute_done, g_stream));
  MDB_CUDA(cudaStreamWaitEvent(g_copy_stream, g_ev_compute_done, 0));
  if (bytes) MDB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_copy_stream));
  MDB_CUDA(cudaEventRecord(g_ev_copy_done, g_copy_stream));
  return 0;
}

int mdb_prefetch_wait(void) {
  if (g_copy_stream) MDB_CUDA(cudaStreamWaitEvent(g_stream, g_ev_copy_done, 0));
  return 0;
}

int mdb_d2d(void* dst, const void* src, size_t bytes) {
  MDB_TRY(ensure_init());
  if (bytes) MDB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
  return 0;
}

int mdb_event_create(void** ev) {
  MDB_TRY(ensure_init());
  cudaEvent_t e;
  MDB_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return 0;
}
int mdb_event_record(void* ev) {
  MDB_CUDA(cudaEventRecord((cudaEvent_t)ev, g_stream));
  return 0;
}
int mdb_event_elapsed_ms(void* start, void* stop, float* ms) {
  MDB_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  MDB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}
int mdb_event_destroy(void* ev) {
  MDB_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
uint64_t mdb_launch_count(void) { return g_launches.load(); }

// ---- CUDA graphs: capture everything the frontend launches on the compute stream between begin
// and end, replay it with one launch (the device-side counterpart of caching.reuse_graph)
int mdb_graph_begin(void) {
  MDB_TRY(ensure_init());
  std::lock_guard<std::mutex> lk(g_mu);
  MDB_REQUIRE(g_alloc.capturing == nullptr, "a graph capture is already open");
  MDB_REQUIRE(!g_prof_on, "switch the profiler off before capturing (timed events cannot be captured)");
  GraphPool* g = new GraphPool();
  cudaError_t e = cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete g;
    return set_error(MDB_ECUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e));
  }
  g->launches = g_launches.load();
  g_alloc.capturing = g;
  return 0;
}

int mdb_graph_end(void** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  GraphPool* g = g_alloc.capturing;
  MDB_REQUIRE(g != nullptr && out != nullptr, "no graph capture is open");
  g_alloc.capturing = nullptr;
  g->launches = g_launches.load() - g->launches;
  cudaError_t e = cudaStreamEndCapture(g_stream, &g->graph);
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (g->graph) cudaGraphDestroy(g->graph);
    g_alloc.dissolve(g);
    delete g;
    return set_error(MDB_ECUDA, "graph capture failed: %s (a capture must not synchronise, read "
                     "results back or upload from pageable memory)", cudaGetErrorString(e));
  }
  *out = (void*)g;
  return 0;
}

int mdb_graph_launch(void* graph) {
  GraphPool* g = (GraphPool*)graph;
  MDB_REQUIRE(g && g->exec, "invalid graph handle");
  MDB_REQUIRE(g_alloc.capturing == nullptr, "cannot replay a graph while another capture is open");
  MDB_CUDA(cudaGraphLaunch(g->exec, g_stream));
  count_launches((int)g->launches);
  return 0;
}

int mdb_graph_info(void* graph, uint64_t* kernel_launches, size_t* pinned_bytes) {
  GraphPool* g = (GraphPool*)graph;
  MDB_REQUIRE(g && g->exec, "invalid graph handle");
  if (kernel_launches) *kernel_launches = g->launches;
  if (pinned_bytes) *pinned_bytes = g->bytes;
  return 0;
}

int mdb_graph_destroy(void* graph) {
  GraphPool* g = (GraphPool*)graph;
  if (!g) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_device >= 0) {
    cudaStreamSynchronize(g_stream);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    g_alloc.dissolve(g);
  }
  delete g;
  return 0;
}

int mdb_prof_enable(int on) {
  MDB_TRY(ensure_init());
  MDB_CUDA(cudaStreamSynchronize(g_stream));
  g_prof_on = on != 0;
  g_prof_used = 0;
  g_prof_depth = 0;
  return 0;
}

int mdb_prof_read(int cls, double* total_ms, uint64_t* calls, double* work) {
  MDB_REQUIRE(cls >= 0 && cls < PROF_NCLASS, "bad profile class %d", cls);
  MDB_CUDA(cudaStreamSynchronize(g_stream));
  double ms = 0.0, w = 0.0;
  uint64_t n = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    if (g_prof[i].cls != cls) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b) != cudaSuccess) { cudaGetLastError(); continue; }
    ms += t; w += g_prof[i].work; ++n;
  }
  *total_ms = ms; *calls = n; *work = w;
  return 0;
}

}  // extern "C"
