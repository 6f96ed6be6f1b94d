// Shared internals of libminidiff_b200 (not part of the public ABI; see include/minidiff_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "minidiff_b200.h"

namespace mdb {

extern cudaStream_t g_stream;   // the one compute stream
extern int g_sm_count;          // 148 on B200
extern int g_device;

int set_error(int code, const char* fmt, ...);
int ensure_init();
void count_launches(int n = 1);

#define MDB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return mdb::set_error(MDB_ECUDA, "%s -> %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                            __FILE__, __LINE__);                                               \
  } while (0)

#define MDB_CHECK_LAUNCH()                                                                     \
  do {                                                                                         \
    mdb::count_launches();                                                                     \
    cudaError_t e__ = cudaPeekAtLastError();                                                   \
    if (e__ != cudaSuccess) {                                                                  \
      cudaGetLastError();                                                                      \
      return mdb::set_error(MDB_ECUDA, "kernel launch failed: %s (%s:%d)",                     \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                      \
    }                                                                                          \
  } while (0)

#define MDB_REQUIRE(cond, ...)                                                                 \
  do {                                                                                         \
    if (!(cond)) return mdb::set_error(MDB_EINVAL, __VA_ARGS__);                               \
  } while (0)

#define MDB_TRY(expr)                                                                          \
  do {                                                                                         \
    int rc__ = (expr);                                                                         \
    if (rc__ != 0) return rc__;                                                                \
  } while (0)

__host__ __device__ inline int dtype_size(int dt) {
  switch (dt) {
    case MDB_BOOL: case MDB_U8: case MDB_I8: return 1;
    case MDB_I16: case MDB_U16: case MDB_F16: return 2;
    case MDB_I32: case MDB_U32: case MDB_F32: return 4;
    default: return 8;
  }
}
inline bool dtype_is_float(int dt) { return dt == MDB_F32 || dt == MDB_F64 || dt == MDB_F16; }

inline int64_t numel(const mdb_array* a) {
  int64_t n = 1;
  for (int i = 0; i < a->ndim; ++i) n *= a->shape[i];
  return n;
}

// Unsigned 32-bit division by a runtime constant via multiply-high (valid for n < 2^31).
struct FastDiv {
  uint32_t d, m, s;
  FastDiv() : d(1), m(0), s(0) {}
  explicit FastDiv(uint32_t div) : d(div) {
    for (s = 0; s < 32; ++s)
      if ((1u << s) >= d) break;
    uint64_t one = 1;
    m = (uint32_t)(((one << 32) * ((one << s) - d)) / d + 1);
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(n, m) + n) >> s; }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};

// Grid size for the streaming kernels.  Measured on B200 (scripts/microbench_ew.py, 2^26 fp32):
// capping the grid at 8 CTAs/SM with a grid-stride loop reached 0.75-0.94 of the HBM copy peak,
// one pass per CTA (no cap) 0.96-1.01 -- the hardware CTA scheduler balances the tail better than a
// static stride loop, so the default is "as many CTAs as there is work".
inline int grid_for(int64_t work_items, int threads, int max_ctas_per_sm = 1 << 14) {
  int64_t want = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)g_sm_count * max_ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// Per-class kernel timing with CUDA events on the compute stream (bench.py's roofline numbers).
// Off by default; when on, every public compute entry point brackets its launches with an event
// pair and records the ALGORITHMIC work of the call (bytes: each distinct input element once at
// its un-broadcast size + each output element once; flops: 2MNK).
enum ProfClass { PROF_ELEMENTWISE = 0, PROF_REDUCE = 1, PROF_GEMM = 2, PROF_OTHER = 3, PROF_NCLASS = 4 };
struct ProfScope {
  int slot;
  ProfScope(int cls, double work);
  ~ProfScope();
};
double algorithmic_bytes(const mdb_array* out, int n_in, const mdb_array* in);

// optional fused epilogue of the CTA-pair GEMM: C = relu?(A@B + bias?) * (mask_src > 0)?
struct GemmEpilogue {
  const float* bias;       // [N], nullable
  int relu;
  const float* mask_src;   // [M, ld_mask], nullable
  int64_t ld_mask;
};

// temp buffer from the caching allocator, returned on scope exit (stream-ordered => safe)
struct TempBuf {
  void* ptr = nullptr;
  int alloc(size_t bytes) { return mdb_alloc(bytes, &ptr); }
  ~TempBuf() { if (ptr) mdb_free(ptr); }
};

}  // namespace mdb
