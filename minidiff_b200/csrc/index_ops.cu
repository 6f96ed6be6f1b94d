// Data-dependent helpers of the backend table on the device (round-1 versions staged
// these through host NumPy): index validation + offset arithmetic for integer-array indexing, stream
// compaction (argwhere / boolean-mask indexing: backend/numpy.py:73-75, ops/definitions.py:279-290),
// isin / unravel_index (tensor.py:503-515), key sort for permutation / shuffle / choice without
// replacement, inclusive scan + binary search for weighted choice (tensor.py:598-659).
// Coverage entry points, not on the BASELINE hot path: simple, bounds-checked kernels.
#include "ew_ops.cuh"

namespace mdb {

// One sticky error slot in device memory: kernels that validate data-dependent arguments write the
// first offending value (+ a code) there; the host reads it back right after the launch -- a 16-byte
// D2H copy and a stream sync, the price of NumPy's synchronous IndexError -- unless a CUDA-graph
// capture is open (a capture cannot synchronise: out-of-range indices are then clamped, documented).
struct ErrSlot { int code; int pad; long long value; };
static ErrSlot* g_err_dev = nullptr;
static ErrSlot* g_err_host = nullptr;

static int err_slot_ready() {
  if (g_err_dev) return 0;
  MDB_CUDA(cudaMalloc(&g_err_dev, sizeof(ErrSlot)));
  MDB_CUDA(cudaMemset(g_err_dev, 0, sizeof(ErrSlot)));
  MDB_CUDA(cudaMallocHost(&g_err_host, sizeof(ErrSlot)));
  return 0;
}
static bool stream_capturing() {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(g_stream, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st != cudaStreamCaptureStatusNone;
}
// returns 0 when no kernel reported an error, else the code (and clears the slot)
static int err_slot_fetch(long long* value) {
  if (stream_capturing()) return 0;
  if (cudaMemcpyAsync(g_err_host, g_err_dev, sizeof(ErrSlot), cudaMemcpyDeviceToHost, g_stream) != cudaSuccess ||
      cudaStreamSynchronize(g_stream) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const int code = g_err_host->code;
  if (code) {
    *value = g_err_host->value;
    cudaMemsetAsync(g_err_dev, 0, sizeof(ErrSlot), g_stream);
  }
  return code;
}
__device__ __forceinline__ void report(ErrSlot* e, int code, long long value) {
  if (atomicCAS(&e->code, 0, code) == 0) e->value = value;
}

// ---- integer-array index -> element offsets, validated ---------------------------------------
// off[i] (+)= wrap(idx[i]) * stride       wrap: negative indices count from the end (NumPy)
struct OffParams {
  long long* off;
  const void* idx;
  int idx_dtype, ndim;
  int64_t shape[MDB_MAX_DIMS], istr[MDB_MAX_DIMS];
  int64_t n, extent, stride;
  int accumulate;
  ErrSlot* err;
};
__global__ void __launch_bounds__(256) index_offsets_kernel(const OffParams p) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
    int64_t rem = i, io = 0;
    for (int d = p.ndim - 1; d >= 0; --d) {
      const int64_t q = rem / p.shape[d];
      io += (rem - q * p.shape[d]) * p.istr[d];
      rem = q;
    }
    long long r = load_as<long long>(p.idx, p.idx_dtype, io);
    const long long raw = r;
    if (r < 0) r += p.extent;
    if (r < 0 || r >= p.extent) {
      report(p.err, 1, raw);
      r = r < 0 ? 0 : (p.extent > 0 ? p.extent - 1 : 0);   // keep the access in bounds whatever happens next
    }
    const long long v = r * p.stride;
    p.off[i] = p.accumulate ? p.off[i] + v : v;
  }
}

// ---- stream compaction ------------------------------------------------------------------------
constexpr int kCompactChunk = 2048;     // elements per CTA (256 threads x 8)
__global__ void __launch_bounds__(256) nonzero_count_kernel(const unsigned char* __restrict__ m, int64_t n,
                                                            long long* __restrict__ block_counts) {
  __shared__ int warp_sums[8];
  const int64_t base = (int64_t)blockIdx.x * kCompactChunk;
  int c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t i = base + j * 256 + threadIdx.x;
    c += (i < n && m[i] != 0) ? 1 : 0;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += warp_sums[w];
    block_counts[blockIdx.x] = t;
  }
}
// exclusive scan of the per-CTA counts, in place, by ONE CTA (the counts of 2^31 elements are 2^20
// values: a few hundred microseconds at worst); total -> counts[nblocks]
__global__ void __launch_bounds__(1024) scan_counts_kernel(long long* counts, int64_t nblocks) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nblocks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    long long v = i < nblocks ? counts[i] : 0, incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = warp_tot[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_tot[threadIdx.x] = wi - w;     // exclusive prefix of the warp totals
    }
    __syncthreads();
    const long long carry = carry_s;
    if (i < nblocks) counts[i] = carry + warp_tot[threadIdx.x >> 5] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[31] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[nblocks] = carry_s;
}
__global__ void __launch_bounds__(256) nonzero_write_kernel(const unsigned char* __restrict__ m, int64_t n,
                                                            const long long* __restrict__ block_offsets,
                                                            long long* __restrict__ out) {
  __shared__ int warp_base[8];
  const int64_t base = (int64_t)blockIdx.x * kCompactChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // element order must be preserved: warp w owns elements [base + w*256, +256), lane-major inside
  unsigned ballots[8];
  int total = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t i = base + warp * 256 + j * 32 + lane;
    ballots[j] = __ballot_sync(0xffffffffu, i < n && m[i] != 0);
    total += __popc(ballots[j]);
  }
  if (lane == 0) warp_base[warp] = total;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += warp_base[w];
  long long pos = block_offsets[blockIdx.x] + before;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t i = base + warp * 256 + j * 32 + lane;
    if (ballots[j] >> lane & 1u) out[pos + __popc(ballots[j] & ((1u << lane) - 1u))] = i;
    pos += __popc(ballots[j]);
  }
}

// ---- unravel_index ----------------------------------------------------------------------------
struct UnravelParams {
  const void* idx;
  int idx_dtype, ndim;
  long long* out;          // [ndim][n]
  int64_t n, total;
  int64_t dims[MDB_MAX_DIMS];
  ErrSlot* err;
};
__global__ void __launch_bounds__(256) unravel_kernel(const UnravelParams p) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
    long long v = load_as<long long>(p.idx, p.idx_dtype, i);
    if (v < 0 || v >= p.total) { report(p.err, 2, v); v = 0; }
    for (int d = p.ndim - 1; d >= 0; --d) {
      const long long q = v / p.dims[d];
      p.out[(int64_t)d * p.n + i] = v - q * p.dims[d];
      v = q;
    }
  }
}

// ---- isin: tiled all-pairs comparison -----------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) isin_kernel(const void* elem, int e_dtype, int64_t n, const void* test,
                                                   int t_dtype, int64_t m, unsigned char* out, int invert) {
  __shared__ T tile[1024];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const T v = i < n ? load_as<T>(elem, e_dtype, i) : T(0);
  bool found = false;
  for (int64_t base = 0; base < m; base += 1024) {
    const int cnt = (int)min((int64_t)1024, m - base);
    for (int j = threadIdx.x; j < cnt; j += 256) tile[j] = load_as<T>(test, t_dtype, base + j);
    __syncthreads();
    for (int j = 0; j < cnt; ++j) found |= (tile[j] == v);
    __syncthreads();
  }
  if (i < n) out[i] = (unsigned char)(found != (invert != 0));
}

// ---- bitonic sort of 64-bit keys (ascending) ---------------------------------------------------
// n is a power of two (the host pads with ~0).  Steps whose partner distance fits a 2048-key tile run
// in shared memory, the wider ones in global memory.
constexpr int kSortTile = 2048;
__device__ __forceinline__ void cmpswap(unsigned long long& a, unsigned long long& b, bool up) {
  if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}
// all steps with j < kSortTile of the phases k in [k_first, k_last]
__global__ void __launch_bounds__(1024) bitonic_local_kernel(unsigned long long* keys, int64_t n, int64_t k_first,
                                                             int64_t k_last) {
  __shared__ unsigned long long s[kSortTile];
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  s[threadIdx.x] = keys[base + threadIdx.x];
  s[threadIdx.x + 1024] = keys[base + threadIdx.x + 1024];
  __syncthreads();
  for (int64_t k = k_first; k <= k_last; k <<= 1) {
    for (int64_t j = min(k >> 1, (int64_t)(kSortTile >> 1)); j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int lo = (int)(2 * (t - (t & (int)(j - 1))) + (t & (int)(j - 1)));    // index with bit j clear
      const int64_t g = base + lo;
      cmpswap(s[lo], s[lo + (int)j], (g & k) == 0);
      __syncthreads();
    }
  }
  keys[base + threadIdx.x] = s[threadIdx.x];
  keys[base + threadIdx.x + 1024] = s[threadIdx.x + 1024];
}
__global__ void __launch_bounds__(256) bitonic_global_kernel(unsigned long long* keys, int64_t n, int64_t j, int64_t k) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n / 2) return;
  const int64_t lo = 2 * (t - (t & (j - 1))) + (t & (j - 1));
  unsigned long long a = keys[lo], b = keys[lo + j];
  if ((a > b) == ((lo & k) == 0)) { keys[lo] = b; keys[lo + j] = a; }
}
// keys[i] = (random 32 bits << 32 | i) for i < n, ~0 for the padding
__global__ void __launch_bounds__(256) perm_keys_kernel(unsigned long long* keys, const float* u, int64_t n, int64_t padded) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= padded) return;
  if (i < n) {
    const unsigned int r = __float_as_uint(u[i]);      // raw random bits supplied by mdb_random_bits
    keys[i] = ((unsigned long long)r << 32) | (unsigned long long)(unsigned int)i;
  } else {
    keys[i] = ~0ull;
  }
}
__global__ void __launch_bounds__(256) perm_extract_kernel(const unsigned long long* keys, long long* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (long long)(keys[i] & 0xffffffffull);
}

// ---- inclusive scan (float64) by one CTA + binary search -----------------------------------------
__global__ void __launch_bounds__(1024) cumsum_f64_kernel(const void* in, int in_dtype, double* out, int64_t n) {
  __shared__ double warp_tot[32];
  __shared__ double carry_s;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    double v = i < n ? load_as<double>(in, in_dtype, i) : 0.0, incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      double w = warp_tot[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_tot[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const double carry = carry_s;
    const double r = carry + warp_tot[threadIdx.x >> 5] + incl;
    if (i < n) out[i] = r;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = r;
    __syncthreads();
  }
}
// out[i] = number of cdf entries <= u[i] * cdf[m-1]   (np.searchsorted(cdf / cdf[-1], u, side="right")), clamped to m-1
__global__ void __launch_bounds__(256) searchsorted_kernel(const double* cdf, int64_t m, const double* u, int64_t n,
                                                           long long* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = u[i] * cdf[m - 1];
  int64_t lo = 0, hi = m;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cdf[mid] <= x) lo = mid + 1; else hi = mid;
  }
  out[i] = lo < m ? lo : m - 1;
}

// arange: out[i] = start + i * step (float64 arithmetic for float outputs like NumPy, exact int64 for ints)
__global__ void __launch_bounds__(256) arange_kernel(void* out, int dtype, int64_t n, double start, double step,
                                                     long long istart, long long istep, int integral) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (integral) store_as<long long>(out, dtype, i, istart + (long long)i * istep);
    else store_as<double>(out, dtype, i, start + (double)i * step);
  }
}

static bool is_contig(const mdb_array* a) {
  int64_t st = 1;
  for (int d = a->ndim - 1; d >= 0; --d) {
    if (a->shape[d] != 1 && a->strides[d] != st) return false;
    st *= a->shape[d];
  }
  return true;
}
static bool is_int_dtype(int dt) {
  return dt == MDB_BOOL || dt == MDB_U8 || dt == MDB_I8 || dt == MDB_I16 || dt == MDB_I32 || dt == MDB_I64 ||
         dt == MDB_U16 || dt == MDB_U32 || dt == MDB_U64;
}

}  // namespace mdb

using namespace mdb;

extern "C" {

int mdb_index_offsets(const mdb_array* off, const mdb_array* idx, int64_t extent, int64_t stride, int accumulate) {
  MDB_TRY(ensure_init());
  MDB_TRY(err_slot_ready());
  MDB_REQUIRE(off && idx && off->ptr && idx->ptr, "index_offsets: device arrays required");
  MDB_REQUIRE(off->dtype == MDB_I64 && is_contig(off), "index_offsets: offsets must be contiguous int64");
  MDB_REQUIRE(is_int_dtype(idx->dtype) && idx->dtype != MDB_BOOL, "arrays used as indices must be of integer (or boolean) type");
  MDB_REQUIRE(idx->ndim == off->ndim, "index_offsets: index must be given broadcast to the offsets' shape");
  OffParams p;
  p.off = (long long*)off->ptr; p.idx = idx->ptr; p.idx_dtype = idx->dtype; p.ndim = off->ndim;
  p.n = 1;
  for (int d = 0; d < off->ndim; ++d) {
    MDB_REQUIRE(idx->shape[d] == off->shape[d] || idx->shape[d] == 1, "index_offsets: shape mismatch on axis %d", d);
    p.shape[d] = off->shape[d];
    p.istr[d] = idx->shape[d] == off->shape[d] ? idx->strides[d] : 0;
    p.n *= off->shape[d];
  }
  p.extent = extent; p.stride = stride; p.accumulate = accumulate; p.err = g_err_dev;
  if (p.n == 0) return 0;
  index_offsets_kernel<<<grid_for(p.n, 256), 256, 0, g_stream>>>(p);
  MDB_CHECK_LAUNCH();
  long long bad = 0;
  if (err_slot_fetch(&bad) == 1)
    return set_error(MDB_EINDEX, "index %lld is out of bounds for axis with size %lld", bad, (long long)extent);
  return 0;
}

int mdb_nonzero(const mdb_array* mask, const mdb_array* out_indices, int64_t* count) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(mask && out_indices && count && mask->ptr && out_indices->ptr, "nonzero: device arrays required");
  MDB_REQUIRE((mask->dtype == MDB_BOOL || mask->dtype == MDB_U8) && is_contig(mask), "nonzero: contiguous bool mask required");
  MDB_REQUIRE(out_indices->dtype == MDB_I64 && is_contig(out_indices), "nonzero: contiguous int64 output required");
  const int64_t n = numel(mask);
  MDB_REQUIRE(numel(out_indices) >= n, "nonzero: output must have room for every element");
  MDB_REQUIRE(!stream_capturing(), "nonzero has a data-dependent result size and cannot be captured in a CUDA graph");
  *count = 0;
  if (n == 0) return 0;
  const int64_t nblocks = (n + kCompactChunk - 1) / kCompactChunk;
  TempBuf counts;
  MDB_TRY(counts.alloc((size_t)(nblocks + 1) * sizeof(long long)));
  nonzero_count_kernel<<<(unsigned)nblocks, 256, 0, g_stream>>>((const unsigned char*)mask->ptr, n, (long long*)counts.ptr);
  MDB_CHECK_LAUNCH();
  scan_counts_kernel<<<1, 1024, 0, g_stream>>>((long long*)counts.ptr, nblocks);
  MDB_CHECK_LAUNCH();
  nonzero_write_kernel<<<(unsigned)nblocks, 256, 0, g_stream>>>((const unsigned char*)mask->ptr, n,
                                                               (const long long*)counts.ptr, (long long*)out_indices->ptr);
  MDB_CHECK_LAUNCH();
  long long total = 0;
  MDB_CUDA(cudaMemcpyAsync(&total, (const long long*)counts.ptr + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, g_stream));
  MDB_CUDA(cudaStreamSynchronize(g_stream));
  *count = total;
  return 0;
}

int mdb_unravel_index(const mdb_array* out, const mdb_array* indices, int ndim, const int64_t* dims) {
  MDB_TRY(ensure_init());
  MDB_TRY(err_slot_ready());
  MDB_REQUIRE(out && indices && out->ptr && indices->ptr && dims, "unravel_index: device arrays required");
  MDB_REQUIRE(ndim >= 1 && ndim <= MDB_MAX_DIMS, "unravel_index: 1..%d dimensions", MDB_MAX_DIMS);
  MDB_REQUIRE(is_int_dtype(indices->dtype) && is_contig(indices), "unravel_index: contiguous integer indices required");
  const int64_t n = numel(indices);
  MDB_REQUIRE(out->dtype == MDB_I64 && is_contig(out) && numel(out) == n * ndim, "unravel_index: output must be int64 [ndim, n]");
  UnravelParams p;
  p.idx = indices->ptr; p.idx_dtype = indices->dtype; p.ndim = ndim; p.out = (long long*)out->ptr; p.n = n; p.total = 1;
  for (int d = 0; d < ndim; ++d) { p.dims[d] = dims[d]; p.total *= dims[d]; }
  p.err = g_err_dev;
  if (n == 0) return 0;
  unravel_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>(p);
  MDB_CHECK_LAUNCH();
  long long bad = 0;
  if (err_slot_fetch(&bad) == 2)
    return set_error(MDB_EINVAL, "index %lld is out of bounds for array with size %lld", bad, (long long)p.total);
  return 0;
}

}  // extern "C"

namespace mdb {
// ascending in-place sort of `padded` (a power of two >= kSortTile) 64-bit keys
static int bitonic_sort_u64(unsigned long long* k, int64_t padded) {
  const unsigned tiles = (unsigned)(padded / kSortTile);
  bitonic_local_kernel<<<tiles, 1024, 0, g_stream>>>(k, padded, 2, kSortTile);
  MDB_CHECK_LAUNCH();
  for (int64_t kk = 2 * kSortTile; kk <= padded; kk <<= 1) {
    for (int64_t j = kk >> 1; j >= kSortTile; j >>= 1) {
      bitonic_global_kernel<<<(unsigned)((padded / 2 + 255) / 256), 256, 0, g_stream>>>(k, padded, j, kk);
      MDB_CHECK_LAUNCH();
    }
    bitonic_local_kernel<<<tiles, 1024, 0, g_stream>>>(k, padded, kk, kk);
    MDB_CHECK_LAUNCH();
  }
  return 0;
}

// isin for LARGE test sets: order-preserving 64-bit keys of the test elements, sorted once, then one
// binary search per element (the all-pairs kernel is O(n*m))
template <typename T> __device__ __forceinline__ unsigned long long order_key(T v);
template <> __device__ __forceinline__ unsigned long long order_key<long long>(long long v) {
  return (unsigned long long)v ^ 0x8000000000000000ull;
}
template <> __device__ __forceinline__ unsigned long long order_key<double>(double v) {
  if (v == 0.0) v = 0.0;                                   // -0.0 == +0.0
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
template <typename T>
__global__ void __launch_bounds__(256) isin_keys_kernel(const void* test, int dtype, int64_t m, int64_t padded,
                                                        unsigned long long* keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= padded) return;
  if (i < m) {
    const T v = load_as<T>(test, dtype, i);
    keys[i] = (v != v) ? ~0ull : order_key<T>(v);        // NaN never equals anything: park it with the padding
  } else {
    keys[i] = ~0ull;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) isin_search_kernel(const void* elem, int dtype, int64_t n,
                                                          const unsigned long long* keys, int64_t m,
                                                          unsigned char* out, int invert) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T v = load_as<T>(elem, dtype, i);
  bool found = false;
  if (v == v) {
    const unsigned long long key = order_key<T>(v);
    int64_t lo = 0, hi = m;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    found = lo < m && keys[lo] == key && key != ~0ull;
  }
  out[i] = (unsigned char)(found != (invert != 0));
}
}  // namespace mdb

extern "C" {

int mdb_isin(const mdb_array* out, const mdb_array* elements, const mdb_array* test, int invert) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && elements && test && out->ptr && elements->ptr, "isin: device arrays required");
  MDB_REQUIRE(is_contig(out) && is_contig(elements) && is_contig(test), "isin: contiguous operands required");
  MDB_REQUIRE(out->dtype == MDB_BOOL || out->dtype == MDB_U8, "isin: bool output required");
  const int64_t n = numel(elements), m = numel(test);
  MDB_REQUIRE(numel(out) == n, "isin: output size mismatch");
  if (n == 0) return 0;
  MDB_REQUIRE(m == 0 || test->ptr, "isin: test elements missing");
  const unsigned grid = (unsigned)((n + 255) / 256);
  const bool ints = is_int_dtype(elements->dtype) && is_int_dtype(test->dtype);
  if (m > 4096) {
    // large test set: sort its order-preserving keys once, binary-search every element
    int64_t padded = kSortTile;
    while (padded < m) padded <<= 1;
    TempBuf keys;
    MDB_TRY(keys.alloc((size_t)padded * sizeof(unsigned long long)));
    unsigned long long* k = (unsigned long long*)keys.ptr;
    const unsigned kgrid = (unsigned)((padded + 255) / 256);
    if (ints) isin_keys_kernel<long long><<<kgrid, 256, 0, g_stream>>>(test->ptr, test->dtype, m, padded, k);
    else isin_keys_kernel<double><<<kgrid, 256, 0, g_stream>>>(test->ptr, test->dtype, m, padded, k);
    MDB_CHECK_LAUNCH();
    MDB_TRY(bitonic_sort_u64(k, padded));
    if (ints) isin_search_kernel<long long><<<grid, 256, 0, g_stream>>>(elements->ptr, elements->dtype, n, k, m,
                                                                     (unsigned char*)out->ptr, invert);
    else isin_search_kernel<double><<<grid, 256, 0, g_stream>>>(elements->ptr, elements->dtype, n, k, m,
                                                               (unsigned char*)out->ptr, invert);
    MDB_CHECK_LAUNCH();
    return 0;
  }
  if (ints)
    isin_kernel<long long><<<grid, 256, 0, g_stream>>>(elements->ptr, elements->dtype, n, test->ptr, test->dtype, m,
                                                      (unsigned char*)out->ptr, invert);
  else
    isin_kernel<double><<<grid, 256, 0, g_stream>>>(elements->ptr, elements->dtype, n, test->ptr, test->dtype, m,
                                                   (unsigned char*)out->ptr, invert);
  MDB_CHECK_LAUNCH();
  return 0;
}

/* out[i] = a uniformly random permutation of 0..n-1: sort of (32 random bits, i) keys; `bits` holds n
 * random 32-bit words (mdb_random_bits) */
int mdb_permutation(const mdb_array* out, const mdb_array* bits) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && bits && out->ptr && bits->ptr, "permutation: device arrays required");
  MDB_REQUIRE(out->dtype == MDB_I64 && is_contig(out) && is_contig(bits) && dtype_size(bits->dtype) == 4,
              "permutation: int64 output and 32-bit random words required");
  const int64_t n = numel(out);
  MDB_REQUIRE(numel(bits) >= n && n < (int64_t(1) << 32), "permutation: need n random words, n < 2^32");
  if (n == 0) return 0;
  int64_t padded = kSortTile;
  while (padded < n) padded <<= 1;
  TempBuf keys;
  MDB_TRY(keys.alloc((size_t)padded * sizeof(unsigned long long)));
  unsigned long long* k = (unsigned long long*)keys.ptr;
  perm_keys_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, g_stream>>>(k, (const float*)bits->ptr, n, padded);
  MDB_CHECK_LAUNCH();
  MDB_TRY(bitonic_sort_u64(k, padded));
  perm_extract_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(k, (long long*)out->ptr, n);
  MDB_CHECK_LAUNCH();
  return 0;
}

int mdb_arange(const mdb_array* out, double start, double step, int64_t istart, int64_t istep, int integral) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && out->ptr && is_contig(out), "arange: contiguous output required");
  const int64_t n = numel(out);
  if (n == 0) return 0;
  arange_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>(out->ptr, out->dtype, n, start, step, istart, istep, integral);
  MDB_CHECK_LAUNCH();
  return 0;
}

int mdb_cumsum_f64(const mdb_array* out, const mdb_array* in) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && in && out->ptr && in->ptr && out->dtype == MDB_F64 && is_contig(out) && is_contig(in) &&
              numel(out) == numel(in), "cumsum: contiguous input and float64 output of the same size required");
  const int64_t n = numel(in);
  if (n == 0) return 0;
  cumsum_f64_kernel<<<1, 1024, 0, g_stream>>>(in->ptr, in->dtype, (double*)out->ptr, n);
  MDB_CHECK_LAUNCH();
  return 0;
}

int mdb_searchsorted_cdf(const mdb_array* out, const mdb_array* cdf, const mdb_array* u) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && cdf && u && out->ptr && cdf->ptr && u->ptr, "searchsorted: device arrays required");
  MDB_REQUIRE(out->dtype == MDB_I64 && cdf->dtype == MDB_F64 && u->dtype == MDB_F64 && is_contig(out) && is_contig(cdf) &&
              is_contig(u) && numel(out) == numel(u) && numel(cdf) > 0, "searchsorted: float64 cdf / samples, int64 output");
  const int64_t n = numel(u);
  if (n == 0) return 0;
  searchsorted_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>((const double*)cdf->ptr, numel(cdf),
                                                                        (const double*)u->ptr, n, (long long*)out->ptr);
  MDB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
