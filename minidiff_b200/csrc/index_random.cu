// Integer-array gather / scatter over the leading axis (getitem, __setitem__, index_add with
// array keys: backend/numpy.py:73-75,105) and the counter-based device RNG behind rand / randn
// (backend/numpy.py:131-134).  Coverage entry points, not on the BASELINE hot path.
#include "ew_ops.cuh"

namespace mdb {

struct RowsParams {
  char* a;              // gather: src / scatter: dst   (indexed side)
  char* b;              // gather: out / scatter: src   (dense side, walked by i)
  const long long* idx;
  int64_t n_idx, a_rows, a_row_stride, b_row_stride;   // strides in elements
  int nin;              // inner dims
  int64_t ishape[MDB_MAX_DIMS], a_istr[MDB_MAX_DIMS], b_istr[MDB_MAX_DIMS];
  int64_t inner, total;
  int esize, dtype, mode;  // mode 0 gather, 1 scatter-assign, 2 scatter-add
  int offsets;             // 1: idx holds validated element OFFSETS (mdb_index_offsets), signed, used as is
};

__global__ void __launch_bounds__(256) rows_kernel(const RowsParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += stride) {
    int64_t i = e / p.inner, j = e - i * p.inner;
    long long r = p.idx[i];
    if (!p.offsets) {
      if (r < 0) r += p.a_rows;
      r = r < 0 ? 0 : (r >= p.a_rows ? p.a_rows - 1 : r);
    }
    int64_t ao = r * p.a_row_stride, bo = i * p.b_row_stride;
    for (int d = p.nin - 1; d >= 0; --d) {
      int64_t q = j / p.ishape[d], k = j - q * p.ishape[d];
      j = q;
      ao += k * p.a_istr[d];
      bo += k * p.b_istr[d];
    }
    if (p.mode == 2) {
      switch (p.dtype) {
        case MDB_F32: atomicAdd((float*)p.a + ao, ((const float*)p.b)[bo]); break;
        case MDB_F64: atomicAdd((double*)p.a + ao, ((const double*)p.b)[bo]); break;
        case MDB_I32: atomicAdd((int*)p.a + ao, ((const int*)p.b)[bo]); break;
        case MDB_I64:
          atomicAdd((unsigned long long*)p.a + ao, (unsigned long long)((const long long*)p.b)[bo]);
          break;
        default: break;
      }
      continue;
    }
    char* src = p.mode == 0 ? p.a + ao * p.esize : p.b + bo * p.esize;
    char* dst = p.mode == 0 ? p.b + bo * p.esize : p.a + ao * p.esize;
    switch (p.esize) {
      case 1: *dst = *src; break;
      case 2: *(short*)dst = *(const short*)src; break;
      case 4: *(int*)dst = *(const int*)src; break;
      default: *(long long*)dst = *(const long long*)src; break;
    }
  }
}

static int rows_op(const mdb_array* indexed, const mdb_array* dense, const mdb_array* idx, int mode) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(indexed && dense && idx && indexed->ptr && dense->ptr && idx->ptr, "device arrays required");
  MDB_REQUIRE(idx->dtype == MDB_I64 && idx->ndim == 1 && (idx->shape[0] <= 1 || idx->strides[0] == 1),
              "index vector must be contiguous int64");
  MDB_REQUIRE(indexed->ndim >= 1 && dense->ndim == indexed->ndim, "gather/scatter: rank mismatch");
  MDB_REQUIRE(indexed->dtype == dense->dtype, "gather/scatter: dtype mismatch");
  MDB_REQUIRE(dense->shape[0] == idx->shape[0] || (mode != 0 && dense->shape[0] == 1),
              "gather/scatter: leading extent must equal the number of indices");
  if (mode == 2)
    MDB_REQUIRE(indexed->dtype == MDB_F32 || indexed->dtype == MDB_F64 || indexed->dtype == MDB_I32 ||
                indexed->dtype == MDB_I64, "index_add supports f32/f64/i32/i64");
  RowsParams p;
  p.a = (char*)indexed->ptr; p.b = (char*)dense->ptr; p.idx = (const long long*)idx->ptr;
  p.n_idx = idx->shape[0]; p.a_rows = indexed->shape[0];
  p.offsets = indexed->shape[0] == MDB_ROWS_ARE_OFFSETS ? 1 : 0;
  p.a_row_stride = indexed->strides[0];
  p.b_row_stride = dense->shape[0] == 1 && idx->shape[0] != 1 ? 0 : dense->strides[0];
  p.nin = indexed->ndim - 1; p.inner = 1;
  for (int d = 0; d < p.nin; ++d) {
    int64_t e = indexed->shape[d + 1];
    MDB_REQUIRE(dense->shape[d + 1] == e || (mode != 0 && dense->shape[d + 1] == 1),
                "gather/scatter: inner extents differ on axis %d", d + 1);
    p.ishape[d] = e;
    p.a_istr[d] = indexed->strides[d + 1];
    p.b_istr[d] = dense->shape[d + 1] == e ? dense->strides[d + 1] : 0;
    p.inner *= e;
  }
  p.total = p.n_idx * p.inner;
  p.esize = dtype_size(indexed->dtype); p.dtype = indexed->dtype; p.mode = mode;
  if (p.total == 0) return 0;
  ProfScope prof(PROF_OTHER, (double)p.total * (2.0 * p.esize + 8.0 / (double)(p.inner ? p.inner : 1)));
  MDB_REQUIRE(p.a_rows > 0, "index out of bounds: indexed axis has extent 0");
  rows_kernel<<<grid_for(p.total, 256), 256, 0, g_stream>>>(p);
  MDB_CHECK_LAUNCH();
  return 0;
}

// ---- Philox4x32-10 (Salmon et al. 2011), one 128-bit block per 4 outputs ----------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32(uint64_t ctr, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

// Stream position: `offset`, or -- when ctr != nullptr -- the library's DEVICE-RESIDENT counter, read by
// every thread at kernel start and advanced by a one-thread kernel queued right behind (rng_advance).
// With the position on the device a captured CUDA graph draws fresh numbers on every replay (a host-
// side offset would be baked into the graph: ADVICE r1).
__global__ void rng_advance_kernel(unsigned long long* ctr, unsigned long long inc) { *ctr += inc; }

__global__ void __launch_bounds__(256) random_kernel(void* out, int dtype, int64_t n, int normal,
                                                     uint64_t seed, uint64_t offset, const unsigned long long* ctr) {
  if (ctr) offset = *ctr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // each iteration produces 2 values from one Philox block (64 random bits per value)
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b * 2 < n; b += stride) {
    uint32_t r[4];
    philox4x32((uint64_t)b + offset, seed, r);
    double u0 = ((((uint64_t)r[0] << 32) | r[1]) >> 11) * (1.0 / 9007199254740992.0);
    double u1 = ((((uint64_t)r[2] << 32) | r[3]) >> 11) * (1.0 / 9007199254740992.0);
    double v0 = u0, v1 = u1;
    if (normal) {  // Box-Muller on (0,1] x [0,1)
      double rad = sqrt(-2.0 * log(1.0 - u0));
      double s, c;
      sincospi(2.0 * u1, &s, &c);
      v0 = rad * c; v1 = rad * s;
    }
    int64_t i = b * 2;
    store_as<double>(out, dtype, i, v0);
    if (i + 1 < n) store_as<double>(out, dtype, i + 1, v1);
  }
}

// raw 32-bit words (keys of the permutation sort)
__global__ void __launch_bounds__(256) random_bits_kernel(uint32_t* out, int64_t n, uint64_t seed, uint64_t offset,
                                                          const unsigned long long* ctr) {
  if (ctr) offset = *ctr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b * 4 < n; b += stride) {
    uint32_t r[4];
    philox4x32((uint64_t)b + offset, seed, r);
    for (int j = 0; j < 4; ++j)
      if (b * 4 + j < n) out[b * 4 + j] = r[j];
  }
}

// integers uniform on [low, low + span): 64 random bits scaled by multiply-high (bias < span / 2^64)
__global__ void __launch_bounds__(256) randint_kernel(void* out, int dtype, int64_t n, long long low,
                                                      unsigned long long span, uint64_t seed, uint64_t offset,
                                                      const unsigned long long* ctr) {
  if (ctr) offset = *ctr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b * 2 < n; b += stride) {
    uint32_t r[4];
    philox4x32((uint64_t)b + offset, seed, r);
    for (int j = 0; j < 2; ++j) {
      const int64_t i = b * 2 + j;
      if (i >= n) break;
      const unsigned long long bits = ((unsigned long long)r[2 * j] << 32) | r[2 * j + 1];
      store_as<long long>(out, dtype, i, low + (long long)__umul64hi(bits, span));
    }
  }
}

// binomial(trials, p): sum of `trials` Bernoulli draws per output (p scalar or one value per output)
__global__ void __launch_bounds__(256) binomial_kernel(long long* out, int64_t n, long long trials, const void* p_arr,
                                                       int p_dtype, double p_imm, uint64_t seed, uint64_t offset,
                                                       const unsigned long long* ctr) {
  if (ctr) offset = *ctr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint64_t blocks_per = (uint64_t)((trials + 3) / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double p = p_arr ? load_as<double>(p_arr, p_dtype, i) : p_imm;
    // compare 32 random bits against p scaled to 2^32 (resolution 2.3e-10)
    const double scaled = p * 4294967296.0;
    const unsigned long long thr = scaled <= 0.0 ? 0ull : (scaled >= 4294967296.0 ? 4294967296ull : (unsigned long long)scaled);
    long long c = 0, left = trials;
    for (uint64_t b = 0; b < blocks_per; ++b) {
      uint32_t r[4];
      philox4x32(offset + (uint64_t)i * blocks_per + b, seed, r);
      for (int j = 0; j < 4 && left > 0; ++j, --left) c += ((unsigned long long)r[j] < thr) ? 1 : 0;
    }
    out[i] = c;
  }
}

}  // namespace mdb

using namespace mdb;

extern "C" {

int mdb_gather_rows(const mdb_array* out, const mdb_array* src, const mdb_array* idx) {
  return rows_op(src, out, idx, 0);
}
int mdb_scatter_rows(const mdb_array* dst, const mdb_array* src, const mdb_array* idx, int add) {
  return rows_op(dst, src, idx, add ? 2 : 1);
}

// offset == MDB_RNG_DEVICE_OFFSET selects the device-resident stream position
static unsigned long long* g_rng_ctr = nullptr;
static int rng_counter(uint64_t offset, unsigned long long** ctr) {
  *ctr = nullptr;
  if (offset != MDB_RNG_DEVICE_OFFSET) return 0;
  if (!g_rng_ctr) {
    MDB_CUDA(cudaMalloc(&g_rng_ctr, sizeof(unsigned long long)));
    MDB_CUDA(cudaMemsetAsync(g_rng_ctr, 0, sizeof(unsigned long long), g_stream));
  }
  *ctr = g_rng_ctr;
  return 0;
}
static int rng_advance(unsigned long long* ctr, uint64_t blocks) {
  if (!ctr) return 0;
  rng_advance_kernel<<<1, 1, 0, g_stream>>>(ctr, (unsigned long long)blocks);
  MDB_CHECK_LAUNCH();
  return 0;
}

int mdb_random_reset(uint64_t position) {
  MDB_TRY(ensure_init());
  unsigned long long* ctr = nullptr;
  MDB_TRY(rng_counter(MDB_RNG_DEVICE_OFFSET, &ctr));
  MDB_CUDA(cudaMemcpyAsync(ctr, &position, sizeof(position), cudaMemcpyHostToDevice, g_stream));
  MDB_CUDA(cudaStreamSynchronize(g_stream));      // `position` lives on this stack frame
  return 0;
}

static int contiguous_out(const mdb_array* out, const char* who) {
  int64_t st = 1;
  for (int d = out->ndim - 1; d >= 0; --d) {
    MDB_REQUIRE(out->shape[d] == 1 || out->strides[d] == st, "%s: output must be contiguous", who);
    st *= out->shape[d];
  }
  return 0;
}

int mdb_random_bits(const mdb_array* out, uint64_t seed, uint64_t offset) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && out->ptr && dtype_size(out->dtype) == 4, "random_bits: 32-bit output required");
  MDB_TRY(contiguous_out(out, "random_bits"));
  const int64_t n = numel(out);
  if (n == 0) return 0;
  unsigned long long* ctr = nullptr;
  MDB_TRY(rng_counter(offset, &ctr));
  random_bits_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, g_stream>>>((uint32_t*)out->ptr, n, seed, offset, ctr);
  MDB_CHECK_LAUNCH();
  return rng_advance(ctr, (uint64_t)(n + 3) / 4);
}

int mdb_randint(const mdb_array* out, int64_t low, int64_t high, uint64_t seed, uint64_t offset) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && out->ptr && (out->dtype == MDB_I64 || out->dtype == MDB_I32 || out->dtype == MDB_I16 ||
                                  out->dtype == MDB_I8 || out->dtype == MDB_U8 || out->dtype == MDB_U16 ||
                                  out->dtype == MDB_U32 || out->dtype == MDB_U64),
              "randint: integer output required");
  MDB_REQUIRE(high > low, "low >= high");
  MDB_TRY(contiguous_out(out, "randint"));
  const int64_t n = numel(out);
  if (n == 0) return 0;
  unsigned long long* ctr = nullptr;
  MDB_TRY(rng_counter(offset, &ctr));
  randint_kernel<<<grid_for((n + 1) / 2, 256), 256, 0, g_stream>>>(out->ptr, out->dtype, n, low,
                                                                  (unsigned long long)(high - low), seed, offset, ctr);
  MDB_CHECK_LAUNCH();
  return rng_advance(ctr, (uint64_t)(n + 1) / 2);
}

int mdb_binomial(const mdb_array* out, int64_t trials, const mdb_array* p, uint64_t seed, uint64_t offset) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && out->ptr && out->dtype == MDB_I64 && p, "binomial: int64 output and a probability required");
  MDB_REQUIRE(trials >= 0, "n < 0");
  MDB_TRY(contiguous_out(out, "binomial"));
  const int64_t n = numel(out);
  if (p->ptr) {
    MDB_REQUIRE(numel(p) == n, "binomial: one probability per output (broadcast on the host side)");
    MDB_TRY(contiguous_out(p, "binomial p"));
  } else {
    MDB_REQUIRE(p->imm >= 0.0 && p->imm <= 1.0, "p < 0, p > 1 or p is NaN");
  }
  if (n == 0) return 0;
  unsigned long long* ctr = nullptr;
  MDB_TRY(rng_counter(offset, &ctr));
  binomial_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>((long long*)out->ptr, n, trials, p->ptr, p->dtype, p->imm,
                                                         seed, offset, ctr);
  MDB_CHECK_LAUNCH();
  return rng_advance(ctr, (uint64_t)n * (uint64_t)((trials + 3) / 4));
}

int mdb_random(const mdb_array* out, int normal, uint64_t seed, uint64_t offset) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && out->ptr && (out->dtype == MDB_F32 || out->dtype == MDB_F64),
              "random: contiguous f32/f64 output required");
  int64_t n = numel(out), st = 1;
  for (int d = out->ndim - 1; d >= 0; --d) {
    MDB_REQUIRE(out->shape[d] == 1 || out->strides[d] == st, "random: output must be contiguous");
    st *= out->shape[d];
  }
  if (n == 0) return 0;
  unsigned long long* ctr = nullptr;
  MDB_TRY(rng_counter(offset, &ctr));
  random_kernel<<<grid_for((n + 1) / 2, 256), 256, 0, g_stream>>>(out->ptr, out->dtype, n, normal,
                                                                 seed, offset, ctr);
  MDB_CHECK_LAUNCH();
  return rng_advance(ctr, (uint64_t)(n + 1) / 2);
}

}  // extern "C"
