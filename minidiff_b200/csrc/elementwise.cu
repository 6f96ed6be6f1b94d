// Elementwise backend functions (backend/numpy.py:19-95 + the in-place dunders of
// tensor.py:269-362) as two kernels:
//   ew_fast    : fp32 compute, <=3 collapsed dims, unit/zero inner strides -> 128-bit vectorised,
//                4 independent loads in flight per thread, grid sized in multiples of the SM count
//   ew_generic : any dtype / any strides (<=8 dims), one element per thread-iteration
#include <algorithm>
#include <cstdlib>
#include <limits>

#include "ew_ops.cuh"

namespace mdb {

// ------------------------------------------------------------------------------------------------
// host-side shape analysis
// ------------------------------------------------------------------------------------------------
int collapse(const mdb_array* out, int n_in, const mdb_array* in, Collapsed* c) {
  const int nd = out->ndim;
  MDB_REQUIRE(nd >= 0 && nd <= MDB_MAX_DIMS, "ndim %d out of range", nd);
  int64_t shape[MDB_MAX_DIMS], ostr[MDB_MAX_DIMS], istr[3][MDB_MAX_DIMS];
  for (int d = 0; d < nd; ++d) {
    shape[d] = out->shape[d];
    ostr[d] = out->strides[d];
  }
  for (int k = 0; k < n_in; ++k) {
    const mdb_array& a = in[k];
    if (a.ptr == nullptr) {  // immediate
      for (int d = 0; d < nd; ++d) istr[k][d] = 0;
      continue;
    }
    MDB_REQUIRE(a.ndim <= nd, "operands could not be broadcast together: input %d has %d dims, "
                "output has %d", k, a.ndim, nd);
    int lead = nd - a.ndim;
    for (int d = 0; d < nd; ++d) {
      if (d < lead) { istr[k][d] = 0; continue; }
      int64_t ext = a.shape[d - lead];
      if (ext == shape[d]) istr[k][d] = (ext == 1) ? 0 : a.strides[d - lead];
      else if (ext == 1) istr[k][d] = 0;
      else
        return set_error(MDB_EINVAL, "operands could not be broadcast together: input %d extent "
                         "%lld vs output extent %lld on axis %d", k, (long long)ext,
                         (long long)shape[d], d);
    }
  }
  // drop extent-1 axes, then merge (outer, inner) pairs every operand walks as one run
  int m = 0;
  for (int d = 0; d < nd; ++d) {
    if (shape[d] == 1) continue;
    if (m > 0) {
      bool merge = c->ostr[m - 1] == ostr[d] * shape[d];
      for (int k = 0; merge && k < n_in; ++k) merge = c->istr[k][m - 1] == istr[k][d] * shape[d];
      if (merge) {
        c->shape[m - 1] *= shape[d];
        c->ostr[m - 1] = ostr[d];
        for (int k = 0; k < n_in; ++k) c->istr[k][m - 1] = istr[k][d];
        continue;
      }
    }
    c->shape[m] = shape[d];
    c->ostr[m] = ostr[d];
    for (int k = 0; k < n_in; ++k) c->istr[k][m] = istr[k][d];
    ++m;
  }
  if (m == 0) {
    c->shape[0] = 1;
    c->ostr[0] = 1;
    for (int k = 0; k < n_in; ++k) c->istr[k][0] = 0;
    m = 1;
  }
  c->ndim = m;
  return 0;
}

static bool is_small_int(int dt) {
  return dt == MDB_BOOL || dt == MDB_U8 || dt == MDB_I8 || dt == MDB_I16 || dt == MDB_U16;
}

// Which arithmetic the kernel computes in.  The result dtype (NumPy promotion, NEP 50 weak Python
// scalars) is decided by the caller and arrives as out->dtype; predicates compare in the promoted
// type of their inputs.
int compute_class(int op, const mdb_array* out, int n_in, const mdb_array* in) {
  if (!op_is_predicate(op) && out->dtype != MDB_BOOL) {
    if (out->dtype == MDB_F32) return CC_F32;
    if (out->dtype == MDB_F64) return CC_F64;
    if (!op_float_only(op)) return CC_I64;
    return CC_F64;
  }
  bool f64 = false, f32 = false, wide_int = false, imm_float = false, any_array = false;
  for (int k = 0; k < n_in; ++k) {
    if (in[k].ptr == nullptr) {
      if (in[k].imm != (double)in[k].imm_i) imm_float = true;
      if (in[k].dtype == MDB_F64 || in[k].dtype == MDB_F32) imm_float = true;
      continue;
    }
    any_array = true;
    int dt = in[k].dtype;
    if (dt == MDB_F64) f64 = true;
    else if (dt == MDB_F32) f32 = true;
    else if (!is_small_int(dt)) wide_int = true;
  }
  if (f64 || (f32 && wide_int)) return CC_F64;
  if (f32) return CC_F32;
  if (imm_float || !any_array) return CC_F64;
  return CC_I64;
}

// ------------------------------------------------------------------------------------------------
// fast kernel
// ------------------------------------------------------------------------------------------------
struct FastParams {
  void* out;
  int32_t os2, os1;
  FastOperand in[3];
  float aux;
  uint32_t total;  // work items = rows * (inner / VEC)
  FastDiv div_lv, div_d1;
};

// U independent work items per thread: every load is issued before the first use (memory-level
// parallelism), then compute + store.  FLAT: the whole problem is one contiguous run, no index
// decode at all.  CHECK=false is the steady-state body (all U items in range).
template <int OP, int NIN, int VEC, bool FLAT, bool CHECK>
__device__ __forceinline__ void ew_fast_body(const FastParams& p, uint32_t base, uint32_t stride) {
  constexpr int U = 4;
  constexpr bool PRED = op_is_predicate(OP);
  uint32_t raw[U][NIN][VEC];
  int64_t ooff[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t w = base + u * stride;
    if (!CHECK || w < p.total) {
      uint32_t i2 = 0, i1 = 0, col;
      if constexpr (FLAT) {
        col = w * VEC;
        ooff[u] = col;
      } else {
        uint32_t row, cv;
        p.div_lv.divmod(w, row, cv);
        p.div_d1.divmod(row, i2, i1);
        col = cv * VEC;
        ooff[u] = (int64_t)(int32_t)i2 * (int64_t)p.os2 + (int64_t)(int32_t)i1 * (int64_t)p.os1 + col;
      }
#pragma unroll
      for (int k = 0; k < NIN; ++k) fast_load_raw<VEC>(p.in[k], i2, i1, col, raw[u][k]);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t w = base + u * stride;
    if (!CHECK || w < p.total) {
      float v[3][VEC], r[VEC];
#pragma unroll
      for (int k = 0; k < NIN; ++k) fast_decode<VEC>(p.in[k], raw[u][k], v[k]);
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        r[j] = apply<OP, float>(v[0][j], NIN > 1 ? v[1][j] : 0.f, NIN > 2 ? v[2][j] : 0.f, p.aux);
      if constexpr (PRED) {
        unsigned char* o = (unsigned char*)p.out + ooff[u];
        if constexpr (VEC == 4) *(uchar4*)o = make_uchar4(r[0] != 0.f, r[1] != 0.f, r[2] != 0.f, r[3] != 0.f);
        else *o = (unsigned char)(r[0] != 0.f);
      } else {
        float* o = (float*)p.out + ooff[u];
        if constexpr (VEC == 4) *(float4*)o = make_float4(r[0], r[1], r[2], r[3]);
        else *o = r[0];
      }
    }
  }
}

template <int OP, int NIN, int VEC, bool FLAT>
__global__ void __launch_bounds__(256, NIN == 1 ? 6 : (NIN == 2 ? 4 : 3)) ew_fast(const FastParams p) {
  constexpr uint32_t U = 4;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t base = blockIdx.x * blockDim.x + threadIdx.x;
  // steady state: all U items of this thread are in range
  for (; (uint64_t)base + (uint64_t)(U - 1) * stride < p.total; base += stride * U)
    ew_fast_body<OP, NIN, VEC, FLAT, false>(p, base, stride);
  if (base < p.total) ew_fast_body<OP, NIN, VEC, FLAT, true>(p, base, stride);
}

// ------------------------------------------------------------------------------------------------
// row kernel for 2-D broadcast forms (fp32, 128-bit vectors, <= 2 operands)
// ------------------------------------------------------------------------------------------------
// ew_fast decodes (row, column) with two magic-number divisions and rebuilds every operand offset in
// 64-bit arithmetic PER ITEM: ~95 warp instructions per float4 item for the outer product
// (N,1)*(1,M), which made that write-only kernel issue-bound (issue slots 74 % busy) at 0.63 of the
// HBM roofline.  Here a CTA owns 1024 consecutive float4 items of ONE row, so the row decode, the
// operand bases and every per-row constant are computed once per thread, and the access form of each
// operand is a template parameter: FV = unit-stride vector along the row (row pitch may be 0: a
// broadcast row vector), FK = constant along the row (immediate, or a scalar picked by the row index).
enum { FV = 0, FK = 1 };

template <int OP, int NIN, int F0, int F1>
__global__ void __launch_bounds__(256, NIN == 1 ? 6 : 5) ew_rows(const FastParams p, const uint32_t lv) {
  constexpr int U = 4;
  const uint32_t row = blockIdx.x;
  uint32_t i2, i1;
  p.div_d1.divmod(row, i2, i1);
  const int64_t off0 = (int64_t)(int32_t)i2 * p.in[0].s2 + (int64_t)(int32_t)i1 * p.in[0].s1;
  const int64_t off1 = NIN > 1 ? (int64_t)(int32_t)i2 * p.in[1].s2 + (int64_t)(int32_t)i1 * p.in[1].s1 : 0;
  const float4* b0 = F0 == FV ? (const float4*)((const float*)p.in[0].ptr + off0) : nullptr;
  const float4* b1 = (NIN > 1 && F1 == FV) ? (const float4*)((const float*)p.in[1].ptr + off1) : nullptr;
  float k0 = 0.f, k1 = 0.f;
  if constexpr (F0 == FK) k0 = p.in[0].kind == K_IMM ? p.in[0].imm : __ldg((const float*)p.in[0].ptr + off0);
  if constexpr (NIN > 1 && F1 == FK) k1 = p.in[1].kind == K_IMM ? p.in[1].imm : __ldg((const float*)p.in[1].ptr + off1);
  float4* out = (float4*)((float*)p.out + (int64_t)(int32_t)i2 * p.os2 + (int64_t)(int32_t)i1 * p.os1);
  const uint32_t c0 = blockIdx.y * (256 * U) + threadIdx.x;
  float4 x0[U], x1[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (c0 + u * 256 < lv) {
      if constexpr (F0 == FV) x0[u] = __ldg(b0 + c0 + u * 256);
      if constexpr (NIN > 1 && F1 == FV) x1[u] = __ldg(b1 + c0 + u * 256);
    }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (c0 + u * 256 < lv) {
      const float4 a = F0 == FV ? x0[u] : make_float4(k0, k0, k0, k0);
      const float4 b = (NIN > 1 && F1 == FV) ? x1[u] : make_float4(k1, k1, k1, k1);
      float4 r;
      r.x = apply<OP, float>(a.x, b.x, 0.f, p.aux);
      r.y = apply<OP, float>(a.y, b.y, 0.f, p.aux);
      r.z = apply<OP, float>(a.z, b.z, 0.f, p.aux);
      r.w = apply<OP, float>(a.w, b.w, 0.f, p.aux);
      out[c0 + u * 256] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// flat kernel with compile-time operand forms (contiguous problems, 128-bit vectors)
// ------------------------------------------------------------------------------------------------
// ew_fast stages every operand in 4 x VEC registers even when it is an immediate or a broadcast
// scalar, and converts masks behind a runtime switch: where(mask, t, 0) / t > 0 / POW_BWD(k, t, e)
// measured 0.87-0.91 of the HBM roofline at 64-80 registers.  Here each operand is FV (fp32 vector),
// FU (u8 / bool vector, i.e. a mask) or FK (constant: immediate or stride-0 scalar), decided on the
// host, so only streamed operands occupy registers and the loop has no decode.
enum { FU = 2 };

template <int F> struct FlatReg { uint4 v; };
template <int F>
__device__ __forceinline__ void flat_load(const FastOperand& o, uint32_t item, FlatReg<F>& r) {
  if constexpr (F == FV) r.v = __ldg((const uint4*)o.ptr + item);
  else if constexpr (F == FU) r.v.x = __ldg((const unsigned int*)o.ptr + item);
}
template <int F>
__device__ __forceinline__ float flat_get(const FlatReg<F>& r, float k, int j) {
  if constexpr (F == FV) return __uint_as_float(j == 0 ? r.v.x : j == 1 ? r.v.y : j == 2 ? r.v.z : r.v.w);
  else if constexpr (F == FU) return (float)((r.v.x >> (8 * j)) & 0xffu);
  else return k;
}
__device__ __forceinline__ float flat_const(const FastOperand& o) {
  if (o.kind == K_IMM) return o.imm;
  return o.kind == K_F32 ? __ldg((const float*)o.ptr) : (float)__ldg((const unsigned char*)o.ptr);
}

template <int OP, int NIN, int F0, int F1, int F2>
__global__ void __launch_bounds__(256, 6) ew_flat(const FastParams p) {
  constexpr int U = 4;
  constexpr bool PRED = op_is_predicate(OP);
  const float k0 = F0 == FK ? flat_const(p.in[0]) : 0.f;
  const float k1 = (NIN > 1 && F1 == FK) ? flat_const(p.in[1]) : 0.f;
  const float k2 = (NIN > 2 && F2 == FK) ? flat_const(p.in[2]) : 0.f;
  const uint32_t w0 = blockIdx.x * (256 * U) + threadIdx.x;
  FlatReg<F0> a[U];
  FlatReg<F1> b[U];
  FlatReg<F2> c[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (w0 + u * 256 < p.total) {
      flat_load<F0>(p.in[0], w0 + u * 256, a[u]);
      if constexpr (NIN > 1) flat_load<F1>(p.in[1], w0 + u * 256, b[u]);
      if constexpr (NIN > 2) flat_load<F2>(p.in[2], w0 + u * 256, c[u]);
    }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (w0 + u * 256 < p.total) {
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        r[j] = apply<OP, float>(flat_get<F0>(a[u], k0, j), NIN > 1 ? flat_get<F1>(b[u], k1, j) : 0.f,
                                NIN > 2 ? flat_get<F2>(c[u], k2, j) : 0.f, p.aux);
      if constexpr (PRED)
        ((uchar4*)p.out)[w0 + u * 256] = make_uchar4(r[0] != 0.f, r[1] != 0.f, r[2] != 0.f, r[3] != 0.f);
      else
        ((float4*)p.out)[w0 + u * 256] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

// the (op, forms) combinations of the hot paths that ew_fast handles below 0.92 of the roofline
template <int OP, int NIN, int F0, int F1, int F2>
static void launch_flat(const FastParams& p) {
  const uint32_t grid = (p.total + 1023) / 1024;
  ew_flat<OP, NIN, F0, F1, F2><<<grid, 256, 0, g_stream>>>(p);
}
static bool try_flat_forms(int op, int n_in, const FastParams& p, const int (&f)[3]) {
  const int key = f[0] * 100 + (n_in > 1 ? f[1] : 0) * 10 + (n_in > 2 ? f[2] : 0);   // FV=0 FK=1 FU=2
  switch (op) {
    case MDB_OP_WHERE:
      if (key == 201) { launch_flat<MDB_OP_WHERE, 3, FU, FV, FK>(p); return true; }
      if (key == 200) { launch_flat<MDB_OP_WHERE, 3, FU, FV, FV>(p); return true; }
      if (key == 210) { launch_flat<MDB_OP_WHERE, 3, FU, FK, FV>(p); return true; }
      return false;
#define MDB_FLAT_CMP(OPID)                                                         \
    case OPID:                                                                     \
      if (key == 10) { launch_flat<OPID, 2, FV, FK, FK>(p); return true; }         \
      if (key == 0) { launch_flat<OPID, 2, FV, FV, FK>(p); return true; }          \
      return false;
    MDB_FLAT_CMP(MDB_OP_GT) MDB_FLAT_CMP(MDB_OP_GE) MDB_FLAT_CMP(MDB_OP_LT) MDB_FLAT_CMP(MDB_OP_LE)
    MDB_FLAT_CMP(MDB_OP_EQ) MDB_FLAT_CMP(MDB_OP_NE)
#undef MDB_FLAT_CMP
    case MDB_OP_POW_BWD_LIN:
      if (key == 101) { launch_flat<MDB_OP_POW_BWD_LIN, 3, FK, FV, FK>(p); return true; }
      if (key == 1) { launch_flat<MDB_OP_POW_BWD_LIN, 3, FV, FV, FK>(p); return true; }
      return false;
    case MDB_OP_POW_BWD:
      if (key == 101) { launch_flat<MDB_OP_POW_BWD, 3, FK, FV, FK>(p); return true; }
      if (key == 1) { launch_flat<MDB_OP_POW_BWD, 3, FV, FV, FK>(p); return true; }
      return false;
    default:
      return false;
  }
}

// binary arithmetic + fused backward forms that occur with broadcast operands in the hot paths
#define MDB_ROW_BINARY_OPS(X)                                                                   \
  X(MDB_OP_ADD) X(MDB_OP_SUB) X(MDB_OP_MUL) X(MDB_OP_DIV) X(MDB_OP_MAXIMUM) X(MDB_OP_MINIMUM)   \
  X(MDB_OP_SIN_BWD) X(MDB_OP_COS_BWD) X(MDB_OP_EXP_BWD) X(MDB_OP_LOG_BWD)

template <int OP>
static bool launch_rows_binary(const FastParams& p, int f0, int f1, dim3 grid, uint32_t lv) {
  if (f0 == FV && f1 == FV) ew_rows<OP, 2, FV, FV><<<grid, 256, 0, g_stream>>>(p, lv);
  else if (f0 == FV && f1 == FK) ew_rows<OP, 2, FV, FK><<<grid, 256, 0, g_stream>>>(p, lv);
  else if (f0 == FK && f1 == FV) ew_rows<OP, 2, FK, FV><<<grid, 256, 0, g_stream>>>(p, lv);
  else return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------------
struct GenOperand {
  const void* ptr;
  int dtype;
  int64_t str[MDB_MAX_DIMS];
  double imm;
  int64_t imm_i;
};
struct GenParams {
  void* out;
  int out_dtype, ndim;
  int64_t shape[MDB_MAX_DIMS], ostr[MDB_MAX_DIMS];
  int64_t total;
  double aux;
  GenOperand in[3];
};

template <typename T> __device__ __forceinline__ T imm_as(const GenOperand& o) {
  if constexpr (std::is_integral_v<T>) return (T)o.imm_i; else return (T)o.imm;
}

template <int OP, int NIN, typename T>
__global__ void __launch_bounds__(256) ew_generic(const GenParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
    int64_t rem = i, oo = 0, off[3] = {0, 0, 0};
    for (int d = p.ndim - 1; d >= 0; --d) {
      int64_t q = rem / p.shape[d];
      int64_t idx = rem - q * p.shape[d];
      rem = q;
      oo += idx * p.ostr[d];
#pragma unroll
      for (int k = 0; k < NIN; ++k) off[k] += idx * p.in[k].str[d];
    }
    T v[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int k = 0; k < NIN; ++k)
      v[k] = p.in[k].ptr ? load_as<T>(p.in[k].ptr, p.in[k].dtype, off[k]) : imm_as<T>(p.in[k]);
    T r = apply<OP, T>(v[0], v[1], v[2], (T)p.aux);
    store_as<T>(p.out, p.out_dtype, oo, r);
  }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
#define MDB_UNARY_OPS(X)                                                                        \
  X(MDB_OP_COPY) X(MDB_OP_NEG) X(MDB_OP_ABS) X(MDB_OP_SIGN) X(MDB_OP_CEIL) X(MDB_OP_FLOOR)      \
  X(MDB_OP_SIN) X(MDB_OP_COS) X(MDB_OP_TAN) X(MDB_OP_SINH) X(MDB_OP_COSH) X(MDB_OP_TANH)        \
  X(MDB_OP_EXP) X(MDB_OP_LOG) X(MDB_OP_SQRT) X(MDB_OP_RECIP) X(MDB_OP_SQUARE)                   \
  X(MDB_OP_LOGICAL_NOT) X(MDB_OP_INVERT) X(MDB_OP_ISNAN) X(MDB_OP_RELU)
#define MDB_BINARY_OPS(X)                                                                       \
  X(MDB_OP_ADD) X(MDB_OP_SUB) X(MDB_OP_MUL) X(MDB_OP_DIV) X(MDB_OP_POW) X(MDB_OP_MOD)           \
  X(MDB_OP_FLOORDIV) X(MDB_OP_MAXIMUM) X(MDB_OP_MINIMUM) X(MDB_OP_EQ) X(MDB_OP_NE) X(MDB_OP_GT) \
  X(MDB_OP_GE) X(MDB_OP_LT) X(MDB_OP_LE) X(MDB_OP_AND) X(MDB_OP_OR) X(MDB_OP_XOR)               \
  X(MDB_OP_SIN_BWD) X(MDB_OP_COS_BWD) X(MDB_OP_EXP_BWD) X(MDB_OP_LOG_BWD) X(MDB_OP_TANH_BWD)    \
  X(MDB_OP_RELU_MASK_BWD)
#define MDB_TERNARY_OPS(X)                                                                      \
  X(MDB_OP_WHERE) X(MDB_OP_CLIP) X(MDB_OP_FMA) X(MDB_OP_POW_BWD) X(MDB_OP_DIV_BWD_Y)             \
  X(MDB_OP_POW_BWD_LIN)

template <int OP, int NIN>
static int launch_fast(const FastParams& p, int vec, bool flat, int grid) {
  if (vec == 4) {
    if (flat) ew_fast<OP, NIN, 4, true><<<grid, 256, 0, g_stream>>>(p);
    else ew_fast<OP, NIN, 4, false><<<grid, 256, 0, g_stream>>>(p);
  } else {
    ew_fast<OP, NIN, 1, false><<<grid, 256, 0, g_stream>>>(p);   // unaligned / ragged: rare
  }
  MDB_CHECK_LAUNCH();
  return 0;
}

template <int OP, int NIN>
static int launch_generic(const GenParams& p, int cc, int grid) {
  if (cc == CC_F32) ew_generic<OP, NIN, float><<<grid, 256, 0, g_stream>>>(p);
  else if (cc == CC_F64) ew_generic<OP, NIN, double><<<grid, 256, 0, g_stream>>>(p);
  else {
    if constexpr (op_float_only(OP))
      return set_error(MDB_ENOTSUP, "op %d is not defined for integer compute", OP);
    else
      ew_generic<OP, NIN, long long><<<grid, 256, 0, g_stream>>>(p);
  }
  MDB_CHECK_LAUNCH();
  return 0;
}

static bool aligned(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }

int elementwise_impl(int op, const mdb_array* out, int n_in, const mdb_array* in_user) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && out->ptr, "elementwise: output must be a device array");
  mdb_array in[3];
  for (int k = 0; k < n_in; ++k) in[k] = in_user[k];

  // NumPy's exact scalar-exponent forms of power (SURVEY finding 5): rewrite to cheaper ops
  if (op == MDB_OP_POW && in[1].ptr == nullptr && out->dtype != MDB_BOOL &&
      (out->dtype == MDB_F32 || out->dtype == MDB_F64)) {
    double e = in[1].imm;
    if (e == 2.0) { op = MDB_OP_SQUARE; n_in = 1; }
    else if (e == 1.0) { op = MDB_OP_COPY; n_in = 1; }
    else if (e == 0.5) { op = MDB_OP_SQRT; n_in = 1; }
    else if (e == -1.0) { op = MDB_OP_RECIP; n_in = 1; }
    else if (e == 0.0) {
      op = MDB_OP_COPY; n_in = 1;
      in[0].ptr = nullptr; in[0].imm = 1.0; in[0].imm_i = 1; in[0].ndim = 0; in[0].dtype = MDB_F64;
    }
  }
  MDB_REQUIRE(n_in == op_arity(op), "op %d takes %d inputs, got %d", op, op_arity(op), n_in);
  double aux = 0.0;
  if (op == MDB_OP_POW_BWD) {
    MDB_REQUIRE(in[2].ptr == nullptr, "POW_BWD needs an immediate exponent");
    aux = in[2].imm - 1.0;
    if (aux == 1.0) op = MDB_OP_POW_BWD_LIN;   // x**1 == x exactly: (g*2)*x, no pow code in the loop
  }

  Collapsed c;
  MDB_TRY(collapse(out, n_in, in, &c));
  int64_t total = 1;
  for (int d = 0; d < c.ndim; ++d) total *= c.shape[d];
  if (total == 0) return 0;
  const int cc = compute_class(op, out, n_in, in);

  // ---- fast path eligibility
  const bool pred = op_is_predicate(op);
  bool fast = cc == CC_F32 && c.ndim <= 3 && c.ostr[c.ndim - 1] == 1 &&
              ((pred && (out->dtype == MDB_BOOL || out->dtype == MDB_U8)) ||
               (!pred && out->dtype == MDB_F32));
  for (int k = 0; fast && k < n_in; ++k) {
    if (in[k].ptr == nullptr) continue;
    int dt = in[k].dtype;
    int64_t s0 = c.istr[k][c.ndim - 1];
    fast = (dt == MDB_F32 || dt == MDB_BOOL || dt == MDB_U8) && (s0 == 0 || s0 == 1);
  }
  if (fast) {
    const int nd = c.ndim;
    int64_t L = c.shape[nd - 1];
    int64_t d1 = nd >= 2 ? c.shape[nd - 2] : 1, d2 = nd >= 3 ? c.shape[nd - 3] : 1;
    int64_t os1 = nd >= 2 ? c.ostr[nd - 2] : 0, os2 = nd >= 3 ? c.ostr[nd - 3] : 0;
    const size_t osz = pred ? 1 : 4;
    bool v4 = (L % 4 == 0) && aligned(out->ptr, 4 * osz) && os1 % 4 == 0 && os2 % 4 == 0;
    FastParams p;
    auto fits = [](int64_t x) { return x > -(int64_t(1) << 31) && x < (int64_t(1) << 31); };
    bool small_strides = fits(os1) && fits(os2);
    p.out = out->ptr; p.os1 = (int32_t)os1; p.os2 = (int32_t)os2; p.aux = (float)aux;
    for (int k = 0; k < n_in; ++k) {
      FastOperand& o = p.in[k];
      o.ptr = in[k].ptr;
      o.imm = (float)in[k].imm;
      o.kind = in[k].ptr == nullptr ? K_IMM : (in[k].dtype == MDB_F32 ? K_F32 : K_U8);
      o.s0 = (int)c.istr[k][nd - 1];
      const int64_t s1 = nd >= 2 ? c.istr[k][nd - 2] : 0, s2 = nd >= 3 ? c.istr[k][nd - 3] : 0;
      small_strides = small_strides && fits(s1) && fits(s2);
      o.s1 = (int32_t)s1;
      o.s2 = (int32_t)s2;
      if (o.kind != K_IMM && o.s0 == 1) {
        size_t esz = o.kind == K_F32 ? 4 : 1;
        v4 = v4 && aligned(o.ptr, 4 * esz) && o.s1 % 4 == 0 && o.s2 % 4 == 0;
      }
    }
    const int vec = v4 ? 4 : 1;
    int64_t items = d2 * d1 * (L / vec);
    const bool flat = nd == 1;
    if (small_strides && items < (int64_t(1) << 31) && d2 * d1 < (int64_t(1) << 31)) {
      p.total = (uint32_t)items;
      p.div_lv = FastDiv((uint32_t)(L / vec));
      p.div_d1 = FastDiv((uint32_t)d1);
      // contiguous problems whose operand forms are in the specialised table
      static const bool no_flat = getenv("MDB_EW_NO_FLAT") != nullptr;      // A/B switch for measurements
      if (!no_flat && flat && vec == 4) {
        int form[3] = {FK, FK, FK};
        bool ok = true;
        for (int k = 0; k < n_in; ++k) {
          const FastOperand& o = p.in[k];
          if (o.kind == K_IMM || o.s0 == 0) form[k] = FK;
          else form[k] = o.kind == K_F32 ? FV : FU;
          // POW_BWD keeps its exponent in p.aux: operand 2 is only a placeholder
        }
        if (ok && try_flat_forms(op, n_in, p, form)) { MDB_CHECK_LAUNCH(); return 0; }
      }
      // 2-D broadcast forms of fp32 binary ops: one CTA per 1024-float4 chunk of a row
      static const bool no_rows = getenv("MDB_EW_NO_ROWS") != nullptr;      // A/B switch for measurements
      if (!no_rows && !flat && vec == 4 && !pred && n_in == 2 && L / 4 >= 512 && d2 * d1 < (int64_t(1) << 31) &&
          (L / 4 + 1023) / 1024 <= 65535) {
        int form[2];
        bool ok = true;
        for (int k = 0; k < 2; ++k) {
          const FastOperand& o = p.in[k];
          if (o.kind == K_IMM) form[k] = FK;
          else if (o.kind != K_F32) ok = false;
          else form[k] = o.s0 == 1 ? FV : FK;
        }
        if (ok && (form[0] == FV || form[1] == FV)) {
          const uint32_t lv = (uint32_t)(L / 4);
          dim3 rgrid((unsigned)(d2 * d1), (unsigned)((lv + 1023) / 1024));
          bool launched = false;
          switch (op) {
#define X(OPID) case OPID: launched = launch_rows_binary<OPID>(p, form[0], form[1], rgrid, lv); break;
            MDB_ROW_BINARY_OPS(X)
#undef X
            default: break;
          }
          if (launched) { MDB_CHECK_LAUNCH(); return 0; }
        }
      }
      const int per_thread = 4;
      int grid = grid_for((items + per_thread - 1) / per_thread, 256);
      switch (op) {
#define X(OPID) case OPID: return launch_fast<OPID, 1>(p, vec, flat, grid);
        MDB_UNARY_OPS(X)
#undef X
#define X(OPID) case OPID: return launch_fast<OPID, 2>(p, vec, flat, grid);
        MDB_BINARY_OPS(X)
#undef X
#define X(OPID) case OPID: return launch_fast<OPID, 3>(p, vec, flat, grid);
        MDB_TERNARY_OPS(X)
#undef X
        default: return set_error(MDB_EINVAL, "unknown elementwise op %d", op);
      }
    }
  }

  // ---- generic path
  GenParams g;
  g.out = out->ptr; g.out_dtype = out->dtype; g.ndim = c.ndim; g.total = total; g.aux = aux;
  for (int d = 0; d < c.ndim; ++d) { g.shape[d] = c.shape[d]; g.ostr[d] = c.ostr[d]; }
  for (int k = 0; k < n_in; ++k) {
    g.in[k].ptr = in[k].ptr; g.in[k].dtype = in[k].dtype;
    g.in[k].imm = in[k].imm; g.in[k].imm_i = in[k].imm_i;
    for (int d = 0; d < c.ndim; ++d) g.in[k].str[d] = c.istr[k][d];
    MDB_REQUIRE(in[k].ptr == nullptr || in[k].dtype != MDB_F16, "float16 is not supported");
  }
  MDB_REQUIRE(out->dtype != MDB_F16, "float16 is not supported");
  int grid = grid_for(total, 256);
  switch (op) {
#define X(OPID) case OPID: return launch_generic<OPID, 1>(g, cc, grid);
    MDB_UNARY_OPS(X)
#undef X
#define X(OPID) case OPID: return launch_generic<OPID, 2>(g, cc, grid);
    MDB_BINARY_OPS(X)
#undef X
#define X(OPID) case OPID: return launch_generic<OPID, 3>(g, cc, grid);
    MDB_TERNARY_OPS(X)
#undef X
    default: return set_error(MDB_EINVAL, "unknown elementwise op %d", op);
  }
}

}  // namespace mdb

extern "C" {

int mdb_elementwise(int op, const mdb_array* out, int n_in, const mdb_array* in) {
  MDB_REQUIRE(n_in >= 1 && n_in <= 3, "elementwise takes 1..3 inputs, got %d", n_in);
  mdb::ProfScope prof(mdb::PROF_ELEMENTWISE, mdb::algorithmic_bytes(out, n_in, in));
  return mdb::elementwise_impl(op, out, n_in, in);
}

// One call per op for the host shim: the operands come as POINTERS to the descriptors the arrays
// already cache (no descriptor array to marshal), and an output whose ptr is NULL is allocated here
// from the caching allocator (contiguous, numel * itemsize bytes) and its address written back --
// alloc + launch in one ABI crossing instead of two.
int mdb_elementwise_new(int op, mdb_array* out, int n_in, const mdb_array* in0, const mdb_array* in1,
                        const mdb_array* in2) {
  MDB_REQUIRE(n_in >= 1 && n_in <= 3 && out && in0 && (n_in < 2 || in1) && (n_in < 3 || in2),
              "elementwise takes 1..3 inputs, got %d", n_in);
  mdb_array in[3];
  in[0] = *in0;
  if (n_in > 1) in[1] = *in1;
  if (n_in > 2) in[2] = *in2;
  bool fresh = false;
  if (!out->ptr) {
    int64_t n = 1;
    for (int d = 0; d < out->ndim; ++d) n *= out->shape[d];
    void* p = nullptr;
    MDB_TRY(mdb_alloc((size_t)(n > 0 ? n : 1) * (size_t)mdb::dtype_size(out->dtype), &p));
    out->ptr = p;
    fresh = true;
  }
  int rc;
  {
    mdb::ProfScope prof(mdb::PROF_ELEMENTWISE, mdb::algorithmic_bytes(out, n_in, in));
    rc = mdb::elementwise_impl(op, out, n_in, in);
  }
  if (rc != 0 && fresh) { mdb_free(out->ptr); out->ptr = nullptr; }
  return rc;
}

int mdb_copy(const mdb_array* out, const mdb_array* in) {
  mdb::ProfScope prof(mdb::PROF_ELEMENTWISE, mdb::algorithmic_bytes(out, 1, in));
  return mdb::elementwise_impl(MDB_OP_COPY, out, 1, in);
}

int mdb_fill(const mdb_array* out, double value) {
  mdb_array imm;
  imm.ptr = nullptr; imm.dtype = MDB_F64; imm.ndim = 0; imm.imm = value;
  imm.imm_i = (int64_t)value;
  mdb::ProfScope prof(mdb::PROF_ELEMENTWISE, mdb::algorithmic_bytes(out, 0, nullptr));
  return mdb::elementwise_impl(MDB_OP_COPY, out, 1, &imm);
}

}  // extern "C"
