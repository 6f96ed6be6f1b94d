// Scalar semantics of every elementwise backend function, shared by the fast, generic and fused
// reduce kernels.  Each op mirrors the NumPy function the reference binds in
// minidiff/backend/numpy.py:19-95; the *_BWD forms evaluate the gradient call chains of
// minidiff/ops/definitions.py step by step with explicitly rounded intrinsics (no FMA
// contraction) so a fused launch gives the same bits as the chain of separate launches.
#pragma once
#include <math.h>
#include <type_traits>

#include "mdb_common.cuh"

namespace mdb {

template <typename T> __device__ __forceinline__ T mul_(T a, T b) { return a * b; }
template <> __device__ __forceinline__ float mul_<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_(T a, T b) { return a + b; }
template <> __device__ __forceinline__ float add_<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T sub_(T a, T b) { return a - b; }
template <> __device__ __forceinline__ float sub_<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double sub_<double>(double a, double b) { return __dsub_rn(a, b); }
template <typename T> __device__ __forceinline__ T div_(T a, T b) { return b == 0 ? T(0) : a / b; }
template <> __device__ __forceinline__ float div_<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_<double>(double a, double b) { return __ddiv_rn(a, b); }

// x**e.  NumPy's scalar-exponent fast paths (np.power(x, 2|1|0|0.5|-1) == x*x | x | 1 | sqrt |
// 1/x bit-exactly: SURVEY finding 5) are honoured for every exponent value, immediate or not --
// they are exact results, so applying them to array exponents as well only removes error.
// The general case goes through double exp/log (|rel err| ~1e-14 => correctly rounded fp32 in
// all but ~1e-7 of cases), far inside the 2-ulp budget that powf's documented 4 ulp would break.
__device__ __forceinline__ float pow_f32(float x, float e) {
  if (e == 2.0f) return __fmul_rn(x, x);
  if (e == 1.0f) return x;
  if (e == 0.0f) return 1.0f;
  if (e == 0.5f) return sqrtf(x);
  if (e == -1.0f) return __fdiv_rn(1.0f, x);
  if (x > 0.0f && x < INFINITY && fabsf(e) < INFINITY) return (float)exp((double)e * log((double)x));
  return powf(x, e);
}
// exp for fp32 storage in fp32 arithmetic only (~26 FP32 ops): n = rint(x log2 e), r = x - n ln2
// kept as an exact head r_hi (Cody-Waite, n * LN2_HI is exact) plus a tiny tail r_lo;
// exp(r) = 1 + r + r^2 P(r) with the sum 1 + r_hi formed error-free (Fast2Sum) so that the only
// half-ulp rounding is the final add; 2^n applied in two exact-or-single-rounding steps (handles
// subnormal results and overflow).  Measured against float64 on 8.2 M points: max 0.66 ulp in the
// normal range (0.75 in the subnormal range), 98.5 % correctly rounded, never more than 2 ulp from
// NumPy's own AVX-512 expf (itself up to 2.4 ulp from the truth).  The first version evaluated a
// degree-10 polynomial in double: correctly rounded but FP64-pipe bound at 0.68 of the HBM roofline.
__device__ __forceinline__ float exp_f32(float x) {
  const float n = rintf(__fmul_rn(x, 1.4426950408889634f));
  const float r_hi = __fmaf_rn(n, -0.693145751953125f, x);          // exact: LN2_HI has 16 significant bits
  const float r_lo = __fmul_rn(n, -1.42860682030941723212e-06f);
  const float r = __fadd_rn(r_hi, r_lo);
  float p = 2.48015873015873015873e-05f;                            // 1/8!
  p = __fmaf_rn(p, r, 1.98412698412698412698e-04f);
  p = __fmaf_rn(p, r, 1.38888888888888888889e-03f);
  p = __fmaf_rn(p, r, 8.33333333333333333333e-03f);
  p = __fmaf_rn(p, r, 4.16666666666666666667e-02f);
  p = __fmaf_rn(p, r, 1.66666666666666666667e-01f);
  p = __fmaf_rn(p, r, 0.5f);
  const float q = __fmul_rn(__fmul_rn(r, r), p);
  const float s = __fadd_rn(1.0f, r_hi);
  const float e = __fsub_rn(r_hi, __fsub_rn(s, 1.0f));              // Fast2Sum tail of 1 + r_hi
  const float res = __fadd_rn(s, __fadd_rn(e, __fadd_rn(r_lo, q)));
  const int ni = (int)n;
  // |x| < 87: e^x is a normal number, so 2^n goes straight into the exponent field (3 integer instructions
  // instead of two scale factors and two multiplies; exp was issue-bound at 0.83 of the HBM roofline)
  if (fabsf(x) < 87.0f) return __int_as_float(__float_as_int(res) + (ni << 23));
  if (!(fabsf(x) < 104.0f)) return x != x ? x : (x > 0.0f ? INFINITY : 0.0f);   // nan, +-inf, saturated
  const int n1 = ni >> 1, n2 = ni - n1;                             // |n| <= 151: both scales are normal
  const float s1 = __int_as_float((n1 + 127) << 23), s2 = __int_as_float((n2 + 127) << 23);
  return __fmul_rn(__fmul_rn(res, s1), s2);
}

__device__ __forceinline__ double pow_f64(double x, double e) {
  if (e == 2.0) return x * x;
  if (e == 1.0) return x;
  if (e == 0.0) return 1.0;
  if (e == 0.5) return sqrt(x);
  if (e == -1.0) return 1.0 / x;
  return pow(x, e);
}
__device__ __forceinline__ long long pow_i64(long long x, long long e) {
  if (e < 0) return (x == 1) ? 1 : (x == -1 ? ((e & 1) ? -1 : 1) : 0);
  long long r = 1;
  while (e) {
    if (e & 1) r *= x;
    x *= x;
    e >>= 1;
  }
  return r;
}

// Python-style modulo / floor division == npy_divmod (numpy/_core/src/npymath/npy_math_internal)
template <typename T> __device__ __forceinline__ T fmod_py(T a, T b) {
  if constexpr (std::is_integral_v<T>) {
    if (b == 0) return 0;
    T r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) r += b;
    return r;
  } else {
    T r = fmod(a, b);
    if (b == 0) return r;  // nan
    if (r != 0) {
      if ((b < 0) != (r < 0)) r += b;
    } else {
      r = copysign(T(0), b);
    }
    return r;
  }
}
template <typename T> __device__ __forceinline__ T floordiv_py(T a, T b) {
  if constexpr (std::is_integral_v<T>) {
    if (b == 0) return 0;
    T q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
  } else {
    if (b == 0) return a / b;
    T mod = fmod(a, b);
    T div = (a - mod) / b;
    if (mod != 0 && ((b < 0) != (mod < 0))) div -= T(1);
    T fl;
    if (div != 0) {
      fl = floor(div);
      if (div - fl > T(0.5)) fl += T(1);
    } else {
      fl = copysign(T(0), a / b);
    }
    return fl;
  }
}

__host__ __device__ constexpr bool op_is_predicate(int op) {
  return op == MDB_OP_LOGICAL_NOT || op == MDB_OP_ISNAN || (op >= MDB_OP_EQ && op <= MDB_OP_XOR);
}
__host__ __device__ constexpr bool op_float_only(int op) {
  return (op >= MDB_OP_SIN && op <= MDB_OP_RECIP) || op == MDB_OP_ISNAN || op >= MDB_OP_SIN_BWD;
}
__host__ __device__ constexpr int op_arity(int op) {
  if (op >= MDB_OP_SIN_BWD)
    return (op == MDB_OP_POW_BWD || op == MDB_OP_DIV_BWD_Y || op == MDB_OP_POW_BWD_LIN) ? 3 : 2;
  return op < 32 ? 1 : (op < 64 ? 2 : 3);
}

// a, b, c: operands in backend-call order; d: host-computed auxiliary immediate (POW_BWD: the
// exponent minus one, formed in double on the host exactly like Python forms `y - 1`).
template <int OP, typename T>
__device__ __forceinline__ T apply(T a, T b, T c, T d) {
  constexpr bool I = std::is_integral_v<T>;
  constexpr bool F32 = std::is_same_v<T, float>;
  (void)b; (void)c; (void)d;
  if constexpr (OP == MDB_OP_COPY) return a;
  else if constexpr (OP == MDB_OP_NEG) return -a;
  else if constexpr (OP == MDB_OP_ABS) { if constexpr (I) return a < 0 ? -a : a; else return fabs(a); }
  else if constexpr (OP == MDB_OP_SIGN) {
    if constexpr (!I) { if (a != a) return a; }
    return T((a > 0) - (a < 0));
  }
  else if constexpr (OP == MDB_OP_CEIL) { if constexpr (I) return a; else return ceil(a); }
  else if constexpr (OP == MDB_OP_FLOOR) { if constexpr (I) return a; else return floor(a); }
  else if constexpr (OP == MDB_OP_SQUARE) return mul_(a, a);
  else if constexpr (OP == MDB_OP_LOGICAL_NOT) return T(a == 0);
  else if constexpr (OP == MDB_OP_INVERT) { if constexpr (I) return ~a; else return a; }
  else if constexpr (OP == MDB_OP_ISNAN) return T(a != a);
  else if constexpr (OP == MDB_OP_RELU) return a > 0 ? a : T(0);
  else if constexpr (OP == MDB_OP_ADD) return add_(a, b);
  else if constexpr (OP == MDB_OP_SUB) return sub_(a, b);
  else if constexpr (OP == MDB_OP_MUL) return mul_(a, b);
  else if constexpr (OP == MDB_OP_DIV) return div_(a, b);
  else if constexpr (OP == MDB_OP_MOD) return fmod_py(a, b);
  else if constexpr (OP == MDB_OP_FLOORDIV) return floordiv_py(a, b);
  else if constexpr (OP == MDB_OP_MAXIMUM) { if constexpr (!I) { if (a != a) return a; if (b != b) return b; } return a > b ? a : b; }
  else if constexpr (OP == MDB_OP_MINIMUM) { if constexpr (!I) { if (a != a) return a; if (b != b) return b; } return a < b ? a : b; }
  else if constexpr (OP == MDB_OP_EQ) return T(a == b);
  else if constexpr (OP == MDB_OP_NE) return T(a != b);
  else if constexpr (OP == MDB_OP_GT) return T(a > b);
  else if constexpr (OP == MDB_OP_GE) return T(a >= b);
  else if constexpr (OP == MDB_OP_LT) return T(a < b);
  else if constexpr (OP == MDB_OP_LE) return T(a <= b);
  else if constexpr (OP == MDB_OP_AND) return T((a != 0) && (b != 0));
  else if constexpr (OP == MDB_OP_OR) return T((a != 0) || (b != 0));
  else if constexpr (OP == MDB_OP_XOR) return T((a != 0) != (b != 0));
  else if constexpr (OP == MDB_OP_WHERE) return a != 0 ? b : c;
  else if constexpr (OP == MDB_OP_CLIP) { T v = a; if (v < b) v = b; if (v > c) v = c; return v; }
  else if constexpr (OP == MDB_OP_FMA) return add_(a, mul_(b, c));
  else if constexpr (OP == MDB_OP_POW) {
    if constexpr (I) return pow_i64(a, b);
    else if constexpr (F32) return pow_f32(a, b);
    else return pow_f64(a, b);
  }
  else if constexpr (I) return a;  // float-only ops are never instantiated for integers
  else if constexpr (OP == MDB_OP_SIN) return sin(a);     // sinf for float (1 ulp), sin for double
  else if constexpr (OP == MDB_OP_COS) return cos(a);
  else if constexpr (OP == MDB_OP_EXP) { if constexpr (F32) return exp_f32(a); else return exp(a); }
  else if constexpr (OP == MDB_OP_LOG) return log(a);
  else if constexpr (OP == MDB_OP_SQRT) return sqrt(a);
  else if constexpr (OP == MDB_OP_RECIP) return div_(T(1), a);
  // coverage ops: evaluated in double so the fp32 result is correctly rounded (CUDA's tanf/sinhf
  // are 3-4 ulp, outside the 2-ulp budget)
  else if constexpr (OP == MDB_OP_TAN) return (T)tan((double)a);
  else if constexpr (OP == MDB_OP_SINH) return (T)sinh((double)a);
  else if constexpr (OP == MDB_OP_COSH) return (T)cosh((double)a);
  else if constexpr (OP == MDB_OP_TANH) return (T)tanh((double)a);
  // fused backward chains: a = upstream grad, b = x (, c = y or exponent)
  else if constexpr (OP == MDB_OP_SIN_BWD) return mul_(a, (T)cos(b));
  else if constexpr (OP == MDB_OP_COS_BWD) return mul_(a, mul_(T(-1), (T)sin(b)));
  else if constexpr (OP == MDB_OP_EXP_BWD) { if constexpr (F32) return mul_(a, exp_f32(b)); else return mul_(a, (T)exp(b)); }
  else if constexpr (OP == MDB_OP_LOG_BWD) return div_(a, b);
  else if constexpr (OP == MDB_OP_TANH_BWD) { T ch = (T)cosh((double)b); return mul_(a, div_(T(1), mul_(ch, ch))); }
  else if constexpr (OP == MDB_OP_POW_BWD) {
    // (grad * y) * x**(y-1); c carries the scalar y, d carries y-1
    T pw; if constexpr (F32) pw = pow_f32(b, d); else pw = pow_f64(b, d);
    return mul_(mul_(a, c), pw);
  }
  else if constexpr (OP == MDB_OP_DIV_BWD_Y) return mul_(a, div_(mul_(T(-1), b), mul_(c, c)));
  else if constexpr (OP == MDB_OP_RELU_MASK_BWD) return mul_(a, T(b > 0));
  else if constexpr (OP == MDB_OP_POW_BWD_LIN) return mul_(mul_(a, c), b);
  else return a;
}

// dtype-generic element access for the generic kernels (uniform switch: no divergence)
template <typename T>
__device__ __forceinline__ T load_as(const void* p, int dtype, int64_t off) {
  switch (dtype) {
    case MDB_F32: return (T)((const float*)p)[off];
    case MDB_F64: return (T)((const double*)p)[off];
    case MDB_I64: return (T)((const long long*)p)[off];
    case MDB_I32: return (T)((const int*)p)[off];
    case MDB_BOOL: case MDB_U8: return (T)((const unsigned char*)p)[off];
    case MDB_I8: return (T)((const signed char*)p)[off];
    case MDB_I16: return (T)((const short*)p)[off];
    case MDB_U16: return (T)((const unsigned short*)p)[off];
    case MDB_U32: return (T)((const unsigned int*)p)[off];
    case MDB_U64: return (T)((const unsigned long long*)p)[off];
    default: return T(0);
  }
}
template <typename T>
__device__ __forceinline__ void store_as(void* p, int dtype, int64_t off, T v) {
  switch (dtype) {
    case MDB_F32: ((float*)p)[off] = (float)v; break;
    case MDB_F64: ((double*)p)[off] = (double)v; break;
    case MDB_I64: ((long long*)p)[off] = (long long)v; break;
    case MDB_I32: ((int*)p)[off] = (int)v; break;
    case MDB_BOOL: ((unsigned char*)p)[off] = (unsigned char)(v != T(0)); break;
    case MDB_U8: ((unsigned char*)p)[off] = (unsigned char)(long long)v; break;
    case MDB_I8: ((signed char*)p)[off] = (signed char)(long long)v; break;
    case MDB_I16: ((short*)p)[off] = (short)(long long)v; break;
    case MDB_U16: ((unsigned short*)p)[off] = (unsigned short)(long long)v; break;
    case MDB_U32: ((unsigned int*)p)[off] = (unsigned int)(long long)v; break;
    case MDB_U64: ((unsigned long long*)p)[off] = (unsigned long long)v; break;
    default: break;
  }
}

// 128-bit read-only load that does not allocate in L1: for operands that are streamed exactly once,
// so that the small broadcast operand of the same kernel (a row vector, a per-row scalar) stays
// L1-resident instead of being evicted by the stream (ncu: 49 % L1 hit rate on sum(t*c, axis=1)).
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// ---- operand access of the fast (fp32-compute) kernels: <=3 collapsed dims, inner stride 0/1 ----
enum { K_IMM = 0, K_F32 = 1, K_U8 = 2 };

struct FastOperand {
  const void* ptr;
  int32_t s2, s1;  // outer strides (elements); 32-bit so offsets are single IMAD.WIDE instructions
  int s0;          // inner stride: 0 or 1
  int kind;
  float imm;
};
// Loads are split in two phases so that nothing DEPENDS on a load until every load of the thread is
// in flight: fast_load_raw only issues the load (raw bits, no broadcast copies, no u8 -> float
// conversion), fast_decode turns the raw bits into the VEC operand values at compute time.
// (ncu on the first version: the `v[j] = s` copies of a stride-0 operand and the I2F of a mask sat
// right behind their load, so the 4 work items of a thread were serialised on L2 latency --
// long_scoreboard 14-22 cycles per issue, 0.47-0.79 of the HBM roofline on broadcast forms.)
template <int VEC>
__device__ __forceinline__ void fast_load_raw(const FastOperand& o, uint32_t i2, uint32_t i1,
                                              uint32_t col, uint32_t (&raw)[VEC]) {
  if (o.kind == K_IMM) return;
  const int64_t off = (int64_t)(int32_t)i2 * (int64_t)o.s2 + (int64_t)(int32_t)i1 * (int64_t)o.s1;
  if (o.kind == K_F32) {
    const float* p = (const float*)o.ptr + off;
    if (o.s0 == 0) {
      raw[0] = __float_as_uint(__ldg(p));
    } else if constexpr (VEC == 4) {
      const uint4 q = __ldg((const uint4*)(p + col));
      raw[0] = q.x; raw[1] = q.y; raw[2] = q.z; raw[3] = q.w;
    } else {
      raw[0] = __float_as_uint(__ldg(p + col));
    }
  } else {
    const unsigned char* p = (const unsigned char*)o.ptr + off;
    if (o.s0 == 0) raw[0] = __ldg(p);
    else if constexpr (VEC == 4) raw[0] = __ldg((const unsigned int*)(p + col));
    else raw[0] = __ldg(p + col);
  }
}
template <int VEC>
__device__ __forceinline__ void fast_decode(const FastOperand& o, const uint32_t (&raw)[VEC], float (&v)[VEC]) {
  if (o.kind == K_IMM) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = o.imm;
  } else if (o.kind == K_F32) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = __uint_as_float(o.s0 ? raw[j] : raw[0]);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = (float)(o.s0 ? ((raw[0] >> (8 * j)) & 0xffu) : raw[0]);
  }
}
template <int VEC>
__device__ __forceinline__ void fast_load(const FastOperand& o, uint32_t i2, uint32_t i1,
                                          uint32_t col, float (&v)[VEC]) {
  uint32_t raw[VEC] = {};
  fast_load_raw<VEC>(o, i2, i1, col, raw);
  fast_decode<VEC>(o, raw, v);
}

// ---- host-side shape analysis shared by elementwise / reduce -----------------------------------
struct Collapsed {
  int ndim;                                  // collapsed rank (>= 1)
  int64_t shape[MDB_MAX_DIMS];               // outermost first
  int64_t ostr[MDB_MAX_DIMS];
  int64_t istr[3][MDB_MAX_DIMS];
};

// Broadcast `in[k]` against out's shape, then drop extent-1 axes and merge neighbours that every
// operand walks contiguously.  Returns non-zero (with the NumPy-style message) on shape mismatch.
int collapse(const mdb_array* out, int n_in, const mdb_array* in, Collapsed* c);

enum ComputeClass { CC_F32 = 0, CC_F64 = 1, CC_I64 = 2 };
int compute_class(int op, const mdb_array* out, int n_in, const mdb_array* in);

}  // namespace mdb
