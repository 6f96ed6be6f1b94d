// Reductions of the backend boundary: sum/mean/max/min/prod/any/all/argmax/argmin
// (backend/numpy.py:20-57) and the fused "gradient -> un-broadcast -> accumulate" form that
// replaces grad-lambda + md.unbroadcast + `grad + new` (topology.py:93-104,
// ops/definitions.py:157-183) with a single pass.
//
// Fast fp32 kernels (sum/mean/max/min over one run of adjacent axes):
//   red_row_warp / red_row_cta : reduce the innermost (contiguous) run -> warp-shuffle + shared-memory
//                                block reduction, rows split across CTAs when there are few rows
//   red_col                    : reduce a middle/outer run while the inner axis stays contiguous;
//                                32x8 thread tiles, 128-bit column vectors, rows split over grid.y
// Both are two-pass when split (partials -> same kernel again), so results are deterministic.
// Everything else (other dtypes, scattered axes, arg-reductions) takes the generic kernel.
#include <algorithm>
#include <cstdlib>
#include <limits>

#include "ew_ops.cuh"

namespace mdb {

int elementwise_impl(int op, const mdb_array* out, int n_in, const mdb_array* in);

enum { R_SUM = 0, R_MAX = 1, R_MIN = 2 };

template <int RED> __device__ __forceinline__ float red_identity() {
  if constexpr (RED == R_SUM) return 0.f;
  else if constexpr (RED == R_MAX) return -INFINITY;
  else return INFINITY;
}
template <int RED> __device__ __forceinline__ float red_combine(float a, float b) {
  if constexpr (RED == R_SUM) return __fadd_rn(a, b);
  else if constexpr (RED == R_MAX) { if (a != a) return a; if (b != b) return b; return a > b ? a : b; }
  else { if (a != a) return a; if (b != b) return b; return a < b ? a : b; }
}
template <int RED> __device__ __forceinline__ float warp_reduce(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_combine<RED>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct RedParams {
  FastOperand in[3];
  float aux;
  float* dst;              // final output or partial buffer
  int64_t os2, os1;        // output strides over (i2, i1) [row kernels] / (i2) + unit inner [col]
  uint32_t d1;             // extent of collapsed dim 1 (row kernels: rows = d2*d1)
  FastDiv div_d1;
  uint32_t L;              // reduced extent in VEC units (row) / reduced rows (col)
  uint32_t seg;            // work per split, same units as L
  uint32_t nsplit;
  uint32_t I;              // col kernel: inner extent (elements)
  uint32_t rows;           // row kernels: number of rows
  int to_partial;          // 1: dst is the partial buffer [.., nsplit, ..]
  int accumulate;          // final write: dst = dst + result
  float divisor;           // > 0: mean -> result / divisor
  // single-launch split reductions (f32 form kernels): the LAST CTA of a row / column tile to
  // deposit its partial folds all partials (fixed order: deterministic) and writes the result.
  unsigned int* tickets;   // self-resetting arrival counters, or nullptr for the two-launch scheme
  float* final_dst;        // where the last CTA writes (dst is the partial buffer)
  // column sums: `csize` CTAs that share a column tile form a thread-block cluster along the split axis
  // and fold their sums through distributed shared memory; `nparts` = nsplit / csize partial rows remain
  uint32_t csize, nparts;
};

__device__ __forceinline__ uint32_t red_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void red_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float red_ld_dsmem(const float* local, uint32_t rank) {
  float v;
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %1, %2;\n\tld.shared::cluster.f32 %0, [ra];\n\t}"
               : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank) : "memory");
  return v;
}

template <int OP, int NIN, int VEC>
__device__ __forceinline__ void load_apply(const RedParams& p, uint32_t i2, uint32_t i1,
                                           uint32_t col, float (&r)[VEC]) {
  float v[3][VEC];
#pragma unroll
  for (int k = 0; k < NIN; ++k) fast_load<VEC>(p.in[k], i2, i1, col, v[k]);
#pragma unroll
  for (int j = 0; j < VEC; ++j)
    r[j] = apply<OP, float>(v[0][j], NIN > 1 ? v[1][j] : 0.f, NIN > 2 ? v[2][j] : 0.f, p.aux);
}

// U items in two phases: every load of the thread is issued before anything depends on one
// (fast_load_raw keeps raw bits only), then decode -> elementwise functor -> fold into acc.
// LATE: operands 1.. are small (immediates, scalars, row / column vectors that live in L1/L2), so
// only operand 0 is streamed through the raw registers and the others are read at fold time; the
// registers saved buy another resident CTA per SM (ncu: the fused gradient sums sat at 33 %
// occupancy and 0.5-0.6 of the HBM roofline with every operand staged).
// (Tried and dropped: hoisting each operand's (kind, stride) decode into per-thread state -- the
// extra live registers and the 5-way mode select made both fused forms slower, 0.54 / 0.40.)
template <int NIN, int VEC, int U, bool LATE>
__device__ __forceinline__ void load_phase(const RedParams& p, const uint32_t (&i2)[U], const uint32_t (&i1)[U],
                                           const uint32_t (&col)[U], const bool (&ok)[U],
                                           uint32_t (&raw)[U][LATE ? 1 : NIN][VEC]) {
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (ok[u]) {
#pragma unroll
      for (int k = 0; k < (LATE ? 1 : NIN); ++k) fast_load_raw<VEC>(p.in[k], i2[u], i1[u], col[u], raw[u][k]);
    }
}
template <int OP, int NIN, int RED, int VEC, int U, bool LATE>
__device__ __forceinline__ void fold_phase(const RedParams& p, const uint32_t (&i2)[U], const uint32_t (&i1)[U],
                                           const uint32_t (&col)[U], const bool (&ok)[U],
                                           const uint32_t (&raw)[U][LATE ? 1 : NIN][VEC], float (&acc)[VEC]) {
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (ok[u]) {
      float v[3][VEC];
#pragma unroll
      for (int k = 0; k < NIN; ++k) {
        if (LATE && k > 0) fast_load<VEC>(p.in[k], i2[u], i1[u], col[u], v[k]);
        else fast_decode<VEC>(p.in[k], raw[u][LATE ? 0 : k], v[k]);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        acc[j] = red_combine<RED>(acc[j], apply<OP, float>(v[0][j], NIN > 1 ? v[1][j] : 0.f,
                                                           NIN > 2 ? v[2][j] : 0.f, p.aux));
    }
}

__device__ __forceinline__ void final_store(const RedParams& p, float* where, float v) {
  if (p.divisor > 0.f) v = __fdiv_rn(v, p.divisor);
  if (p.accumulate) v = __fadd_rn(*where, v);
  *where = v;
}

// one warp per row (short rows, many rows)
template <int OP, int NIN, int RED, int VEC>
__global__ void __launch_bounds__(256) red_row_warp(const RedParams p) {
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (uint32_t row = blockIdx.x * 8 + wib; row < p.rows; row += gridDim.x * 8) {
    uint32_t i2, i1;
    p.div_d1.divmod(row, i2, i1);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = red_identity<RED>();
    for (uint32_t c = lane; c < p.L; c += 32) {
      float r[VEC];
      load_apply<OP, NIN, VEC>(p, i2, i1, c * VEC, r);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = red_combine<RED>(acc[j], r[j]);
    }
    float a = acc[0];
    if constexpr (VEC == 4)
      a = red_combine<RED>(red_combine<RED>(acc[0], acc[1]), red_combine<RED>(acc[2], acc[3]));
    a = warp_reduce<RED>(a);
    if (lane == 0) final_store(p, p.dst + (int64_t)i2 * p.os2 + (int64_t)i1 * p.os1, a);
  }
}

// one CTA per (row, split): long rows
template <int OP, int NIN, int RED, int VEC, bool LATE = false>
__global__ void __launch_bounds__(256, NIN == 1 ? 6 : (LATE ? 5 : (NIN == 2 ? 4 : 2))) red_row_cta(const RedParams p) {
  constexpr int U = 4;
  __shared__ float sm[8];
  const uint32_t row = blockIdx.x, split = blockIdx.y;
  uint32_t i2, i1;
  p.div_d1.divmod(row, i2, i1);
  const uint32_t start = split * p.seg;
  const uint32_t end = min(start + p.seg, p.L);
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = red_identity<RED>();
  for (uint32_t c = start + threadIdx.x; c < end; c += 256 * U) {
    uint32_t raw[U][LATE ? 1 : NIN][VEC];
    uint32_t a2[U], a1[U], cc[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { a2[u] = i2; a1[u] = i1; cc[u] = (c + u * 256) * VEC; ok[u] = c + u * 256 < end; }
    load_phase<NIN, VEC, U, LATE>(p, a2, a1, cc, ok, raw);
    fold_phase<OP, NIN, RED, VEC, U, LATE>(p, a2, a1, cc, ok, raw, acc);
  }
  float a = acc[0];
  if constexpr (VEC == 4)
    a = red_combine<RED>(red_combine<RED>(acc[0], acc[1]), red_combine<RED>(acc[2], acc[3]));
  a = warp_reduce<RED>(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    float b = threadIdx.x < 8 ? sm[threadIdx.x] : red_identity<RED>();
    b = warp_reduce<RED>(b);
    if (threadIdx.x == 0) {
      if (p.to_partial) p.dst[(int64_t)row * p.nsplit + split] = b;
      else final_store(p, p.dst + (int64_t)i2 * p.os2 + (int64_t)i1 * p.os1, b);
    }
  }
}

// reduce over collapsed dim 1 (rows), keep dim 2 (outer) and dim 0 (inner, contiguous)
template <int OP, int NIN, int RED, int VEC, bool LATE = false>
__global__ void __launch_bounds__(256, NIN == 1 ? 6 : (LATE ? 5 : (NIN == 2 ? 4 : 2))) red_col(const RedParams p) {
  constexpr int U = 4;
  __shared__ float sm[8][32][VEC];
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint32_t col = (blockIdx.x * 32 + tx) * VEC;
  const uint32_t split = blockIdx.y, o2 = blockIdx.z;
  const bool active = col < p.I;
  const uint32_t r0 = split * p.seg, r1 = min(r0 + p.seg, p.L);
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = red_identity<RED>();
  if (active) {
    for (uint32_t r = r0 + ty; r < r1; r += 8 * U) {
      uint32_t raw[U][LATE ? 1 : NIN][VEC];
      uint32_t a2[U], a1[U], cc[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { a2[u] = o2; a1[u] = r + u * 8; cc[u] = col; ok[u] = r + u * 8 < r1; }
      load_phase<NIN, VEC, U, LATE>(p, a2, a1, cc, ok, raw);
      fold_phase<OP, NIN, RED, VEC, U, LATE>(p, a2, a1, cc, ok, raw, acc);
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[ty][tx][j] = acc[j];
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float a = sm[0][tx][j];
#pragma unroll
      for (int t = 1; t < 8; ++t) a = red_combine<RED>(a, sm[t][tx][j]);
      if (p.to_partial) p.dst[((int64_t)o2 * p.nsplit + split) * p.I + col + j] = a;
      else final_store(p, p.dst + (int64_t)o2 * p.os2 + col + j, a);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Form-specialised fp32 kernels (<= 2 operands, 128-bit vectors).  ncu on the generic kernels above:
// sum(t*c, axis=1) and sum(t*a, axis=0) executed ~95 warp instructions per float4 item (runtime
// kind / stride decode, 64-bit offset math per operand per item) and were ISSUE-bound (issue slots
// 72-78 % busy) at 0.62-0.64 of the HBM roofline, while the plain sum needs 36.  Here the access
// form of every operand is a template parameter, so the inner loop is loads + math only:
//   FV  unit-stride fp32 vector (row kernel: along the row; col kernel: 4 adjacent columns, any row pitch)
//   FK  constant for the whole thread (immediate, or a scalar that does not move along the reduced axis)
//   FS  col kernel only: one fp32 scalar per reduced row (e.g. a (N,1) column against (N,M) data)
enum { FV = 0, FK = 1, FS = 2 };

template <int F>
__device__ __forceinline__ float4 form_value(const float4& v, float k) {
  if constexpr (F == FV) return v;
  else return make_float4(k, k, k, k);
}
template <int OP, int NIN, int RED>
__device__ __forceinline__ void fold4(const RedParams& p, const float4& a, const float4& b, float (&acc)[4]) {
  acc[0] = red_combine<RED>(acc[0], apply<OP, float>(a.x, NIN > 1 ? b.x : 0.f, 0.f, p.aux));
  acc[1] = red_combine<RED>(acc[1], apply<OP, float>(a.y, NIN > 1 ? b.y : 0.f, 0.f, p.aux));
  acc[2] = red_combine<RED>(acc[2], apply<OP, float>(a.z, NIN > 1 ? b.z : 0.f, 0.f, p.aux));
  acc[3] = red_combine<RED>(acc[3], apply<OP, float>(a.w, NIN > 1 ? b.w : 0.f, 0.f, p.aux));
}
__device__ __forceinline__ float const_of(const FastOperand& o, int64_t off) {
  return o.kind == K_IMM ? o.imm : __ldg((const float*)o.ptr + off);
}

template <int OP, int NIN, int RED, int F0, int F1>
__global__ void __launch_bounds__(256, (NIN == 2 && F0 == FV && F1 == FV) ? 5 : 6) red_row_f32(const RedParams p) {
  constexpr int U = 4;
  __shared__ float sm[8];
  const uint32_t row = blockIdx.x, split = blockIdx.y;
  uint32_t i2, i1;
  p.div_d1.divmod(row, i2, i1);
  const uint32_t start = split * p.seg;
  const uint32_t end = min(start + p.seg, p.L);
  const int64_t off0 = (int64_t)(int32_t)i2 * p.in[0].s2 + (int64_t)(int32_t)i1 * p.in[0].s1;
  const int64_t off1 = NIN > 1 ? (int64_t)(int32_t)i2 * p.in[1].s2 + (int64_t)(int32_t)i1 * p.in[1].s1 : 0;
  const float4* b0 = F0 == FV ? (const float4*)((const float*)p.in[0].ptr + off0) : nullptr;
  const float4* b1 = (NIN > 1 && F1 == FV) ? (const float4*)((const float*)p.in[1].ptr + off1) : nullptr;
  const float k0 = F0 == FK ? const_of(p.in[0], off0) : 0.f;
  const float k1 = (NIN > 1 && F1 == FK) ? const_of(p.in[1], off1) : 0.f;
  // an operand whose row pitch is zero is re-read by every row: keep it in L1, stream the others past it
  const bool keep0 = p.in[0].s1 == 0 && p.in[0].s2 == 0, keep1 = NIN > 1 && p.in[1].s1 == 0 && p.in[1].s2 == 0;
  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = red_identity<RED>();
  for (uint32_t c = start + threadIdx.x; c < end; c += 256 * U) {
    float4 x0[U], x1[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c + u * 256 < end) {
        if constexpr (F0 == FV) x0[u] = keep0 ? __ldg(b0 + c + u * 256) : ldg_stream4(b0 + c + u * 256);
        if constexpr (NIN > 1 && F1 == FV) x1[u] = keep1 ? __ldg(b1 + c + u * 256) : ldg_stream4(b1 + c + u * 256);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c + u * 256 < end) fold4<OP, NIN, RED>(p, form_value<F0>(x0[u], k0), form_value<F1>(x1[u], k1), acc);
  }
  float a = red_combine<RED>(red_combine<RED>(acc[0], acc[1]), red_combine<RED>(acc[2], acc[3]));
  a = warp_reduce<RED>(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x < 32) {
    float b = threadIdx.x < 8 ? sm[threadIdx.x] : red_identity<RED>();
    b = warp_reduce<RED>(b);
    if (threadIdx.x == 0) {
      s_last = 0;
      if (p.to_partial) {
        p.dst[(int64_t)row * p.nsplit + split] = b;
        if (p.tickets) {
          __threadfence();
          s_last = atomicAdd(&p.tickets[row], 1u) == p.nsplit - 1;
        }
      } else {
        final_store(p, p.dst + (int64_t)i2 * p.os2 + (int64_t)i1 * p.os1, b);
      }
    }
  }
  if (!p.tickets) return;
  __syncthreads();
  if (!s_last) return;
  // last CTA of this row: fold the nsplit partials (strided per thread, then the block tree)
  __threadfence();
  const float* part = p.dst + (int64_t)row * p.nsplit;
  float t = red_identity<RED>();
  for (uint32_t k = threadIdx.x; k < p.nsplit; k += 256) t = red_combine<RED>(t, __ldcg(part + k));
  t = warp_reduce<RED>(t);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    float b = threadIdx.x < 8 ? sm[threadIdx.x] : red_identity<RED>();
    b = warp_reduce<RED>(b);
    if (threadIdx.x == 0) {
      final_store(p, p.final_dst + (int64_t)i2 * p.os2 + (int64_t)i1 * p.os1, b);
      p.tickets[row] = 0;                      // ready for the next launch (and for graph replays)
    }
  }
}

// TX = threads across a column tile (TX * 4 columns, at least one 128-B line per row); 256 / TX row lanes
template <int OP, int NIN, int RED, int F0, int F1, int TX>
__global__ void __launch_bounds__(256, 4) red_col_f32(const RedParams p) {
  constexpr int TY = 256 / TX;
  // 8 rows (float4 each) in flight per thread for single-vector forms: the one-wave cluster launch has at most
  // 4 CTAs per SM and needs ~16 MB outstanding to cover the HBM bandwidth-delay product
  constexpr int U = (NIN == 2 && F0 == FV && F1 == FV) ? 4 : 8;
  __shared__ float sm[TY][TX][4];
  const uint32_t tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const uint32_t col = (blockIdx.x * TX + tx) * 4;
  const uint32_t split = blockIdx.y, o2 = blockIdx.z;
  const bool active = col < p.I;
  const uint32_t r0 = split * p.seg, r1 = min(r0 + p.seg, p.L);
  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = red_identity<RED>();
  if (active) {
    // per-thread bases: vectors start at this thread's 4 columns, scalars at column 0
    const float* q0 = (const float*)p.in[0].ptr + (int64_t)(int32_t)o2 * p.in[0].s2 + (F0 == FV ? col : 0);
    const float* q1 = NIN > 1 ? (const float*)p.in[1].ptr + (int64_t)(int32_t)o2 * p.in[1].s2 + (F1 == FV ? col : 0) : nullptr;
    const int64_t st0 = p.in[0].s1, st1 = NIN > 1 ? p.in[1].s1 : 0;
    const float k0 = F0 == FK ? const_of(p.in[0], (int64_t)(int32_t)o2 * p.in[0].s2) : 0.f;
    const float k1 = (NIN > 1 && F1 == FK) ? const_of(p.in[1], (int64_t)(int32_t)o2 * p.in[1].s2) : 0.f;
    for (uint32_t r = r0 + ty; r < r1; r += TY * U) {
      float4 x0[U], x1[U];
      float s0v[U], s1v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r + u * TY < r1) {
          const int64_t rr = (int64_t)(r + u * TY);
          if constexpr (F0 == FV) x0[u] = st0 ? ldg_stream4((const float4*)(q0 + rr * st0)) : __ldg((const float4*)q0);
          if constexpr (F0 == FS) s0v[u] = __ldg(q0 + rr * st0);
          if constexpr (NIN > 1 && F1 == FV) x1[u] = st1 ? ldg_stream4((const float4*)(q1 + rr * st1)) : __ldg((const float4*)q1);
          if constexpr (NIN > 1 && F1 == FS) s1v[u] = __ldg(q1 + rr * st1);
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r + u * TY < r1)
          fold4<OP, NIN, RED>(p, form_value<F0>(x0[u], F0 == FS ? s0v[u] : k0),
                              form_value<F1>(x1[u], F1 == FS ? s1v[u] : k1), acc);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sm[ty][tx][j] = acc[j];
  __syncthreads();
  if (p.csize > 1) {
    // cluster of csize CTAs (same column tile, consecutive row ranges): fold through distributed shared memory,
    // in rank order, on the rank-0 CTA -- csize times fewer partial rows, no second pass over them when nparts == 1
    __shared__ float cs[TX * 4];
    if (ty == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = sm[0][tx][j];
#pragma unroll
        for (int t = 1; t < TY; ++t) a = red_combine<RED>(a, sm[t][tx][j]);
        cs[tx * 4 + j] = a;                                   // columns past I hold the identity
      }
    }
    red_cluster_sync();
    const uint32_t crank = red_cluster_ctarank();
    if (crank == 0 && threadIdx.x < TX * 4) {
      float a = cs[threadIdx.x];
      for (uint32_t r = 1; r < p.csize; ++r) a = red_combine<RED>(a, red_ld_dsmem(&cs[threadIdx.x], r));
      const uint32_t c = blockIdx.x * (TX * 4) + threadIdx.x;
      if (c < p.I) {
        if (p.to_partial) p.dst[((int64_t)o2 * p.nparts + split / p.csize) * p.I + c] = a;
        else final_store(p, p.dst + (int64_t)o2 * p.os2 + c, a);
      }
    }
    red_cluster_sync();                                       // peers stay resident until rank 0 has read them
    if (crank != 0) return;
  } else if (ty == 0 && active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = sm[0][tx][j];
#pragma unroll
      for (int t = 1; t < TY; ++t) a = red_combine<RED>(a, sm[t][tx][j]);
      if (p.to_partial) p.dst[((int64_t)o2 * p.nparts + split) * p.I + col + j] = a;
      else final_store(p, p.dst + (int64_t)o2 * p.os2 + col + j, a);
    }
  }
  if (!p.tickets || !p.to_partial) return;
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  const unsigned int tile = o2 * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) s_last = atomicAdd(&p.tickets[tile], 1u) == p.nparts - 1;
  __syncthreads();
  if (!s_last) return;
  // last CTA of this column tile: fold the partial rows of all splits, 8 row groups then the tree
  __threadfence();
  float t4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) t4[j] = red_identity<RED>();
  if (active) {
    for (uint32_t k = ty; k < p.nparts; k += TY) {
      const float4 v = __ldcg((const float4*)(p.dst + ((int64_t)o2 * p.nparts + k) * p.I + col));
      t4[0] = red_combine<RED>(t4[0], v.x); t4[1] = red_combine<RED>(t4[1], v.y);
      t4[2] = red_combine<RED>(t4[2], v.z); t4[3] = red_combine<RED>(t4[3], v.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sm[ty][tx][j] = t4[j];
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = sm[0][tx][j];
#pragma unroll
      for (int t = 1; t < TY; ++t) a = red_combine<RED>(a, sm[t][tx][j]);
      final_store(p, p.final_dst + (int64_t)o2 * p.os2 + col + j, a);
    }
  }
  if (threadIdx.x == 0) p.tickets[tile] = 0;
}

// ------------------------------------------------------------------------------------------------
// generic fallback: one thread per output element, any dtype / strides / axes
// ------------------------------------------------------------------------------------------------
struct GenRedParams {
  const void* in;
  void* out;
  int in_dtype, out_dtype, red;
  int nk, nr;                              // kept / reduced axis counts
  int64_t kshape[MDB_MAX_DIMS], kistr[MDB_MAX_DIMS], kostr[MDB_MAX_DIMS];
  int64_t rshape[MDB_MAX_DIMS], ristr[MDB_MAX_DIMS];
  int64_t n_out, n_red;
};

template <typename T>
__global__ void __launch_bounds__(128) red_generic(const GenRedParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < p.n_out; o += stride) {
    int64_t rem = o, ibase = 0, obase = 0;
    for (int d = p.nk - 1; d >= 0; --d) {
      int64_t q = rem / p.kshape[d], idx = rem - q * p.kshape[d];
      rem = q;
      ibase += idx * p.kistr[d];
      obase += idx * p.kostr[d];
    }
    T acc = T(0);
    long long best = 0;
    bool first = true;
    if (p.red == MDB_RED_PROD || p.red == MDB_RED_ALL) acc = T(1);
    for (int64_t r = 0; r < p.n_red; ++r) {
      int64_t rr = r, off = ibase;
      for (int d = p.nr - 1; d >= 0; --d) {
        int64_t q = rr / p.rshape[d], idx = rr - q * p.rshape[d];
        rr = q;
        off += idx * p.ristr[d];
      }
      T v = load_as<T>(p.in, p.in_dtype, off);
      switch (p.red) {
        case MDB_RED_SUM: case MDB_RED_MEAN: acc += v; break;
        case MDB_RED_PROD: acc *= v; break;
        case MDB_RED_ANY: acc = T((acc != T(0)) || (v != T(0))); break;
        case MDB_RED_ALL: acc = T((acc != T(0)) && (v != T(0))); break;
        case MDB_RED_MAX:
          if (first || (acc == acc && (v > acc || v != v))) acc = v;
          break;
        case MDB_RED_MIN:
          if (first || (acc == acc && (v < acc || v != v))) acc = v;
          break;
        case MDB_RED_ARGMAX:
          if (first || (acc == acc && (v > acc || v != v))) { acc = v; best = r; }
          break;
        case MDB_RED_ARGMIN:
          if (first || (acc == acc && (v < acc || v != v))) { acc = v; best = r; }
          break;
      }
      first = false;
    }
    if (p.red == MDB_RED_MEAN) acc = acc / (T)p.n_red;
    if (p.red == MDB_RED_ARGMAX || p.red == MDB_RED_ARGMIN)
      store_as<long long>(p.out, p.out_dtype, obase, best);
    else
      store_as<T>(p.out, p.out_dtype, obase, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// host logic
// ------------------------------------------------------------------------------------------------
struct RedPlan {          // iteration space collapsed to [d2, d1, d0]
  int pattern;            // 0: reduce d0 (row); 1: reduce d1 keeping d0 (col); -1: none
  int64_t d2, d1, d0;
  int64_t istr[3][3];     // per input: strides over (d2, d1, d0)
  int64_t ostr[3];        // output strides over (d2, d1, d0) (0 on reduced dims)
};

// full: iteration shape; red[d]: axis reduced; istr[k][d] / ostr[d]: strides (0 where broadcast)
static bool plan_fast(int nd, const int64_t* full, const bool* red, int n_in,
                      const int64_t (*istr)[MDB_MAX_DIMS], const int64_t* ostr, RedPlan* pl) {
  int64_t shape[MDB_MAX_DIMS], is[3][MDB_MAX_DIMS], os[MDB_MAX_DIMS];
  bool rf[MDB_MAX_DIMS];
  int m = 0;
  for (int d = 0; d < nd; ++d) {
    if (full[d] == 1) continue;
    if (m > 0 && rf[m - 1] == red[d]) {
      bool merge = red[d] || os[m - 1] == ostr[d] * full[d];
      for (int k = 0; merge && k < n_in; ++k) merge = is[k][m - 1] == istr[k][d] * full[d];
      if (merge) {
        shape[m - 1] *= full[d];
        os[m - 1] = ostr[d];
        for (int k = 0; k < n_in; ++k) is[k][m - 1] = istr[k][d];
        continue;
      }
    }
    shape[m] = full[d]; rf[m] = red[d]; os[m] = red[d] ? 0 : ostr[d];
    for (int k = 0; k < n_in; ++k) is[k][m] = istr[k][d];
    ++m;
  }
  if (m == 0 || m > 3) return false;
  // recognise [R], [K,R], [K,K,R] (row) and [R,K], [K,R,K] (col)
  int pattern = -1;
  if (rf[m - 1]) {
    pattern = 0;
    for (int d = 0; d < m - 1; ++d) if (rf[d]) return false;
  } else if (m >= 2 && rf[m - 2]) {
    pattern = 1;
    if (m == 3 && rf[0]) return false;
  } else {
    return false;
  }
  pl->pattern = pattern;
  int64_t sh3[3] = {1, 1, 1};
  for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) pl->istr[k][j] = 0;
  pl->ostr[0] = pl->ostr[1] = pl->ostr[2] = 0;
  for (int d = 0; d < m; ++d) {
    int slot = 3 - m + d;
    sh3[slot] = shape[d];
    pl->ostr[slot] = os[d];
    for (int k = 0; k < n_in; ++k) pl->istr[k][slot] = is[k][d];
  }
  pl->d2 = sh3[0]; pl->d1 = sh3[1]; pl->d0 = sh3[2];
  for (int k = 0; k < n_in; ++k) {
    int64_t s0 = pl->istr[k][2];
    if (s0 != 0 && s0 != 1) return false;
  }
  if (pattern == 1 && pl->ostr[2] != 1) return false;
  return true;
}

static bool aligned16(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }

#define MDB_FUSED_RED_OPS(X)                                                                    \
  X(MDB_OP_COPY, 1) X(MDB_OP_NEG, 1) X(MDB_OP_MUL, 2) X(MDB_OP_DIV, 2) X(MDB_OP_SIN_BWD, 2)     \
  X(MDB_OP_COS_BWD, 2) X(MDB_OP_EXP_BWD, 2) X(MDB_OP_LOG_BWD, 2) X(MDB_OP_RELU_MASK_BWD, 2)     \
  X(MDB_OP_POW_BWD, 3) X(MDB_OP_DIV_BWD_Y, 3) X(MDB_OP_POW_BWD_LIN, 3)

// arrival counters for the single-launch split reductions: 64 Ki zero-initialised words, each reset
// by the CTA that finishes its row / column tile (so launches and CUDA-graph replays never clear them)
constexpr int64_t kTickets = 65536;
static unsigned int* ticket_pool() {
  static unsigned int* pool = nullptr;
  if (!pool) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(g_stream, &st);
    if (st != cudaStreamCaptureStatusNone) return nullptr;      // first use inside a capture: two launches
    if (cudaMalloc(&pool, kTickets * sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); pool = nullptr; return nullptr; }
    cudaMemsetAsync(pool, 0, kTickets * sizeof(unsigned int), g_stream);
  }
  return pool;
}

// operand forms for the specialised kernels; returns false when an operand needs the generic path
static bool classify_forms(const RedPlan& pl, const RedParams& p, int n_in, int (&form)[2]) {
  if (n_in > 2) return false;
  form[0] = form[1] = FV;
  for (int k = 0; k < n_in; ++k) {
    const FastOperand& o = p.in[k];
    if (o.kind == K_IMM) { form[k] = FK; continue; }
    if (o.kind != K_F32) return false;
    if (o.s0 == 1) form[k] = FV;
    else if (pl.pattern == 0) form[k] = FK;                    // row kernel: a scalar per row
    else form[k] = (o.s1 != 0) ? FS : FK;                      // col kernel
  }
  if (n_in == 1) return form[0] == FV;
  return form[0] == FV || form[1] == FV;                       // at least one operand streams
}

template <int OP, int NIN, int RED>
static bool launch_row_forms(const int (&f)[2], dim3 grid, const RedParams& p) {
  if constexpr (NIN == 1) {
    red_row_f32<OP, 1, RED, FV, FV><<<grid, 256, 0, g_stream>>>(p);
    return true;
  } else if constexpr (NIN == 2) {
    if (f[0] == FV && f[1] == FV) red_row_f32<OP, 2, RED, FV, FV><<<grid, 256, 0, g_stream>>>(p);
    else if (f[0] == FV && f[1] == FK) red_row_f32<OP, 2, RED, FV, FK><<<grid, 256, 0, g_stream>>>(p);
    else if (f[0] == FK && f[1] == FV) red_row_f32<OP, 2, RED, FK, FV><<<grid, 256, 0, g_stream>>>(p);
    else return false;
    return true;
  } else {
    return false;
  }
}
// plain launch, or -- when p.csize > 1 -- clusters of csize CTAs along the split axis (grid.y)
template <typename K>
static void launch_col_kernel(K kern, dim3 grid, const RedParams& p) {
  if (p.csize <= 1) {
    kern<<<grid, 256, 0, g_stream>>>(p);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(256, 1, 1); cfg.dynamicSmemBytes = 0; cfg.stream = g_stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 1; attr.val.clusterDim.y = p.csize; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, p);
}
template <int OP, int NIN, int RED, int TX>
static bool launch_col_forms_tx(const int (&f)[2], dim3 grid, const RedParams& p) {
  if constexpr (NIN == 1) {
    launch_col_kernel(red_col_f32<OP, 1, RED, FV, FV, TX>, grid, p);
    return true;
  } else if constexpr (NIN == 2) {
    if (f[0] == FV && f[1] == FV) launch_col_kernel(red_col_f32<OP, 2, RED, FV, FV, TX>, grid, p);
    else if (f[0] == FV && f[1] == FS) launch_col_kernel(red_col_f32<OP, 2, RED, FV, FS, TX>, grid, p);
    else if (f[0] == FS && f[1] == FV) launch_col_kernel(red_col_f32<OP, 2, RED, FS, FV, TX>, grid, p);
    else if (f[0] == FV && f[1] == FK) launch_col_kernel(red_col_f32<OP, 2, RED, FV, FK, TX>, grid, p);
    else if (f[0] == FK && f[1] == FV) launch_col_kernel(red_col_f32<OP, 2, RED, FK, FV, TX>, grid, p);
    else return false;
    return true;
  } else {
    return false;
  }
}
// the narrow tiles are only instantiated for the sums autodiff issues on big operands (plain sums, products,
// the ReLU mask): every other fused form keeps 128-column tiles (compile time)
template <int OP>
constexpr bool kNarrowColTiles = (OP == MDB_OP_COPY || OP == MDB_OP_MUL || OP == MDB_OP_RELU_MASK_BWD);

template <int OP, int NIN, int RED>
static bool launch_col_forms(const int (&f)[2], dim3 grid, const RedParams& p, int tx) {
  if constexpr (kNarrowColTiles<OP>) {
    if (tx == 16) return launch_col_forms_tx<OP, NIN, RED, 16>(f, grid, p);
    if (tx == 8) return launch_col_forms_tx<OP, NIN, RED, 8>(f, grid, p);
  }
  return launch_col_forms_tx<OP, NIN, RED, 32>(f, grid, p);
}

template <int OP, int NIN, int RED>
static int launch_fast_red(const RedPlan& pl, RedParams p, int vec, float* out, bool accumulate,
                           float divisor) {
  p.accumulate = accumulate; p.divisor = divisor;
  p.tickets = nullptr; p.final_dst = nullptr;
  p.csize = 1; p.nparts = 1;
  int form[2];
  static const bool no_forms = getenv("MDB_RED_GENERIC") != nullptr;       // A/B switch for measurements
  const bool forms_ok = !no_forms && vec == 4 && classify_forms(pl, p, NIN, form);
  // LATE kernels: operand 0 streams, every other operand has a footprint of at most 64 KB
  bool late = NIN > 1;
  for (int k = 1; k < NIN; ++k) {
    if (p.in[k].kind == K_IMM) continue;
    int64_t foot = 1;
    const int64_t ext[3] = {pl.d2, pl.d1, pl.d0};
    for (int j = 0; j < 3; ++j) if (pl.istr[k][j] != 0) foot *= ext[j];
    late = late && foot <= 16384;
  }
  if (pl.pattern == 0) {
    const int64_t rows = pl.d2 * pl.d1;
    const uint32_t L = (uint32_t)(pl.d0 / vec);
    p.rows = (uint32_t)rows; p.L = L; p.d1 = (uint32_t)pl.d1; p.div_d1 = FastDiv((uint32_t)pl.d1);
    p.os2 = pl.ostr[0]; p.os1 = pl.ostr[1];
    if (pl.d0 < 2048) {
      p.dst = out; p.to_partial = 0; p.nsplit = 1; p.seg = L;
      int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)g_sm_count * 16);
      if (vec == 4) red_row_warp<OP, NIN, RED, 4><<<grid, 256, 0, g_stream>>>(p);
      else red_row_warp<OP, NIN, RED, 1><<<grid, 256, 0, g_stream>>>(p);
      MDB_CHECK_LAUNCH();
      return 0;
    }
    // long rows: split so that >= ~4 CTAs per SM exist, each split >= 4096 work items
    // many short CTAs (like the 8192-row axis=1 case, 0.94 of peak) beat few long ones
    int64_t want = ((int64_t)g_sm_count * 56 + rows - 1) / rows;   // ~9 waves of 2-iteration CTAs: short tail
    int64_t maxsplit = std::max<int64_t>(1, L / 2048);
    uint32_t nsplit = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(std::min(want, maxsplit), 65535));
    uint32_t seg = (L + nsplit - 1) / nsplit;
    nsplit = (L + seg - 1) / seg;
    p.seg = seg; p.nsplit = nsplit;
    TempBuf tmp;
    if (nsplit > 1) {
      MDB_TRY(tmp.alloc((size_t)rows * nsplit * sizeof(float)));
      p.dst = (float*)tmp.ptr; p.to_partial = 1;
    } else {
      p.dst = out; p.to_partial = 0;
    }
    dim3 grid((unsigned)rows, nsplit);
    if (vec == 4 && forms_ok && nsplit > 1 && rows <= kTickets) {
      p.tickets = ticket_pool();
      p.final_dst = out;
    }
    bool used_forms = false;
    if (vec == 4) {
      if (forms_ok && (used_forms = launch_row_forms<OP, NIN, RED>(form, grid, p))) {}
      else if (NIN > 1 && late) red_row_cta<OP, NIN, RED, 4, (NIN > 1)><<<grid, 256, 0, g_stream>>>(p);
      else red_row_cta<OP, NIN, RED, 4><<<grid, 256, 0, g_stream>>>(p);
    } else {
      red_row_cta<OP, NIN, RED, 1><<<grid, 256, 0, g_stream>>>(p);
    }
    MDB_CHECK_LAUNCH();
    if (nsplit > 1 && !(used_forms && p.tickets)) {  // second pass over the [rows, nsplit] partials
      RedParams q = p;
      q.in[0].ptr = tmp.ptr; q.in[0].kind = K_F32; q.in[0].s0 = 1;
      q.in[0].s1 = (int32_t)nsplit; q.in[0].s2 = (int32_t)((int64_t)nsplit * pl.d1);
      q.L = nsplit; q.seg = nsplit; q.nsplit = 1; q.dst = out; q.to_partial = 0;
      q.rows = (uint32_t)rows;
      if (nsplit <= 256) {          // a warp per row of partials
        int grid2 = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)g_sm_count * 16);
        red_row_warp<MDB_OP_COPY, 1, RED, 1><<<grid2, 256, 0, g_stream>>>(q);
      } else {                      // long rows of partials (full reductions): a CTA per row
        red_row_cta<MDB_OP_COPY, 1, RED, 1><<<dim3((unsigned)rows, 1), 256, 0, g_stream>>>(q);
      }
      MDB_CHECK_LAUNCH();
    }
    return 0;
  }
  // column pattern
  const int64_t R = pl.d1, I = pl.d0, O2 = pl.d2;
  p.L = (uint32_t)R; p.I = (uint32_t)I; p.os2 = pl.ostr[0];
  // f32 form kernels: ONE launch, ONE resident wave.  The 8 CTAs that share a column tile and own consecutive row
  // ranges form a thread-block cluster along the split axis and fold their column sums through distributed shared
  // memory (rank order: deterministic) -- no partial rows in global memory, no second pass, no tickets.  The tile is
  // narrowed (128 / 64 / 32 columns, never less than one 128-B line per row) until >= 3 CTAs per SM exist.
  // Measured on 2^26 elements (profiles/r02_microbench_elementwise.txt): two launches over 3.5 k short CTAs
  // 50.3 us = 0.81 of the HBM peak; clusters over the same short CTAs 60 us (a cluster holds its slots until its
  // slowest member is done, at each of 4 wave boundaries); one wave, 4 rows in flight per thread 47.7 us; 8 rows
  // 46.0 us = 0.89; more than 8 splits (partial rows + tickets among the cluster leaders) 50-58 us.
  // MDB_RED_COL_CLUSTER=0: the two-launch scheme (also used when the rows are too few to split 8 ways).
  static const bool col_cluster = !(getenv("MDB_RED_COL_CLUSTER") && atoi(getenv("MDB_RED_COL_CLUSTER")) == 0);
  int tx = 32;
  uint32_t csize = 1;
  // (inputs of >= 2^27 elements keep the two-launch scheme: its many short CTAs are bandwidth-bound there as well,
  //  and inside the power-capped C4 step it measured 5 % faster than the single wave: 0.458 vs 0.479 ms per step)
  if (col_cluster && vec == 4 && forms_ok && R >= 512 && R * I * O2 < (int64_t(1) << 27)) {
    for (int t : {32, 16, 8}) {
      tx = t;
      if (!kNarrowColTiles<OP> || ((I + 4 * t - 1) / (4 * t)) * O2 * 8 >= (int64_t)g_sm_count * 3) break;
    }
    if (((I + 4 * tx - 1) / (4 * tx)) * O2 * 16 >= (int64_t)g_sm_count * 3) csize = 8;      // >= 1.5 CTAs per SM
    else tx = 32;                                        // too few columns for one wave: split the rows further, two launches
  }
  const int tile = tx * vec;
  const int64_t gx = (I + tile - 1) / tile;
  int64_t want = ((int64_t)g_sm_count * 24 + gx * O2 - 1) / (gx * O2);   // ~4 full waves (8 CTAs/SM gave 1.37 waves: 27 % tail)
  int64_t maxsplit = std::max<int64_t>(1, R / 64);
  uint32_t nsplit = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(std::min(want, maxsplit), 1024));
  if ((int64_t)nsplit * I * O2 >= (int64_t(1) << 31)) nsplit = 1;
  if (csize > 1) nsplit = csize;
  uint32_t seg = (uint32_t)((R + nsplit - 1) / nsplit);
  if (csize == 1) nsplit = (uint32_t)((R + seg - 1) / seg);
  const uint32_t nparts = nsplit / csize;
  p.seg = seg; p.nsplit = nsplit; p.csize = csize; p.nparts = nparts;
  TempBuf tmp;
  if (nparts > 1) {
    MDB_TRY(tmp.alloc((size_t)O2 * nparts * I * sizeof(float)));
    p.dst = (float*)tmp.ptr; p.to_partial = 1;
  } else {
    p.dst = out; p.to_partial = 0;
  }
  dim3 grid((unsigned)gx, nsplit, (unsigned)O2);
  // Per-CTA tickets are NOT used for column sums: measured 52.8 vs 50.6 us on (8192,8192) -- the fence + ticket in
  // each of the 3.5 k short CTAs costs more than the 5 us second launch it saves.  MDB_RED_COL_TICKETS=1 turns the
  // scheme on for measurements.
  static const bool col_tickets = getenv("MDB_RED_COL_TICKETS") != nullptr;
  if (col_tickets && vec == 4 && forms_ok && nparts > 1 && gx * O2 <= kTickets) {
    p.tickets = ticket_pool();
    p.final_dst = out;
  }
  bool used_forms = false;
  if (vec == 4) {
    if (forms_ok && (used_forms = launch_col_forms<OP, NIN, RED>(form, grid, p, tx))) {}
    else if (csize > 1) return MDB_EINVAL;               // unreachable: classify_forms() admits only launchable forms
    else if (NIN > 1 && late) red_col<OP, NIN, RED, 4, (NIN > 1)><<<grid, 256, 0, g_stream>>>(p);
    else red_col<OP, NIN, RED, 4><<<grid, 256, 0, g_stream>>>(p);
  } else {
    red_col<OP, NIN, RED, 1><<<grid, 256, 0, g_stream>>>(p);
  }
  MDB_CHECK_LAUNCH();
  if (nparts > 1 && !(used_forms && p.tickets)) {
    RedParams q = p;
    q.in[0].ptr = tmp.ptr; q.in[0].kind = K_F32; q.in[0].s0 = 1;
    q.in[0].s1 = (int32_t)I; q.in[0].s2 = (int32_t)((int64_t)nparts * I);
    q.L = nparts; q.seg = nparts; q.nsplit = 1; q.csize = 1; q.nparts = 1; q.dst = out; q.to_partial = 0;
    const bool v4 = (vec == 4);  // partial rows are 16B aligned iff I % 4 == 0 (true when vec == 4)
    dim3 grid2((unsigned)gx, 1, (unsigned)O2);
    if (v4) red_col<MDB_OP_COPY, 1, RED, 4><<<grid2, 256, 0, g_stream>>>(q);
    else red_col<MDB_OP_COPY, 1, RED, 1><<<grid2, 256, 0, g_stream>>>(q);
    MDB_CHECK_LAUNCH();
  }
  return 0;
}

static int generic_reduce(int red, const mdb_array* out, const mdb_array* in, uint32_t axis_mask) {
  GenRedParams g;
  g.in = in->ptr; g.out = out->ptr; g.in_dtype = in->dtype; g.out_dtype = out->dtype; g.red = red;
  g.nk = g.nr = 0; g.n_out = 1; g.n_red = 1;
  for (int d = 0; d < in->ndim; ++d) {
    if (axis_mask & (1u << d)) {
      g.rshape[g.nr] = in->shape[d]; g.ristr[g.nr] = in->strides[d]; ++g.nr;
      g.n_red *= in->shape[d];
    } else {
      g.kshape[g.nk] = in->shape[d]; g.kistr[g.nk] = in->strides[d];
      g.kostr[g.nk] = out->strides[d]; ++g.nk;
      g.n_out *= in->shape[d];
    }
  }
  if (g.n_out == 0) return 0;
  if (g.n_red == 0 && (red == MDB_RED_MAX || red == MDB_RED_MIN || red == MDB_RED_ARGMAX ||
                       red == MDB_RED_ARGMIN))
    return set_error(MDB_EINVAL, "zero-size array to reduction operation which has no identity");
  int grid = grid_for(g.n_out, 128);
  const bool int_acc = !dtype_is_float(in->dtype) &&
                       (red != MDB_RED_MEAN) && !dtype_is_float(out->dtype);
  if (int_acc) red_generic<long long><<<grid, 128, 0, g_stream>>>(g);
  else red_generic<double><<<grid, 128, 0, g_stream>>>(g);
  MDB_CHECK_LAUNCH();
  return 0;
}

// shared driver for mdb_reduce (op = COPY) and mdb_elementwise_reduce
static int reduce_driver(int op, int red, const mdb_array* out, int n_in, const mdb_array* in,
                         const int64_t* full, const bool* redax, int nd, bool accumulate) {
  int64_t istr[3][MDB_MAX_DIMS], ostr[MDB_MAX_DIMS];
  int64_t n_red = 1, n_out = 1;
  for (int d = 0; d < nd; ++d) {
    ostr[d] = redax[d] ? 0 : (out->shape[d] == 1 ? 0 : out->strides[d]);
    if (redax[d]) n_red *= full[d]; else n_out *= full[d];
  }
  bool fast_ok = out->dtype == MDB_F32 && (red == MDB_RED_SUM || red == MDB_RED_MEAN ||
                                            red == MDB_RED_MAX || red == MDB_RED_MIN);
  for (int k = 0; k < n_in; ++k) {
    const mdb_array& a = in[k];
    if (a.ptr == nullptr) { for (int d = 0; d < nd; ++d) istr[k][d] = 0; continue; }
    fast_ok = fast_ok && (a.dtype == MDB_F32 || ((a.dtype == MDB_BOOL || a.dtype == MDB_U8) && op != MDB_OP_COPY));
    int lead = nd - a.ndim;
    MDB_REQUIRE(lead >= 0, "reduce: input %d has more dims than the iteration space", k);
    for (int d = 0; d < nd; ++d) {
      if (d < lead) { istr[k][d] = 0; continue; }
      int64_t ext = a.shape[d - lead];
      MDB_REQUIRE(ext == full[d] || ext == 1, "operands could not be broadcast together: extent "
                  "%lld vs %lld on axis %d", (long long)ext, (long long)full[d], d);
      istr[k][d] = (ext == 1) ? 0 : a.strides[d - lead];
    }
  }
  if (n_out == 0) return 0;
  RedPlan pl;
  fast_ok = fast_ok && n_red > 0 && n_red < (int64_t(1) << 31) && n_out < (int64_t(1) << 31) &&
            plan_fast(nd, full, redax, n_in, istr, ostr, &pl);
  if (fast_ok && pl.pattern == 1 && pl.d2 > 65535) fast_ok = false;
  if (fast_ok) {                                   // kernel operands carry 32-bit strides
    const int64_t lim = int64_t(1) << 31;
    for (int k = 0; k < n_in; ++k)
      for (int j = 0; j < 3; ++j) fast_ok = fast_ok && pl.istr[k][j] < lim && pl.istr[k][j] > -lim;
    fast_ok = fast_ok && pl.d2 * pl.d1 * 1024 < lim;   // room for the [rows, nsplit] partials indexing
  }
  if (fast_ok) {
    RedParams p;
    p.aux = 0.f;
    bool v4 = pl.d0 % 4 == 0;
    for (int k = 0; k < n_in; ++k) {
      FastOperand& o = p.in[k];
      o.ptr = in[k].ptr; o.imm = (float)in[k].imm;
      o.kind = in[k].ptr == nullptr ? K_IMM : (in[k].dtype == MDB_F32 ? K_F32 : K_U8);
      o.s2 = (int32_t)pl.istr[k][0]; o.s1 = (int32_t)pl.istr[k][1]; o.s0 = (int)pl.istr[k][2];
      if (o.kind != K_IMM && o.s0 == 1) {
        size_t esz = o.kind == K_F32 ? 4 : 1;
        v4 = v4 && aligned16(o.ptr, 4 * esz) && o.s1 % 4 == 0 && o.s2 % 4 == 0;
      }
    }
    if (op == MDB_OP_POW_BWD) {
      MDB_REQUIRE(in[2].ptr == nullptr, "POW_BWD needs an immediate exponent");
      p.aux = (float)(in[2].imm - 1.0);
      if (in[2].imm == 2.0) op = MDB_OP_POW_BWD_LIN;
    }
    if (pl.pattern == 1) v4 = v4 && aligned16(out->ptr, 16) && pl.ostr[0] % 4 == 0;
    const int vec = v4 ? 4 : 1;
    float divisor = red == MDB_RED_MEAN ? (float)n_red : 0.f;
    int r = red == MDB_RED_MAX ? R_MAX : (red == MDB_RED_MIN ? R_MIN : R_SUM);
    float* o = (float*)out->ptr;
    if (op == MDB_OP_COPY) {
      if (r == R_SUM) return launch_fast_red<MDB_OP_COPY, 1, R_SUM>(pl, p, vec, o, accumulate, divisor);
      if (r == R_MAX) return launch_fast_red<MDB_OP_COPY, 1, R_MAX>(pl, p, vec, o, accumulate, divisor);
      return launch_fast_red<MDB_OP_COPY, 1, R_MIN>(pl, p, vec, o, accumulate, divisor);
    }
    if (r == R_SUM) {
      switch (op) {
#define X(OPID, N) case OPID: return launch_fast_red<OPID, N, R_SUM>(pl, p, vec, o, accumulate, divisor);
        MDB_FUSED_RED_OPS(X)
#undef X
        default: break;
      }
    }
  }
  // ---- fallback: materialise the elementwise result (if any), then generic / staged reduce
  mdb_array src;
  TempBuf tmp;
  if (op != MDB_OP_COPY || n_in != 1 || in[0].ptr == nullptr || in[0].ndim != nd) {
    int64_t n = n_red * n_out;
    src.dtype = out->dtype == MDB_BOOL ? in[0].dtype : out->dtype;
    if (red == MDB_RED_ARGMAX || red == MDB_RED_ARGMIN || red == MDB_RED_ANY || red == MDB_RED_ALL)
      src.dtype = in[0].dtype;
    MDB_TRY(tmp.alloc((size_t)std::max<int64_t>(n, 1) * dtype_size(src.dtype)));
    src.ptr = tmp.ptr; src.ndim = nd; src.imm = 0; src.imm_i = 0;
    int64_t st = 1;
    for (int d = nd - 1; d >= 0; --d) { src.shape[d] = full[d]; src.strides[d] = st; st *= full[d]; }
    MDB_TRY(elementwise_impl(op, &src, n_in, in));
  } else {
    src = in[0];
    for (int d = 0; d < nd; ++d) if (src.shape[d] != full[d]) { src.shape[d] = full[d]; src.strides[d] = 0; }
  }
  uint32_t mask = 0;
  for (int d = 0; d < nd; ++d) if (redax[d]) mask |= 1u << d;
  if (!accumulate) return generic_reduce(red, out, &src, mask);
  // accumulate: reduce into a temp shaped like out, then out += temp
  TempBuf t2;
  mdb_array o2 = *out;
  MDB_TRY(t2.alloc((size_t)std::max<int64_t>(n_out, 1) * dtype_size(out->dtype)));
  o2.ptr = t2.ptr;
  int64_t st = 1;
  for (int d = nd - 1; d >= 0; --d) { o2.strides[d] = st; st *= o2.shape[d]; }
  MDB_TRY(generic_reduce(red, &o2, &src, mask));
  mdb_array pair[2] = {*out, o2};
  return elementwise_impl(MDB_OP_ADD, out, 2, pair);
}

}  // namespace mdb

using namespace mdb;

extern "C" {

int mdb_reduce(int red, const mdb_array* out, const mdb_array* in, uint32_t axis_mask) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(in && in->ptr && out && out->ptr, "reduce: device arrays required");
  MDB_REQUIRE(out->ndim == in->ndim, "reduce: out must be given in keepdims form");
  const int nd = in->ndim;
  bool redax[MDB_MAX_DIMS];
  int64_t full[MDB_MAX_DIMS];
  for (int d = 0; d < nd; ++d) {
    redax[d] = (axis_mask >> d) & 1u;
    full[d] = in->shape[d];
    MDB_REQUIRE(out->shape[d] == (redax[d] ? 1 : in->shape[d]), "reduce: bad output extent on axis %d", d);
  }
  ProfScope prof(PROF_REDUCE, algorithmic_bytes(out, 1, in));
  if (nd == 0) return elementwise_impl(MDB_OP_COPY, out, 1, in);
  return reduce_driver(MDB_OP_COPY, red, out, 1, in, full, redax, nd, false);
}

int mdb_elementwise_reduce(int op, const mdb_array* out, int n_in, const mdb_array* in,
                           int accumulate) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(out && out->ptr && n_in >= 1 && n_in <= 3, "elementwise_reduce: bad arguments");
  MDB_REQUIRE(n_in == op_arity(op), "op %d takes %d inputs, got %d", op, op_arity(op), n_in);
  // iteration space = broadcast of all inputs; out is right-aligned against it, axes where out's
  // extent is 1 (or missing) but the iteration extent is larger are summed away
  int nd = out->ndim;
  for (int k = 0; k < n_in; ++k) if (in[k].ptr && in[k].ndim > nd) nd = in[k].ndim;
  MDB_REQUIRE(nd <= MDB_MAX_DIMS, "too many dimensions");
  int64_t full[MDB_MAX_DIMS];
  for (int d = 0; d < nd; ++d) full[d] = 1;
  for (int k = 0; k < n_in; ++k) {
    if (!in[k].ptr) continue;
    int lead = nd - in[k].ndim;
    for (int d = 0; d < in[k].ndim; ++d) {
      int64_t e = in[k].shape[d];
      if (e != 1) {
        MDB_REQUIRE(full[lead + d] == 1 || full[lead + d] == e,
                    "operands could not be broadcast together on axis %d", lead + d);
        full[lead + d] = e;
      }
    }
  }
  mdb_array o = *out;  // right-align out
  const int olead = nd - out->ndim;
  bool redax[MDB_MAX_DIMS];
  for (int d = 0; d < nd; ++d) {
    int64_t oe = d < olead ? 1 : out->shape[d - olead];
    o.shape[d] = oe;
    o.strides[d] = d < olead ? 0 : out->strides[d - olead];
    MDB_REQUIRE(oe == full[d] || oe == 1, "elementwise_reduce: output extent %lld does not match "
                "iteration extent %lld on axis %d", (long long)oe, (long long)full[d], d);
    redax[d] = (oe == 1 && full[d] != 1);
  }
  o.ndim = nd;
  ProfScope prof(PROF_REDUCE, algorithmic_bytes(out, n_in, in) +
                                  (accumulate ? algorithmic_bytes(out, 0, nullptr) : 0.0));
  return reduce_driver(op, MDB_RED_SUM, &o, n_in, in, full, redax, nd, accumulate != 0);
}

}  // extern "C"
