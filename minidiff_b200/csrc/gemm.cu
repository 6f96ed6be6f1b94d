// matmul (backend/numpy.py:84) and its two gradient GEMMs (ops/definitions.py:487-492):
// dispatcher + the fp32 CUDA-core kernel used for shapes the tcgen05 path does not take
// (small / unaligned operands, e.g. the reference tests' 10x30 @ 30x20).
#include <algorithm>

#include "mdb_common.cuh"

namespace mdb {

int gemm_tcgen05(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate, const GemmEpilogue* epi);
extern int g_last_plan[8];
extern int g_knob_raster, g_knob_group, g_knob_hint_a, g_knob_hint_b, g_knob_hint_c, g_knob_streamk, g_knob_l2_budget_mb, g_knob_max_clusters, g_knob_split, g_knob_chunk, g_knob_rz_gain;
static int g_force_path = 0;
extern int g_gemm_flags;
// launches per GEMM kernel since the last reset (mdb_gemm_stats): tests assert with it that a
// workload really ran on the kernel it is meant to cover (a silent fallback cannot hide)
uint64_t g_gemm_path[MDB_GEMM_NPATHS] = {};

// 64x64 output tile per CTA, K step 16, 4x4 micro-tile per thread; operands addressed through
// (row, col) element strides so NN / NT / TN views need no copies.
template <bool ACC>
__global__ void __launch_bounds__(256) sgemm_simt(int M, int N, int K, const float* __restrict__ A,
                                                  int64_t sam, int64_t sak,
                                                  const float* __restrict__ B, int64_t sbk,
                                                  int64_t sbn, float* C, int64_t scm, int64_t scn) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
  float acc[4][4] = {};
  // loader mapping: pick the thread->element order that walks the operand's unit-stride axis
  const bool a_k_fast = (sak == 1);
  const bool b_n_fast = (sbn == 1);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;  // 1024 elements of the 64x16 A tile
      int mm = a_k_fast ? e / BK : e % BM;
      int kk = a_k_fast ? e % BK : e / BM;
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? A[(int64_t)gm * sam + (int64_t)gk * sak] : 0.f;
      int nn = b_n_fast ? e % BN : e / BK;
      int kb = b_n_fast ? e / BN : e % BK;
      int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < N && gkb < K) ? B[(int64_t)gkb * sbk + (int64_t)gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tr + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tc + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gm = m0 + tr + i, gn = n0 + tc + j;
      if (gm < M && gn < N) {
        float* p = C + (int64_t)gm * scm + (int64_t)gn * scn;
        *p = ACC ? __fadd_rn(*p, acc[i][j]) : acc[i][j];
      }
    }
}

// Batched / double-precision matmul on the CUDA cores: the same 64x64x16 tiling for T = float or double,
// every operand addressed through element strides (views, broadcast batch axes with stride 0), all
// matrices of the batch in ONE launch (grid.z walks the flattened batch).  Stands behind matmul for
// float64 / integer operands (NumPy's default dtype: Tensor(python floats), zeros, rand ...) and for
// stacked operands (backend/numpy.py:84 is np.matmul, which broadcasts leading axes).
struct BatchParams {
  int M, N, K, nbatch_dims;
  int64_t bshape[MDB_MAX_DIMS], sa[MDB_MAX_DIMS], sb[MDB_MAX_DIMS], sc[MDB_MAX_DIMS];   // batch axes
  int64_t sam, sak, sbk, sbn, scm, scn;
  int64_t batches;
};
template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_batched(const BatchParams p, const T* __restrict__ A,
                                                         const T* __restrict__ B, T* C) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
  const bool a_k_fast = (p.sak == 1), b_n_fast = (p.sbn == 1);
  for (int64_t batch = blockIdx.z; batch < p.batches; batch += gridDim.z) {
    int64_t rem = batch, oa = 0, ob = 0, oc = 0;
    for (int d = p.nbatch_dims - 1; d >= 0; --d) {
      const int64_t q = rem / p.bshape[d], i = rem - q * p.bshape[d];
      rem = q;
      oa += i * p.sa[d]; ob += i * p.sb[d]; oc += i * p.sc[d];
    }
    const T* Ab = A + oa;
    const T* Bb = B + ob;
    T* Cb = C + oc;
    T acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        const int mm = a_k_fast ? e / BK : e % BM, kk = a_k_fast ? e % BK : e / BM;
        const int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < p.M && gk < p.K) ? Ab[(int64_t)gm * p.sam + (int64_t)gk * p.sak] : T(0);
        const int nn = b_n_fast ? e % BN : e / BK, kb = b_n_fast ? e / BN : e % BK;
        const int gn = n0 + nn, gkb = k0 + kb;
        Bs[kb][nn] = (gn < p.N && gkb < p.K) ? Bb[(int64_t)gkb * p.sbk + (int64_t)gn * p.sbn] : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        T a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][tr + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tc + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gm = m0 + tr + i, gn = n0 + tc + j;
        if (gm < p.M && gn < p.N) Cb[(int64_t)gm * p.scm + (int64_t)gn * p.scn] = acc[i][j];
      }
  }
}

// fp32 variant for matrices of at least 128 x 128: 128 x 128 x 8 tiles, 8 x 8 outputs per thread read from
// shared memory as float4, the next k-tile prefetched into registers while the current one is used.
// (The 64 x 64 kernel above measured 1.5 TFLOP/s on 64 stacked 256^3 products: 4 x 4 outputs per thread
// leave it bound by shared-memory loads.)
__global__ void __launch_bounds__(256) gemm_simt_batched_f32_128(const BatchParams p, const float* __restrict__ A,
                                                                 const float* __restrict__ B, float* C) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tr = (tid / 16) * 8, tc = (tid % 16) * 8;
  const bool a_k_fast = (p.sak == 1), b_n_fast = (p.sbn == 1);
  // loader mapping: 1024 elements per operand tile, 4 per thread, walking the operand's unit-stride axis
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + i * 256;
    am[i] = a_k_fast ? e / BK : e % BM; ak[i] = a_k_fast ? e % BK : e / BM;
    bn[i] = b_n_fast ? e % BN : e / BK; bk[i] = b_n_fast ? e / BN : e % BK;
  }
  for (int64_t batch = blockIdx.z; batch < p.batches; batch += gridDim.z) {
    int64_t rem = batch, oa = 0, ob = 0, oc = 0;
    for (int d = p.nbatch_dims - 1; d >= 0; --d) {
      const int64_t q = rem / p.bshape[d], i = rem - q * p.bshape[d];
      rem = q;
      oa += i * p.sa[d]; ob += i * p.sb[d]; oc += i * p.sc[d];
    }
    const float* Ab = A + oa;
    const float* Bb = B + ob;
    float* Cb = C + oc;
    float acc[8][8] = {};
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gm = m0 + am[i], gk = k0 + ak[i];
        ra[i] = (gm < p.M && gk < p.K) ? Ab[(int64_t)gm * p.sam + (int64_t)gk * p.sak] : 0.f;
        const int gn = n0 + bn[i], gkb = k0 + bk[i];
        rb[i] = (gn < p.N && gkb < p.K) ? Bb[(int64_t)gkb * p.sbk + (int64_t)gn * p.sbn] : 0.f;
      }
    };
    auto stash = [&](int buf) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { As[buf][ak[i]][am[i]] = ra[i]; Bs[buf][bk[i]][bn[i]] = rb[i]; }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < p.K; k0 += BK) {
      const bool more = k0 + BK < p.K;
      if (more) fetch(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a0 = *(const float4*)&As[buf][kk][tr], a1 = *(const float4*)&As[buf][kk][tr + 4];
        const float4 b0 = *(const float4*)&Bs[buf][kk][tc], b1 = *(const float4*)&Bs[buf][kk][tc + 4];
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (more) stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gm = m0 + tr + i, gn = n0 + tc + j;
        if (gm < p.M && gn < p.N) Cb[(int64_t)gm * p.scm + (int64_t)gn * p.scn] = acc[i][j];
      }
    __syncthreads();
  }
}

static int gemm_simt(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate) {
  const int M = (int)a->shape[0], K = (int)a->shape[1], N = (int)b->shape[1];
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  if (accumulate)
    sgemm_simt<true><<<grid, 256, 0, g_stream>>>(M, N, K, (const float*)a->ptr, a->strides[0],
        a->strides[1], (const float*)b->ptr, b->strides[0], b->strides[1], (float*)c->ptr,
        c->strides[0], c->strides[1]);
  else
    sgemm_simt<false><<<grid, 256, 0, g_stream>>>(M, N, K, (const float*)a->ptr, a->strides[0],
        a->strides[1], (const float*)b->ptr, b->strides[0], b->strides[1], (float*)c->ptr,
        c->strides[0], c->strides[1]);
  MDB_CHECK_LAUNCH();
  ++g_gemm_path[MDB_GEMM_PATH_SIMT];
  return 0;
}

}  // namespace mdb

using namespace mdb;

extern "C" {

int mdb_gemm_config(int force_path) {
  MDB_REQUIRE(force_path >= 0 && force_path <= 2, "force_path must be 0, 1 or 2");
  g_force_path = force_path;
  return 0;
}

int mdb_gemm_stats(uint64_t* counts, int reset) {
  if (counts)
    for (int i = 0; i < MDB_GEMM_NPATHS; ++i) counts[i] = g_gemm_path[i];
  if (reset)
    for (int i = 0; i < MDB_GEMM_NPATHS; ++i) g_gemm_path[i] = 0;
  return 0;
}

int mdb_gemm_last_plan(int* out8) {
  for (int i = 0; i < 8; ++i) out8[i] = g_last_plan[i];
  return 0;
}

int mdb_gemm_tune(int flags) {
  g_gemm_flags = flags;
  return 0;
}

int mdb_gemm_knob(int knob, int value) {
  switch (knob) {
    case MDB_GEMM_KNOB_RASTER: g_knob_raster = value; break;
    case MDB_GEMM_KNOB_GROUP: g_knob_group = value; break;
    case MDB_GEMM_KNOB_HINT_A: g_knob_hint_a = value; break;
    case MDB_GEMM_KNOB_HINT_B: g_knob_hint_b = value; break;
    case MDB_GEMM_KNOB_HINT_C: g_knob_hint_c = value; break;
    case MDB_GEMM_KNOB_STREAMK: g_knob_streamk = value; break;
    case MDB_GEMM_KNOB_MAX_CLUSTERS: g_knob_max_clusters = value; break;
    case MDB_GEMM_KNOB_SPLIT: g_knob_split = value; break;
    case MDB_GEMM_KNOB_CHUNK: g_knob_chunk = value; break;
    case MDB_GEMM_KNOB_RZ_GAIN: g_knob_rz_gain = value; break;
    case MDB_GEMM_KNOB_L2_BUDGET_MB: g_knob_l2_budget_mb = value > 0 ? value : 32; break;
    default: return set_error(MDB_EINVAL, "unknown GEMM knob %d", knob);
  }
  return 0;
}

static int check_gemm_args(const mdb_array* c, const mdb_array* a, const mdb_array* b) {
  MDB_REQUIRE(a && b && c && a->ptr && b->ptr && c->ptr, "gemm: device arrays required");
  MDB_REQUIRE(a->ndim == 2 && b->ndim == 2 && c->ndim == 2, "gemm: operands must be 2-D");
  MDB_REQUIRE(a->dtype == MDB_F32 && b->dtype == MDB_F32 && c->dtype == MDB_F32,
              "gemm: fp32 operands required");
  if (a->shape[1] != b->shape[0])
    return set_error(MDB_EINVAL, "matmul: Input operand 1 has a mismatch in its core dimension 0, "
                     "with gufunc signature (n?,k),(k,m?)->(n?,m?) (size %lld is different from %lld)",
                     (long long)b->shape[0], (long long)a->shape[1]);
  MDB_REQUIRE(c->shape[0] == a->shape[0] && c->shape[1] == b->shape[1], "gemm: bad output shape");
  MDB_REQUIRE(a->shape[0] < (1ll << 31) && a->shape[1] < (1ll << 31) && b->shape[1] < (1ll << 31),
              "gemm: extents must fit in int32");
  return 0;
}

int mdb_gemm_fused(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate,
                   const mdb_array* bias, int relu, const mdb_array* mask_src) {
  MDB_TRY(ensure_init());
  MDB_TRY(check_gemm_args(c, a, b));
  MDB_REQUIRE(c->shape[0] > 0 && c->shape[1] > 0 && a->shape[1] > 0, "gemm_fused: empty operands");
  GemmEpilogue epi = {nullptr, relu ? 1 : 0, nullptr, 0};
  if (bias) {
    MDB_REQUIRE(bias->ptr && bias->dtype == MDB_F32 && bias->ndim == 1 && bias->shape[0] == c->shape[1] &&
                bias->strides[0] == 1, "gemm_fused: bias must be a contiguous fp32 vector of length N");
    epi.bias = (const float*)bias->ptr;
  }
  if (mask_src) {
    MDB_REQUIRE(mask_src->ptr && mask_src->dtype == MDB_F32 && mask_src->ndim == 2 &&
                mask_src->shape[0] == c->shape[0] && mask_src->shape[1] == c->shape[1] && mask_src->strides[1] == 1,
                "gemm_fused: mask source must be a row-major fp32 matrix of C's shape");
    epi.mask_src = (const float*)mask_src->ptr;
    epi.ld_mask = mask_src->strides[0];
  }
  ProfScope prof(PROF_GEMM, 2.0 * (double)a->shape[0] * (double)a->shape[1] * (double)b->shape[1]);
  return gemm_tcgen05(c, a, b, accumulate, &epi);
}

int mdb_gemm_batched(const mdb_array* c, const mdb_array* a, const mdb_array* b) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(a && b && c && a->ptr && b->ptr && c->ptr, "gemm_batched: device arrays required");
  MDB_REQUIRE(a->ndim >= 2 && a->ndim == b->ndim && a->ndim == c->ndim, "gemm_batched: operands must share their rank (>= 2)");
  MDB_REQUIRE(a->dtype == b->dtype && a->dtype == c->dtype && (a->dtype == MDB_F32 || a->dtype == MDB_F64),
              "gemm_batched: float32 or float64 operands of one dtype required");
  const int nd = a->ndim, nb = nd - 2;
  if (a->shape[nd - 1] != b->shape[nd - 2])
    return set_error(MDB_EINVAL, "matmul: Input operand 1 has a mismatch in its core dimension 0, "
                     "with gufunc signature (n?,k),(k,m?)->(n?,m?) (size %lld is different from %lld)",
                     (long long)b->shape[nd - 2], (long long)a->shape[nd - 1]);
  MDB_REQUIRE(c->shape[nd - 2] == a->shape[nd - 2] && c->shape[nd - 1] == b->shape[nd - 1], "gemm_batched: bad output shape");
  BatchParams p;
  p.M = (int)a->shape[nd - 2]; p.K = (int)a->shape[nd - 1]; p.N = (int)b->shape[nd - 1];
  MDB_REQUIRE(a->shape[nd - 2] < (1ll << 31) && a->shape[nd - 1] < (1ll << 31) && b->shape[nd - 1] < (1ll << 31),
              "gemm_batched: extents must fit in int32");
  p.nbatch_dims = nb; p.batches = 1;
  for (int d = 0; d < nb; ++d) {
    const int64_t e = c->shape[d];
    MDB_REQUIRE((a->shape[d] == e || a->shape[d] == 1) && (b->shape[d] == e || b->shape[d] == 1),
                "gemm_batched: batch axis %d does not broadcast", d);
    p.bshape[d] = e;
    p.sa[d] = a->shape[d] == e ? a->strides[d] : 0;
    p.sb[d] = b->shape[d] == e ? b->strides[d] : 0;
    p.sc[d] = c->strides[d];
    p.batches *= e;
  }
  p.sam = a->strides[nd - 2]; p.sak = a->strides[nd - 1];
  p.sbk = b->strides[nd - 2]; p.sbn = b->strides[nd - 1];
  p.scm = c->strides[nd - 2]; p.scn = c->strides[nd - 1];
  if (p.batches == 0 || p.M == 0 || p.N == 0) return 0;
  ProfScope prof(PROF_GEMM, 2.0 * (double)p.batches * p.M * (double)p.K * p.N);
  dim3 grid((p.N + 63) / 64, (p.M + 63) / 64, (unsigned)std::min<int64_t>(p.batches, 32768));
  if (a->dtype == MDB_F32 && p.M >= 128 && p.N >= 128) {
    dim3 grid128((p.N + 127) / 128, (p.M + 127) / 128, (unsigned)std::min<int64_t>(p.batches, 32768));
    gemm_simt_batched_f32_128<<<grid128, 256, 0, g_stream>>>(p, (const float*)a->ptr, (const float*)b->ptr, (float*)c->ptr);
  } else if (a->dtype == MDB_F32)
    gemm_simt_batched<float><<<grid, 256, 0, g_stream>>>(p, (const float*)a->ptr, (const float*)b->ptr, (float*)c->ptr);
  else
    gemm_simt_batched<double><<<grid, 256, 0, g_stream>>>(p, (const double*)a->ptr, (const double*)b->ptr, (double*)c->ptr);
  MDB_CHECK_LAUNCH();
  ++g_gemm_path[MDB_GEMM_PATH_SIMT_BATCHED];
  return 0;
}

int mdb_gemm(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate) {
  MDB_TRY(ensure_init());
  MDB_REQUIRE(a && b && c && a->ptr && b->ptr && c->ptr, "gemm: device arrays required");
  MDB_REQUIRE(a->ndim == 2 && b->ndim == 2 && c->ndim == 2, "gemm: operands must be 2-D");
  MDB_REQUIRE(a->dtype == MDB_F32 && b->dtype == MDB_F32 && c->dtype == MDB_F32,
              "gemm: fp32 operands required");
  if (a->shape[1] != b->shape[0])
    return set_error(MDB_EINVAL, "matmul: Input operand 1 has a mismatch in its core dimension 0, "
                     "with gufunc signature (n?,k),(k,m?)->(n?,m?) (size %lld is different from %lld)",
                     (long long)b->shape[0], (long long)a->shape[1]);
  MDB_REQUIRE(c->shape[0] == a->shape[0] && c->shape[1] == b->shape[1], "gemm: bad output shape");
  MDB_REQUIRE(a->shape[0] < (1ll << 31) && a->shape[1] < (1ll << 31) && b->shape[1] < (1ll << 31),
              "gemm: extents must fit in int32");
  if (c->shape[0] == 0 || c->shape[1] == 0) return 0;
  if (a->shape[1] == 0) return accumulate ? 0 : mdb_fill(c, 0.0);
  ProfScope prof(PROF_GEMM, 2.0 * (double)a->shape[0] * (double)a->shape[1] * (double)b->shape[1]);
  if (g_force_path != 1) {
    int rc = gemm_tcgen05(c, a, b, accumulate, nullptr);
    if (rc == 0) return 0;
    if (rc != MDB_ENOTSUP || g_force_path == 2) return rc;
  }
  return gemm_simt(c, a, b, accumulate);
}

}  // extern "C"
