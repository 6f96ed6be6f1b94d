// fp32 GEMM on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), 3xTF32 split.
//
//   C[M,N] (+)= A[M,K] @ B[K,N]      A, B: any 2-D fp32 views with one unit stride each
//
// stands behind backend.matmul (reference backend/numpy.py:84 -> OpenBLAS sgemm) and its two
// gradient GEMMs dA = dC @ B^T, dB = A^T @ dC (ops/definitions.py:487-492), whose operands arrive
// as transposed *views*: nothing is copied, the operand's major-ness selects the UMMA descriptor:
//     A: stride_k == 1 -> K-major     stride_m == 1 -> MN-major
//     B: stride_k == 1 -> K-major     stride_n == 1 -> MN-major      (B is consumed as N x K)
//
// 3xTF32 (north_star): x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi); the tensor cores
// accumulate  hi*hi + hi*lo + lo*hi  in fp32 TMEM, dropping only lo*lo (2^-22 relative).
// The split is done once per operand by a streaming pre-pass into compact hi/lo planes
// (v1; the planes are what TMA loads).
//
// Accumulation: the tensor core adds into the fp32 TMEM accumulator with TRUNCATION (measured here:
// error grew linearly with K, 1.7e-2 at K=8192 when one TMEM chain ran over all of K).  So a chain
// never runs longer than kChunk k-blocks (128 K): after each chunk the accumulator is handed to
// the epilogue warps, which add it into fp32 REGISTERS with round-to-nearest, while the MMA warp
// already fills the other TMEM buffer (same promotion idea as Ootomo & Yokota's tensor-core SGEMM).
//
// Kernel anatomy (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer : cp.async.bulk.tensor (SWIZZLE_128B) of A_hi/A_lo/B_hi/B_lo k-blocks
//                              into a kStages-deep shared-memory ring, mbarrier complete_tx
//   warp 1      MMA issuer   : one lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8),
//                              3 MMAs per k-step; tcgen05.commit releases ring slots / publishes
//                              the accumulator
//   warp 2      TMEM allocator
//   warps 4..7  epilogue     : tcgen05.ld (32 lanes x 32 columns per warp-instruction) -> registers
//                              -> global (optionally C += ...), double-buffered TMEM accumulators so
//                              the epilogue of tile i overlaps the main loop of tile i+1
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "mdb_common.cuh"

namespace mdb {

namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 32;           // fp32 elements per k-block == one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes of K per instruction
constexpr int kThreads = 256;
constexpr int kChunk = 4;        // k-blocks per in-TMEM accumulation chain (128 K) before promotion

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Spin on a phase parity.  A deadlock (protocol bug) traps after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spins == 64) t0 = clock64();
    if (spins > 64 && (spins & 1023) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4   [16,30) LBO>>4   [32,46) SBO>>4   [46,48) version=1   [61,64) layout type
// layout type 2 = SWIZZLE_128B (K-major operands: 16-B chunks XOR row%8, 8-row / 1024-B atoms)
// layout type 1 = SWIZZLE_128B_BASE32B (the ONLY layout the tensor core accepts for MN-major
//                 32-bit operands: 32-B chunks XOR row%4, 4-row / 512-B atoms; TMA writes it with
//                 CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

#define MDB_TMEM_LD32(taddr, r)                                                                          \
  asm volatile(                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                          \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),         \
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),       \
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),       \
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                            \
      : "r"(taddr)                                                                                       \
      : "memory")

struct Params {
  int M, N, K;
  int a_mn_major, b_mn_major;   // operand major-ness (0: K-major, 1: MN-major)
  float* C;
  int64_t ldc;
  int accumulate;
  int tiles_m, tiles_n, group_m;
  int debug;   // MDB_GEMM_DEBUG: 1 = epilogue stores a sentinel instead of the result
};

// smem ring: per stage [A_hi | A_lo | B_hi | B_lo], each operand tile is (rows x 128 B), 1024-B atoms
template <int BN, int kStages>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int RING_BYTES = kStages * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = RING_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

__device__ __forceinline__ void tile_coords(const Params& p, int t, int& m_blk, int& n_blk) {
  // groups of `group_m` tile-rows are walked column by column: neighbouring CTAs share B panels
  // and a small set of A panels (L2 reuse across the wave)
  const int per_group = p.group_m * p.tiles_n;
  const int g = t / per_group, r = t - g * per_group;
  const int rows = min(p.group_m, p.tiles_m - g * p.group_m);
  m_blk = g * p.group_m + (r % rows);
  n_blk = r / rows;
}

template <int BN, int kStages>
__global__ void __launch_bounds__(kThreads, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const Params p) {
  using S = Smem<BN, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 atoms: 1024-B aligned
  uint64_t* full_bar = (uint64_t*)(smem + S::RING_BYTES);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BN;         // two accumulator stages (power of two >= 32)
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_k = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(p, t, m_blk, n_blk);
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * S::STAGE_BYTES;
          const uint32_t a_hi = smem_u32(st), a_lo = a_hi + S::A_BYTES;
          const uint32_t b_hi = a_lo + S::A_BYTES, b_lo = b_hi + S::B_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          const int k0 = kb * BK;
          if (!p.a_mn_major) {           // plane is [M][K], K contiguous: one (32 x BM) box
            tma_load_2d(a_hi, &map_a_hi, &full_bar[stage], k0, m0);
            tma_load_2d(a_lo, &map_a_lo, &full_bar[stage], k0, m0);
          } else {                       // plane is [K][M], M contiguous: BM/32 boxes of (32 x 32)
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) {
              tma_load_2d(a_hi + c * 4096, &map_a_hi, &full_bar[stage], m0 + 32 * c, k0);
              tma_load_2d(a_lo + c * 4096, &map_a_lo, &full_bar[stage], m0 + 32 * c, k0);
            }
          }
          if (!p.b_mn_major) {           // plane is [N][K], K contiguous
            tma_load_2d(b_hi, &map_b_hi, &full_bar[stage], k0, n0);
            tma_load_2d(b_lo, &map_b_lo, &full_bar[stage], k0, n0);
          } else {                       // plane is [K][N], N contiguous
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
              tma_load_2d(b_hi + c * 4096, &map_b_hi, &full_bar[stage], n0 + 32 * c, k0);
              tma_load_2d(b_lo + c * 4096, &map_b_lo, &full_bar[stage], n0 + 32 * c, k0);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=TF32 [7,10), b=TF32
    // [10,13), a_major bit 15, b_major bit 16, N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) |
                           ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    // K-major  : rows of 128 B, 8-row atoms 1024 B apart (SBO); a k-step advances 32 B inside the row
    // MN-major : 32-element column chunks 4096 B apart (LBO), 4-k-row atoms 512 B apart (SBO);
    //            a k-step (8 k-rows = 2 atoms) advances 1024 B
    const uint32_t a_lbo = p.a_mn_major ? 4096 : 16, b_lbo = p.b_mn_major ? 4096 : 16;
    const uint32_t a_sbo = p.a_mn_major ? 512 : 1024, b_sbo = p.b_mn_major ? 512 : 1024;
    const uint32_t a_lt = p.a_mn_major ? 1 : 2, b_lt = p.b_mn_major ? 1 : 2;
    const uint32_t a_kstep = p.a_mn_major ? 1024 : UMMA_K * 4, b_kstep = p.b_mn_major ? 1024 : UMMA_K * 4;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_k; ++kb) {
        const bool chunk_start = (kb % kChunk) == 0;
        const bool chunk_end = ((kb + 1) % kChunk) == 0 || kb == num_k - 1;
        if (chunk_start) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);    // epilogue has drained this accumulator
          tcgen05_fence_after();
        }
        const uint32_t tmem_d = tmem_base + acc * BN;
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          uint8_t* st = smem + stage * S::STAGE_BYTES;
          const uint32_t a_hi = smem_u32(st), a_lo = a_hi + S::A_BYTES;
          const uint32_t b_hi = a_lo + S::A_BYTES, b_lo = b_hi + S::B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da_hi = make_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t da_lo = make_desc(a_lo + k * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t db_hi = make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
            const uint64_t db_lo = make_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
            umma_tf32(tmem_d, da_lo, db_hi, idesc, !(chunk_start && k == 0));   // small terms first
            umma_tf32(tmem_d, da_hi, db_lo, idesc, 1);
            umma_tf32(tmem_d, da_hi, db_hi, idesc, 1);
          }
          umma_commit(&empty_bar[stage]);                 // ring slot free once these MMAs retire
          if (chunk_end) umma_commit(&tmem_full[acc]);    // chunk accumulator ready for promotion
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    const int q = warp & 3;                               // TMEM lane quarter this warp may touch
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (p.ldc % 4 == 0) && (((uintptr_t)p.C & 15) == 0);
    const int num_chunks = (num_k + kChunk - 1) / kChunk;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(p, t, m_blk, n_blk);
      const int row = m_blk * BM + q * 32 + lane;
      const int n0 = n_blk * BN;
      float sum[BN];                                       // this thread's row of the C tile
#pragma unroll
      for (int j = 0; j < BN; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32);
          MDB_TMEM_LD32(taddr, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] = __fadd_rn(sum[c * 32 + j], __uint_as_float(r[j]));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);      // MMA warp may overwrite this buffer
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (p.debug == 1) {
#pragma unroll
        for (int j = 0; j < BN; ++j) sum[j] = 7.0f;
      }
      if (row < p.M) {
        float* crow = p.C + (int64_t)row * p.ldc;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(sum[c * 32 + j], sum[c * 32 + j + 1], sum[c * 32 + j + 2], sum[c * 32 + j + 3]);
              float4* dst = (float4*)(crow + col0 + j);
              if (p.accumulate) {
                const float4 o = *dst;
                v = make_float4(__fadd_rn(o.x, v.x), __fadd_rn(o.y, v.y), __fadd_rn(o.z, v.z), __fadd_rn(o.w, v.w));
              }
              *dst = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) {
                float v = sum[c * 32 + j];
                if (p.accumulate) v = __fadd_rn(crow[col0 + j], v);
                crow[col0 + j] = v;
              }
          }
        }
      }
    }
  }

  // ------------------------------------------ teardown ---------------------------------------
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- operand pre-pass: hi = rn_tf32(x), lo = rn_tf32(x - hi) into compact planes ----------------
// The plane keeps the operand's memory order: [outer][inner] with `inner` the unit-stride axis.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, int64_t s_outer,
                                                         int64_t s_inner, int outer, int inner, int ld,
                                                         float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = (int64_t)outer * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / ld), c = (int)(i - (int64_t)o * ld);
    float h = 0.f, l = 0.f;
    if (c < inner) {
      const float x = src[(int64_t)o * s_outer + (int64_t)c * s_inner];
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
      h = __uint_as_float(hb);
      const float rem = __fsub_rn(x, h);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
      l = __uint_as_float(lb);
      if (!(fabsf(x) < INFINITY)) { h = x; l = 0.f; }   // keep inf / nan in the hi plane only
    }
    hi[i] = h;
    lo[i] = l;
  }
}

}  // namespace tc

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    return set_error(MDB_ECUDA, "cuTensorMapEncodeTiled is unavailable in this driver");
  }
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

// 2-D fp32 plane [outer][inner] (inner contiguous, row pitch ld floats), box = (32 x box_rows)
static int make_map(CUtensorMap* map, const float* base, int inner, int outer, int ld, int box_rows,
                    bool mn_major) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MDB_ECUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return 0;
}

struct Plane {      // compact hi/lo copy of one operand, in the operand's memory order
  TempBuf hi, lo;
  int inner = 0, outer = 0, ld = 0;
  bool mn_major = false;
};

// mn = extent of the operand's M (or N) axis, k = extent of K; s_mn / s_k its element strides
static int split_operand(const float* src, int mn, int k, int64_t s_mn, int64_t s_k, Plane* pl) {
  if (s_k != 1 && s_mn == 1) {                       // MN-major plane [k][mn]
    pl->mn_major = true; pl->outer = k; pl->inner = mn;
  } else {                                           // K-major plane [mn][k] (also the gather
    pl->mn_major = false; pl->outer = mn; pl->inner = k;   // target for doubly-strided views)
  }
  pl->ld = (pl->inner + 3) & ~3;
  const size_t bytes = (size_t)pl->outer * pl->ld * sizeof(float);
  MDB_TRY(pl->hi.alloc(bytes));
  MDB_TRY(pl->lo.alloc(bytes));
  const int64_t s_outer = pl->mn_major ? s_k : s_mn, s_inner = pl->mn_major ? s_mn : s_k;
  const int64_t total = (int64_t)pl->outer * pl->ld;
  tc::split_tf32_kernel<<<grid_for(total, 256), 256, 0, g_stream>>>(src, s_outer, s_inner, pl->outer, pl->inner,
                                                                  pl->ld, (float*)pl->hi.ptr, (float*)pl->lo.ptr);
  MDB_CHECK_LAUNCH();
  return 0;
}

template <int BN, int kStages>
static int launch(const CUtensorMap maps[4], const tc::Params& p) {
  using S = tc::Smem<BN, kStages>;
  auto kern = tc::gemm_3xtf32_kernel<BN, kStages>;
  static bool configured = false;
  if (!configured) {
    MDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = std::min(tiles, g_sm_count);
  kern<<<grid, tc::kThreads, S::TOTAL, g_stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  MDB_CHECK_LAUNCH();
  return 0;
}

int gemm_tcgen05(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate) {
  const int64_t M = a->shape[0], K = a->shape[1], N = b->shape[1];
  // small problems are launch-latency bound: the CUDA-core kernel is as fast and needs no pre-pass
  if (M * N * K < (int64_t(1) << 21) || K < 32 || M < 32 || N < 32)
    return set_error(MDB_ENOTSUP, "problem too small for the tensor-core path");
  if (c->strides[1] != 1)
    return set_error(MDB_ENOTSUP, "output must be row-major for the tensor-core path");
  MDB_TRY(load_encode());

  Plane pa, pb;
  MDB_TRY(split_operand((const float*)a->ptr, (int)M, (int)K, a->strides[0], a->strides[1], &pa));
  MDB_TRY(split_operand((const float*)b->ptr, (int)N, (int)K, b->strides[1], b->strides[0], &pb));

  constexpr int BN = 128, kStages = 3;
  CUtensorMap maps[4];
  const int a_box = pa.mn_major ? 32 : tc::BM, b_box = pb.mn_major ? 32 : BN;
  MDB_TRY(make_map(&maps[0], (const float*)pa.hi.ptr, pa.inner, pa.outer, pa.ld, a_box, pa.mn_major));
  MDB_TRY(make_map(&maps[1], (const float*)pa.lo.ptr, pa.inner, pa.outer, pa.ld, a_box, pa.mn_major));
  MDB_TRY(make_map(&maps[2], (const float*)pb.hi.ptr, pb.inner, pb.outer, pb.ld, b_box, pb.mn_major));
  MDB_TRY(make_map(&maps[3], (const float*)pb.lo.ptr, pb.inner, pb.outer, pb.ld, b_box, pb.mn_major));

  tc::Params p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.a_mn_major = pa.mn_major; p.b_mn_major = pb.mn_major;
  p.C = (float*)c->ptr; p.ldc = c->strides[0];
  p.accumulate = accumulate;
  p.tiles_m = (int)((M + tc::BM - 1) / tc::BM);
  p.tiles_n = (int)((N + BN - 1) / BN);
  p.group_m = 16;
  const char* dbg = getenv("MDB_GEMM_DEBUG");
  p.debug = dbg ? atoi(dbg) : 0;
  return launch<BN, kStages>(maps, p);
}

}  // namespace mdb
