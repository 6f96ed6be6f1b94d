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
// 3xTF32 (north_star): x = hi + lo, the tensor cores accumulate  lo*hi + hi*lo + hi*hi  in fp32 TMEM,
// dropping only lo*lo (~2^-20 relative).
//
// RAW mode (the fast path).  Measured on this hardware: kind::tf32 simply IGNORES the low 13 mantissa
// bits of its operands (feeding raw fp32 as the hi plane keeps fp32-class accuracy), so a raw fp32
// tile is its own hi plane: TMA loads the operand tiles UNTOUCHED straight from the user's arrays,
// and four converter warps derive only   lo = rn_tf32(x - trunc_tf32(x))   shared memory -> shared
// memory (same swizzled offset, so the layout is irrelevant to them), fence.proxy.async, and hand
// the stage to the MMA warp.  No pre-pass, no temporary planes, one read of A and B.
// (Dead ends, with ncu numbers in profiles/: splitting in a pre-pass costs 14 % of device time and
// doubles the L2->SM traffic; splitting in registers of LDG-producer warps is instruction-issue
// bound -- 4.3e9 warp instructions, tensor pipe 28 % active -- because every CTA re-splits its tiles.)
//
// PRE-SPLIT mode (fallback for operands TMA cannot address: unaligned base / pitch, two non-unit
// strides): a streaming pre-pass gathers them into compact hi/lo planes which TMA then loads.
//
// Accumulation: the tensor core adds into the fp32 TMEM accumulator with TRUNCATION (measured here:
// error grew linearly with K, 1.7e-2 at K=8192 when one TMEM chain ran over all of K).  So a chain
// never runs longer than kChunk k-blocks (128 K): after each chunk the accumulator is handed to
// the epilogue warps, which add it into fp32 REGISTERS with round-to-nearest, while the MMA warp
// already fills the other TMEM buffer (same promotion idea as Ootomo & Yokota's tensor-core SGEMM).
//
// Kernel anatomy (one CTA per SM, persistent over output tiles, 512 threads = 4 warpgroups whose
// register shares are re-partitioned with setmaxnreg: 40 / 216 / 128 / 128):
//   warp 0      TMA producer : cp.async.bulk.tensor.2d (SWIZZLE_128B for K-major operands,
//                              SWIZZLE_128B_ATOM_32B for MN-major ones) into a 3-stage 64 KB ring
//   warp 1      MMA issuer   : one lane issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8), 12 per
//                              k-block; tcgen05.commit frees ring slots / publishes accumulators
//   warp 2      TMEM allocator
//   warps 4-7   epilogue     : tcgen05.ld -> fp32 register sums (promotion) -> global (C or C +=)
//   warps 8-15  converters   : raw tile -> lo tile (RAW mode only)
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

namespace mdb {

namespace tc {

struct Params {
  int M, N, K;
  int raw;                      // 1: RAW mode (converter warps make the lo tiles), 0: PRE-SPLIT mode
  int a_mn_major, b_mn_major;   // operand major-ness (0: K-major, 1: MN-major)
  float* C;
  int64_t ldc;
  int accumulate;
  int tiles_m, tiles_n, group_m;
  int debug;   // MDB_GEMM_DEBUG: 1 = epilogue stores a sentinel instead of the result
  float rz_gain;   // compensation of the accumulator's truncation bias per MMA instruction (gemm_pair.cuh, epilogue)
  int flags;   // tuning switches (mdb_gemm_tune): bit0 issue hi*hi before waiting for lo (measured 4 % slower),
               // bit1 round lo with cvt.rna (9 % slower than integer rounding), bit2 do not round lo at all
               // (5 % faster, max error +16 %: default)
};

// Shared memory: two rings of operand tiles, every tile is (rows x 128 B) in 1024-B aligned atoms.
//   hi ring  (kHi deep): [A_hi | B_hi] per slot -- the TMA target; in RAW mode these are the raw tiles
//   lo ring  (kLo deep): [A_lo | B_lo] per slot -- written by the converter warps (RAW) or TMA (PRE-SPLIT)
// The lo tiles are produced just in time, so 2 slots suffice; that leaves room for 5 hi slots, i.e.
// 3-4 k-blocks of TMA lookahead (with one shared 3-stage ring the convert step cost a whole stage of
// lookahead and the kernel ran depth-limited: measured 5.6 ms vs 4.3 ms per 8192^3 GEMM).
template <int BN, int kHi, int kLo>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int SLOT_BYTES = A_BYTES + B_BYTES;
  static constexpr int HI_BYTES = kHi * SLOT_BYTES;
  static constexpr int LO_BYTES = kLo * SLOT_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int TOTAL = HI_BYTES + LO_BYTES + BAR_BYTES + 1024;  // + alignment slack
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

// ring position helper: slot index + phase parity
struct Ring {
  int slot = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int depth) {
    if (++slot == depth) { slot = 0; phase ^= 1; }
  }
};

template <int BN, int kHi, int kLo>
__global__ void __launch_bounds__(kThreads, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const Params p) {
  using S = Smem<BN, kHi, kLo>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1024-B aligned
  uint8_t* lo_ring = smem + S::HI_BYTES;
  uint64_t* hi_full = (uint64_t*)(smem + S::HI_BYTES + S::LO_BYTES);   // hi slot landed (TMA complete_tx)
  uint64_t* hi_empty = hi_full + kHi;                                  // hi slot consumed (MMA commit)
  uint64_t* lo_full = hi_empty + kHi;                                  // lo slot written
  uint64_t* lo_empty = lo_full + kLo;                                  // lo slot consumed (MMA commit)
  uint64_t* tmem_full = lo_empty + kLo;          // [2]
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
    for (int s = 0; s < kHi; ++s) { mbar_init(&hi_full[s], 1); mbar_init(&hi_empty[s], 1); }
    for (int s = 0; s < kLo; ++s) {
      mbar_init(&lo_full[s], p.raw ? kConverterThreads : 1);   // RAW: every converter thread arrives
      mbar_init(&lo_empty[s], 1);
    }
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

  // setmaxnreg sits at the top of each role's branch (ptxas applies the new limit to the code it
  // dominates): control warpgroup 40, epilogue 216, the two converter warpgroups keep the launch value (128)
  if (warp == 2 || warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else if (warp == 0) {
    // ===================================== TMA producer =====================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {                 // uniform-datapath issue, see tc_common.cuh
      Ring hi, lo;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(p, t, m_blk, n_blk);
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          const int k0 = kb * BK;
          mbar_wait(&hi_empty[hi.slot], hi.phase ^ 1);
          const uint32_t a_hi = smem_u32(smem + hi.slot * S::SLOT_BYTES), b_hi = a_hi + S::A_BYTES;
          uint64_t* hbar = &hi_full[hi.slot];
          mbar_expect_tx(hbar, S::SLOT_BYTES);
          if (!p.a_mn_major) {           // memory is [M][K], K contiguous: one (32 x BM) box
            tma_load_2d(a_hi, &map_a_hi, hbar, k0, m0);
          } else {                       // memory is [K][M], M contiguous: BM/32 boxes of (32 x 32)
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(a_hi + c * 4096, &map_a_hi, hbar, m0 + 32 * c, k0);
          }
          if (!p.b_mn_major) {           // memory is [N][K], K contiguous
            tma_load_2d(b_hi, &map_b_hi, hbar, k0, n0);
          } else {                       // memory is [K][N], N contiguous
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) tma_load_2d(b_hi + c * 4096, &map_b_hi, hbar, n0 + 32 * c, k0);
          }
          hi.advance(kHi);
          if (!p.raw) {                  // PRE-SPLIT: the lo planes come by TMA as well
            mbar_wait(&lo_empty[lo.slot], lo.phase ^ 1);
            const uint32_t a_lo = smem_u32(lo_ring + lo.slot * S::SLOT_BYTES), b_lo = a_lo + S::A_BYTES;
            uint64_t* lbar = &lo_full[lo.slot];
            mbar_expect_tx(lbar, S::SLOT_BYTES);
            if (!p.a_mn_major) {
              tma_load_2d(a_lo, &map_a_lo, lbar, k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 32; ++c) tma_load_2d(a_lo + c * 4096, &map_a_lo, lbar, m0 + 32 * c, k0);
            }
            if (!p.b_mn_major) {
              tma_load_2d(b_lo, &map_b_lo, lbar, k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 32; ++c) tma_load_2d(b_lo + c * 4096, &map_b_lo, lbar, n0 + 32 * c, k0);
            }
            lo.advance(kLo);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
    Ring hi, lo;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_k; ++kb) {
        const bool chunk_start = (kb % kChunk) == 0;
        const bool chunk_end = ((kb + 1) % kChunk) == 0 || kb == num_k - 1;
        if (chunk_start) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);    // epilogue has drained this accumulator
          tcgen05_fence_after();
        }
        const uint32_t tmem_d = tmem_base + acc * BN;
        mbar_wait(&hi_full[hi.slot], hi.phase);          // TMA bytes of the hi tiles are visible
        tcgen05_fence_after();
        const uint32_t a_hi = smem_u32(smem + hi.slot * S::SLOT_BYTES), b_hi = a_hi + S::A_BYTES;
        const uint32_t a_lo = smem_u32(lo_ring + lo.slot * S::SLOT_BYTES), b_lo = a_lo + S::A_BYTES;
        const bool early = p.flags & 1;
        if (early && elect_one()) {                      // hi*hi needs only the raw tiles: issue it
#pragma unroll                                           // while the converters still work on lo
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_tf32(tmem_d, make_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt),
                      make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt), idesc, !(chunk_start && k == 0));
        }
        __syncwarp();
        mbar_wait(&lo_full[lo.slot], lo.phase);          // lo tiles written (converters / TMA)
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da_hi = make_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t da_lo = make_desc(a_lo + k * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t db_hi = make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
            const uint64_t db_lo = make_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
            umma_tf32(tmem_d, da_lo, db_hi, idesc, early || !(chunk_start && k == 0));
            umma_tf32(tmem_d, da_hi, db_lo, idesc, 1);
            if (!early) umma_tf32(tmem_d, da_hi, db_hi, idesc, 1);
          }
          umma_commit(&lo_empty[lo.slot]);                // both slots are free once these MMAs retire
          umma_commit(&hi_empty[hi.slot]);
          if (chunk_end) umma_commit(&tmem_full[acc]);    // chunk accumulator ready for promotion
        }
        __syncwarp();
        hi.advance(kHi);
        lo.advance(kLo);
        if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ===================================== converters (RAW mode) =============================
    // lo = rn_tf32(x - trunc_tf32(x)) for every element of the two raw tiles of a hi slot, written
    // at the same byte offset of the lo slot (the swizzle is a pure address permutation, so a
    // linear sweep over the tiles is layout-agnostic and bank-conflict free).
    if (p.raw) {
      const int t = threadIdx.x - 8 * 32;            // 0..255
      Ring hi, lo;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&hi_full[hi.slot], hi.phase);
          mbar_wait(&lo_empty[lo.slot], lo.phase ^ 1);
          const uint32_t src = smem_u32(smem + hi.slot * S::SLOT_BYTES);
          const uint32_t dst = smem_u32(lo_ring + lo.slot * S::SLOT_BYTES);
          constexpr int kVecs = S::SLOT_BYTES / 16 / kConverterThreads;   // 8 float4 per thread
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float4 v[kVecs / 2];
#pragma unroll
            for (int j = 0; j < kVecs / 2; ++j)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w)
                           : "r"(src + (t + (half * (kVecs / 2) + j) * kConverterThreads) * 16));
#pragma unroll
            for (int j = 0; j < kVecs / 2; ++j) {
              float e[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float h = __uint_as_float(__float_as_uint(e[i]) & 0xFFFFE000u);   // what the MMA sees
                // round-half-away to tf32 with integer ops (cvt.rna.tf32 runs on the quarter-rate
                // conversion pipe: 8192 of them per k-block cost ~512 cycles per SM, measured as
                // the bottleneck); sign-magnitude bits make +0x1000 then mask exactly that rounding
                if (p.flags & 4) {
                  e[i] = __fsub_rn(e[i], h);          // leave the rounding of lo to the tensor core (truncation)
                } else if (p.flags & 2) {
                  uint32_t lb;
                  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(__fsub_rn(e[i], h)));
                  e[i] = __uint_as_float(lb);
                } else {
                  e[i] = __uint_as_float((__float_as_uint(__fsub_rn(e[i], h)) + 0x1000u) & 0xFFFFE000u);
                }
              }
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (t + (half * (kVecs / 2) + j) * kConverterThreads) * 16),
                           "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3])
                           : "memory");
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to UMMA
          mbar_arrive(&lo_full[lo.slot]);
          hi.advance(kHi);
          lo.advance(kLo);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
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
        const float comp = 1.f + p.rz_gain * (float)(min(kChunk, num_k - ch * kChunk) * 12);
        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32);
          MDB_TMEM_LD32(taddr, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          // chain sums whose low mantissa bits are ALL zero lost nothing (exact inputs, e.g. small integers): keep
          // them exact.  Tested per batch of 32 columns (16 LOP3 + 1 select; a per-element test tripled the ALU work
          // of the promotion and cost the converter-bound fast split 8 %).
          uint32_t low = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) low |= r[j];
          const float cj = (low & 0xFu) ? comp : 1.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] = __fmaf_rn(__uint_as_float(r[j]), cj, sum[c * 32 + j]);
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


#include "gemm_pair.cuh"      // CTA-pair kernel (the default for large problems)

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
int g_gemm_flags = 4;   // default: lo is left unrounded (the tensor core truncates it); see mdb_gemm_tune
                        // bit4 (16): never use the CTA-pair kernel; bit5 (32): always use it when legal
double g_pair_speedup = 1.45;   // per-flop rate of the pair kernel relative to the single-CTA one

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

template <int BN, int kHi, int kLo>
static int launch(const CUtensorMap maps[4], const tc::Params& p) {
  using S = tc::Smem<BN, kHi, kLo>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory budget");
  auto kern = tc::gemm_3xtf32_kernel<BN, kHi, kLo>;
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

typedef void (*PairKernel)(const CUtensorMap, const CUtensorMap, const tc::PairParams);
extern uint64_t g_gemm_path[MDB_GEMM_NPATHS];

// tuning knobs (mdb_gemm_knob): -1 = automatic
int g_knob_raster = -1, g_knob_group = -1, g_knob_hint_a = -1, g_knob_hint_b = -1, g_knob_hint_c = -1;
int g_knob_streamk = -1;          // -1 auto, 0 never, 1 whenever the split is legal
int g_knob_max_clusters = -1;     // cap on co-resident CTA pairs (leave SMs to a concurrent NCCL kernel); -1 = all
int g_knob_split = -1;            // 0 (and -1 = default) = 3xTF32, 1 = TF32 + 2 BF16 cross terms (gemm_pair.cuh "hybrid": faster, ~2.5x the error)
int g_knob_chunk = -1;            // k-blocks per in-TMEM accumulation chain; -1 = default (kChunk)
int g_knob_rz_gain = -1;          // accumulator-truncation compensation per MMA instruction, in units of 1e-10; -1 = default
int g_knob_l2_budget_mb = 32;     // an operand up to this size is walked whole per band of tiles (stays L2-resident)

// Calibrated on hardware (scripts/gemm_split_check.py): the value that minimises the rms error against float64;
// theory: E[0.5 ulp(x) / x] / 2 = 2.15e-8 per instruction.  3xTF32: rms 1.24e-6 -> 0.53e-6, max 6.8e-6 -> 3.9e-6 of
// the result's rms (NumPy/OpenBLAS sgemm: 0.34e-6 / 3.0e-6).
static float rz_gain_for(bool hybrid) {
  if (g_knob_rz_gain >= 0) return 1e-10f * (float)g_knob_rz_gain;
  return hybrid ? 200e-10f : 250e-10f;
}

constexpr int kPairHi = 4, kPairLo = 3;
using PairSmem = tc::Smem<tc::PBN, kPairHi, kPairLo>;
static_assert(PairSmem::TOTAL <= 227 * 1024, "shared memory budget");

// one-time: opt in to the shared-memory size and ask how many CTA pairs can be co-resident
static int pair_max_clusters(int* out) {
  static int max_clusters = 0;
  if (!max_clusters) {
    auto kern = tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, false, false>;
    MDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::TOTAL));
    MDB_CUDA(cudaFuncSetAttribute(tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, true, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::TOTAL));
    MDB_CUDA(cudaFuncSetAttribute(tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::TOTAL));
    MDB_CUDA(cudaFuncSetAttribute(tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, true, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::TOTAL));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g_sm_count / 2 * 2, 1, 1);
    cfg.blockDim = dim3(tc::kPairThreads, 1, 1);
    cfg.dynamicSmemBytes = PairSmem::TOTAL;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = g_sm_count / 2;
    }
    max_clusters = std::min(n, g_sm_count / 2);
  }
  *out = max_clusters;
  return 0;
}

// stream-K workspace: one 256 x 256 fp32 slot + 16 flags per cluster.  Allocated once, outside any
// CUDA-graph capture (a capture cannot cudaMalloc): while unavailable the planner stays data-parallel.
static float* g_sk_partials = nullptr;
static uint32_t* g_sk_flags = nullptr;
static bool sk_workspace_ready() {
  if (g_sk_partials) return true;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(g_stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;
  }
  const size_t clusters = (size_t)g_sm_count / 2;
  void *a = nullptr, *b = nullptr;
  if (cudaMalloc(&a, clusters * tc::kSkSlotFloats * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&b, clusters * 16 * sizeof(uint32_t)) != cudaSuccess ||
      cudaMemsetAsync(b, 0, clusters * 16 * sizeof(uint32_t), g_stream) != cudaSuccess) {
    cudaGetLastError();
    if (a) cudaFree(a);
    if (b) cudaFree(b);
    return false;
  }
  g_sk_partials = (float*)a;
  g_sk_flags = (uint32_t*)b;
  return true;
}

// Work split + tile order + L2 hints of one CTA-pair launch.
static void plan_pair(tc::PairParams& q, int max_clusters, int* grid_clusters, bool* streamk) {
  const int tiles = q.tiles_m * q.tiles_n, C = max_clusters;
  const int num_k = (q.K + tc::BK - 1) / tc::BK;
  const double panel = 256.0 * (double)q.K * 4.0;                // one 256-row operand panel, bytes
  const double a_total = (double)q.M * q.K * 4.0, b_total = (double)q.N * q.K * 4.0;
  const double budget = (double)g_knob_l2_budget_mb * 1048576.0;
  // ---- tile order (measured with ncu, profiles/r02_gemm_dram_traffic.md).  The clusters of a wave march
  // through k together, so inside a wave every operand panel is fetched from DRAM about once; across
  // waves the L2 keeps a panel only when the panels of one wave (74 tiles: (rows + cols) * 256*K*4 B)
  // stay well below its effective capacity -- at K = 4096 they are ~70 MB and nothing survives, with
  // or without eviction hints (evict_last on the re-used operand changed the DRAM bytes by < 5 %,
  // evict_first on the streamed one made them 40 % worse: it breaks the sharing INSIDE the wave).
  //   * smaller operand <= budget (32 MB): walk ALL its panels per band, the other operand streams
  //     past once (fwd1 / dh2 / fwd3 of the C4 step: DRAM bytes 0.99-1.19 x algorithmic);
  //   * otherwise near-square waves (8 x 9.25 tiles), the minimum of (rows + cols).
  int raster, group, hint_a = 0, hint_b = 0;
  (void)panel;
  if (std::min(a_total, b_total) <= budget) {
    if (b_total <= a_total) { raster = 1; group = q.tiles_n; }
    else { raster = 0; group = q.tiles_m; }
  } else {
    raster = b_total < a_total ? 1 : 0;
    group = 8;
  }
  group = std::min(group, raster ? q.tiles_n : q.tiles_m);
  if (g_knob_raster >= 0) raster = g_knob_raster;
  if (g_knob_group > 0) group = std::min(g_knob_group, raster ? q.tiles_n : q.tiles_m);
  q.raster = raster; q.group = std::max(group, 1);
  q.hint_a = g_knob_hint_a >= 0 ? g_knob_hint_a : hint_a;
  q.hint_b = g_knob_hint_b >= 0 ? g_knob_hint_b : hint_b;
  const double c_bytes = (double)q.M * q.N * 4.0;
  (void)c_bytes;
  q.hint_c = g_knob_hint_c >= 0 ? g_knob_hint_c : 0;
  // ---- work split
  q.dp_tiles = tiles; q.sk_clusters = 0; q.sk_share = 0;
  q.sk_partials = nullptr; q.sk_flags = nullptr;
  *grid_clusters = std::min(tiles, C);
  *streamk = false;
  const int w = tiles / C, R = tiles % C;
  if (R == 0 || g_knob_streamk == 0 || num_k < 2 * tc::kChunk) return;
  const double t_dp = w + 1.0;
  int sk_clusters = 0, share = 0;
  double t_sk = t_dp;
  const int s = C / R;
  if (s >= 2) {
    // even split: every remaining tile is cut into s equal k-ranges; the clusters of one range march
    // through k together, so operand sharing in L2 is as good as in a data-parallel wave
    share = ((num_k + s - 1) / s + tc::kChunk - 1) / tc::kChunk * tc::kChunk;
    sk_clusters = (int)(((long long)R * num_k + share - 1) / share);
    t_sk = w + (double)share / num_k + 0.02;                   // + fix-up
  } else if (std::min(a_total + b_total, ((R + q.group - 1) / q.group + q.group) * panel) <= 80.0 * 1048576.0 ||
             g_knob_streamk == 1) {
    // true stream-K of the last wave: its clusters drift apart in k, which costs nothing only while
    // the operand panels those R tiles touch stay in L2 between the drifting readers
    const long long total = (long long)R * num_k;
    share = (int)(((total + C - 1) / C + tc::kChunk - 1) / tc::kChunk * tc::kChunk);
    sk_clusters = (int)((total + share - 1) / share);
    t_sk = w + (double)share / num_k + 0.02;
  }
  if (sk_clusters == 0 || share < 2 * tc::kChunk) return;
  if (g_knob_streamk != 1 && t_sk > 0.96 * t_dp) return;
  if (!sk_workspace_ready()) return;
  q.dp_tiles = w * C; q.sk_clusters = sk_clusters; q.sk_share = share;
  q.sk_partials = g_sk_partials; q.sk_flags = g_sk_flags;
  *grid_clusters = w > 0 ? C : sk_clusters;
  *streamk = true;
}

int g_last_plan[8] = {};   // mdb_gemm_last_plan: clusters, raster, group, dp_tiles, sk_clusters, sk_share, hints, tiles

static int launch_pair(const CUtensorMap& map_a, const CUtensorMap& map_b, tc::PairParams& q) {
  int max_clusters = 0;
  MDB_TRY(pair_max_clusters(&max_clusters));
  if (g_knob_max_clusters > 0) max_clusters = std::min(max_clusters, g_knob_max_clusters);
  int clusters = 0;
  bool streamk = false;
  plan_pair(q, max_clusters, &clusters, &streamk);
  g_last_plan[0] = clusters; g_last_plan[1] = q.raster; g_last_plan[2] = q.group; g_last_plan[3] = q.dp_tiles;
  g_last_plan[4] = q.sk_clusters; g_last_plan[5] = q.sk_share;
  g_last_plan[6] = q.hint_a * 100 + q.hint_b * 10 + q.hint_c; g_last_plan[7] = q.tiles_m * q.tiles_n;
  static const bool timing = getenv("MDB_GEMM_TIMING") != nullptr;
  const bool hybrid = g_knob_split == 1;
  q.rz_gain = rz_gain_for(hybrid);
  q.chunk = g_knob_chunk > 0 ? std::min(g_knob_chunk, 64) : tc::kChunk;
  if (!timing) {
    if (hybrid)
      tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, false, true>
          <<<2 * clusters, tc::kPairThreads, PairSmem::TOTAL, g_stream>>>(map_a, map_b, q);
    else
      tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, false, false>
          <<<2 * clusters, tc::kPairThreads, PairSmem::TOTAL, g_stream>>>(map_a, map_b, q);
    MDB_CHECK_LAUNCH();
    ++g_gemm_path[streamk ? MDB_GEMM_PATH_TC_PAIR_STREAMK : MDB_GEMM_PATH_TC_PAIR];
    return 0;
  }
  // diagnostic mode: per-role stall cycles, averaged over leader / follower CTAs, printed to stderr
  static unsigned long long* dbuf = nullptr;
  if (!dbuf) MDB_CUDA(cudaMalloc(&dbuf, 16 * 8 * 2 * 148));
  MDB_CUDA(cudaMemsetAsync(dbuf, 0, 16 * 8 * 2 * clusters, g_stream));
  q.timing = dbuf;
  if (hybrid)
    tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, true, true>
        <<<2 * clusters, tc::kPairThreads, PairSmem::TOTAL, g_stream>>>(map_a, map_b, q);
  else
    tc::gemm_3xtf32_pair_kernel<kPairHi, kPairLo, true, false>
        <<<2 * clusters, tc::kPairThreads, PairSmem::TOTAL, g_stream>>>(map_a, map_b, q);
  MDB_CHECK_LAUNCH();
  ++g_gemm_path[streamk ? MDB_GEMM_PATH_TC_PAIR_STREAMK : MDB_GEMM_PATH_TC_PAIR];
  std::vector<unsigned long long> h(16 * 2 * clusters);
  MDB_CUDA(cudaMemcpyAsync(h.data(), dbuf, h.size() * 8, cudaMemcpyDeviceToHost, g_stream));
  MDB_CUDA(cudaStreamSynchronize(g_stream));
  static const char* names[14] = {"prod.wait_hi_empty", "prod.total", "mma.wait_lo_full", "mma.wait_tmem_empty", "mma.total",
                                  "-", "conv.wait_hi_full", "conv.wait_lo_empty", "conv.work", "conv.fence+arrive",
                                  "conv.total", "epi.wait_tmem_full", "epi.total", "epi.wait_streamk"};
  const int tiles = q.tiles_m * q.tiles_n;
  const double kblocks = (double)((q.K + tc::BK - 1) / tc::BK) * tiles / clusters;
  fprintf(stderr, "[gemm timing] pair M=%d N=%d K=%d raster=%d group=%d dp_tiles=%d sk=%dx%d  (cycles per k-block, "
                  "leader | follower)\n", q.M, q.N, q.K, q.raster, q.group, q.dp_tiles, q.sk_clusters, q.sk_share);
  for (int i = 0; i < 14; ++i) {
    if (i == 5) continue;
    double sacc[2] = {0, 0};
    for (int c = 0; c < 2 * clusters; ++c) sacc[c & 1] += (double)h[c * 16 + i];
    const double per = (double)clusters * kblocks;
    fprintf(stderr, "  %-22s %9.1f | %9.1f\n", names[i], sacc[0] / per, sacc[1] / per);
  }
  return 0;
}

// Wave efficiency of a persistent grid: useful tile-slots / occupied tile-slots.
static double wave_eff(int64_t tiles, int workers) {
  const int64_t waves = (tiles + workers - 1) / workers;
  return (double)tiles / (double)(waves * workers);
}

// Can TMA read this operand in place?  Needs a unit stride along k (K-major) or along m/n
// (MN-major), a 16-byte aligned base and a row pitch that is a multiple of 16 bytes.
static bool tma_addressable(const float* ptr, int64_t s_mn, int64_t s_k, int mn, int k, bool* mn_major,
                            int64_t* pitch) {
  if (((uintptr_t)ptr & 15) != 0) return false;
  if (s_k == 1 && s_mn >= k) { *mn_major = false; *pitch = s_mn; }
  else if (s_mn == 1 && s_k >= mn) { *mn_major = true; *pitch = s_k; }
  else return false;
  return *pitch % 4 == 0 && *pitch < (int64_t(1) << 36);
}

// epi: optional fused epilogue (bias / relu / mask); only the CTA-pair kernel implements it, so a
// call that carries one returns MDB_ENOTSUP when the pair kernel cannot take the problem
int gemm_tcgen05(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate, const GemmEpilogue* epi) {
  const int64_t M = a->shape[0], K = a->shape[1], N = b->shape[1];
  // small problems are launch-latency bound: the CUDA-core kernel is as fast
  if (M * N * K < (int64_t(1) << 21) || K < 32 || M < 32 || N < 32)
    return set_error(MDB_ENOTSUP, "problem too small for the tensor-core path");
  if (c->strides[1] != 1)
    return set_error(MDB_ENOTSUP, "output must be row-major for the tensor-core path");
  MDB_TRY(load_encode());

  constexpr int BN = 128, kHi = 5, kLo = 2;
  CUtensorMap maps[4];
  tc::Params p;
  Plane pa, pb;
  bool a_mn = false, b_mn = false;
  int64_t a_pitch = 0, b_pitch = 0;
  static const bool no_raw = getenv("MDB_GEMM_PRESPLIT") != nullptr;   // A/B switch for measurements
  const bool raw = !no_raw &&
      tma_addressable((const float*)a->ptr, a->strides[0], a->strides[1], (int)M, (int)K, &a_mn, &a_pitch) &&
      tma_addressable((const float*)b->ptr, b->strides[1], b->strides[0], (int)N, (int)K, &b_mn, &b_pitch);
  if (epi && !(raw && M > tc::BM && N > tc::PBN))
    return set_error(MDB_ENOTSUP, "fused epilogue needs the CTA-pair kernel (M > 128, N > 128, TMA-addressable operands)");
  if (raw) {
    // TMA reads the user's arrays directly; the raw tile doubles as the hi plane
    if (!a_mn) MDB_TRY(make_map(&maps[0], (const float*)a->ptr, (int)K, (int)M, (int)a_pitch, tc::BM, false));
    else MDB_TRY(make_map(&maps[0], (const float*)a->ptr, (int)M, (int)K, (int)a_pitch, 32, true));
    if (!b_mn) MDB_TRY(make_map(&maps[2], (const float*)b->ptr, (int)K, (int)N, (int)b_pitch, BN, false));
    else MDB_TRY(make_map(&maps[2], (const float*)b->ptr, (int)N, (int)K, (int)b_pitch, 32, true));
    maps[1] = maps[0];
    maps[3] = maps[2];
    p.a_mn_major = a_mn; p.b_mn_major = b_mn;
  } else {
    MDB_TRY(split_operand((const float*)a->ptr, (int)M, (int)K, a->strides[0], a->strides[1], &pa));
    MDB_TRY(split_operand((const float*)b->ptr, (int)N, (int)K, b->strides[1], b->strides[0], &pb));
    const int a_box = pa.mn_major ? 32 : tc::BM, b_box = pb.mn_major ? 32 : BN;
    MDB_TRY(make_map(&maps[0], (const float*)pa.hi.ptr, pa.inner, pa.outer, pa.ld, a_box, pa.mn_major));
    MDB_TRY(make_map(&maps[1], (const float*)pa.lo.ptr, pa.inner, pa.outer, pa.ld, a_box, pa.mn_major));
    MDB_TRY(make_map(&maps[2], (const float*)pb.hi.ptr, pb.inner, pb.outer, pb.ld, b_box, pb.mn_major));
    MDB_TRY(make_map(&maps[3], (const float*)pb.lo.ptr, pb.inner, pb.outer, pb.ld, b_box, pb.mn_major));
    p.a_mn_major = pa.mn_major; p.b_mn_major = pb.mn_major;
  }
  if (raw && (!(g_gemm_flags & 16) || epi) && M > tc::BM && N > tc::PBN) {
    // CTA-pair kernel unless wave quantisation of its 256 x 256 tiles eats the gain (measured ratio
    // of the two kernels' per-flop rates: see profiles/)
    const int64_t pm = (M + 255) / 256, pn = (N + 255) / 256;
    const int64_t sm = (M + tc::BM - 1) / tc::BM, sn = (N + BN - 1) / BN;
    const double useful_pair = (double)M * N / ((double)pm * pn * 65536.0);
    const double useful_single = (double)M * N / ((double)sm * sn * 16384.0);
    // (the pair kernel's last wave can be stream-K split, so its wave loss is at most ~1/(2C) tile)
    int maxc = 0;
    MDB_TRY(pair_max_clusters(&maxc));
    const int64_t pt = pm * pn;
    double pair_waves = (double)((pt + maxc - 1) / maxc);
    if (g_knob_streamk != 0 && pt % maxc != 0 && K >= 2048) {
      const int64_t R = pt % maxc, sdiv = maxc / R;
      const double last = sdiv >= 2 ? 1.0 / (double)sdiv : 1.0;
      pair_waves = (double)(pt / maxc) + last + 0.03;
    }
    const double e_pair = (double)pt / (pair_waves * maxc) * useful_pair * g_pair_speedup;
    const double e_single = wave_eff(sm * sn, g_sm_count) * useful_single;
    if (e_pair >= e_single || (g_gemm_flags & 32) || epi) {
      tc::PairParams q = {};
      q.M = (int)M; q.N = (int)N; q.K = (int)K;
      q.a_mn_major = a_mn; q.b_mn_major = b_mn;
      q.C = (float*)c->ptr; q.ldc = c->strides[0];
      q.accumulate = accumulate;
      q.tiles_m = (int)pm; q.tiles_n = (int)pn;
      q.flags = g_gemm_flags;
      q.timing = nullptr;
      if (epi) { q.bias = epi->bias; q.relu = epi->relu; q.mask_src = epi->mask_src; q.ld_mask = epi->ld_mask; }
      return launch_pair(maps[0], maps[2], q);
    }
  }
  p.raw = raw ? 1 : 0;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.C = (float*)c->ptr; p.ldc = c->strides[0];
  p.accumulate = accumulate;
  p.tiles_m = (int)((M + tc::BM - 1) / tc::BM);
  p.tiles_n = (int)((N + BN - 1) / BN);
  p.group_m = 16;
  const char* dbg = getenv("MDB_GEMM_DEBUG");
  p.debug = dbg ? atoi(dbg) : 0;
  p.flags = g_gemm_flags;
  p.rz_gain = rz_gain_for(false);
  MDB_TRY((launch<BN, kHi, kLo>(maps, p)));
  ++g_gemm_path[raw ? MDB_GEMM_PATH_TC_SINGLE : MDB_GEMM_PATH_TC_PRESPLIT];
  return 0;
}

}  // namespace mdb
