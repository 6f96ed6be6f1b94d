// CTA-pair (cta_group::2) 3xTF32 GEMM kernel.  Included by gemm_tcgen05.cu inside namespace mdb::tc
// (shares Smem / Ring with the single-CTA kernel defined there).
#pragma once

// =====================================================================================================
// CTA-PAIR kernel (cta_group::2): one 256 x 256 output tile per pair of SMs, RAW mode only.
//
// Why pairs: at M=128, N=128 a kind::tf32 MMA reads 8 KB of shared memory per 64 tensor-core cycles --
// the whole 128 B/clk of an SM -- and the TMA writes (32 KB per k-block) and the converter warps (64 KB
// per k-block) compete for the same port: the single-CTA kernel measured 47-54 % tensor-pipe
// utilisation.  As a pair, each SM still stages a 128-row A tile and a 128-row B tile per k-block
// (32 KB raw + 32 KB lo), but every MMA is 256 x 256 x 8: each SM's tensor core works 128 cycles on
// the same 8 KB, so the operand traffic per flop halves (MMA 64 B/clk + converters 42 + TMA 21).
//
//   CTA r of the pair loads   A rows  [m0 + 128 r, +128)   and   B rows (n) [n0 + 128 r, +128)
//   and owns accumulator rows [m0 + 128 r, +128) x all 256 columns in ITS tensor memory.
//
// Protocol (barriers live at the same offsets in both CTAs):
//   hi_full   local   TMA complete_tx                       -> local converters
//   lo_full   LEADER  one arrive per converter warp of BOTH CTAs -> leader MMA warp;
//                     it also covers "raw tiles landed" for both CTAs (converters waited on hi_full)
//   hi_empty / lo_empty / tmem_full   both   multicast tcgen05.commit from the leader
//   tmem_empty LEADER one arrive per epilogue warp of both CTAs
// Remote arrives use the plain mbarrier.arrive.shared::cluster form (tc_common.cuh explains why the
// .release.cluster form is ~1000 cycles slower and what makes the plain one sufficient here).
// 512 threads: control warpgroup (TMA, MMA, TMEM alloc), 2 epilogue warpgroups (128 columns each,
// 128 fp32 promotion registers per thread), 1 converter warpgroup; setmaxnreg 40 / 200 / 64 from a
// launch value of 128.  setmaxnreg.inc only draws on registers that other warps RELEASED with .dec,
// so the shares balance: released 88*128 + 64*128 = 19456 >= requested 72*256 = 18432.
//
// Work distribution (round 2).  A cluster's work is a list of SEGMENTS (tile, kb0, kb1):
//   * data-parallel part: tiles [0, dp_tiles) whole, round-robin over the clusters (dp_tiles is a
//     multiple of the cluster count: full waves only);
//   * stream-K part: the k-blocks of the remaining tiles, linearised tile-major, are cut into
//     `sk_share`-sized contiguous ranges, one per cluster (c < sk_clusters).  A range may end inside a
//     tile and the next cluster continues it.  The cluster whose segment STARTS a tile (kb0 == 0)
//     owns it: clusters that continue the tile (kb0 > 0; always their FIRST segment, so it is
//     finished early) deposit their fp32 register sums in a per-cluster slot of a global workspace
//     and raise a flag per epilogue warp; the owner -- whose partial segment is its LAST one -- adds
//     the slots in cluster order (deterministic) and stores.  A wait only ever targets the first
//     segment of a higher-numbered cluster, which never waits itself, so the scheme cannot deadlock
//     even when not all clusters are co-resident.
//   This removes the wave-quantisation loss of the last, partially filled wave (dW2 of the C4 step:
//   256 tiles on 74 clusters = 3.46 waves -> 3 + 34 tiles split two ways = 3.5 instead of 4).
// Tile order (raster): groups of `group` tile-rows walked column by column (raster 0) or groups of
// `group` tile-columns walked row by row (raster 1), so that the clusters of a wave share few
// distinct operand panels; TMA loads and C stores can carry L2 eviction hints (off by default: measured
// useless to harmful, profiles/r02_gemm_dram_traffic.md).
// Fused epilogue (linear_relu): C = relu?( A@B + bias? ) * (mask_src > 0)?
// Operand split (template kHybrid): false = 3xTF32, the default (converters write fp32 lo tiles, 12 MMAs per k-block);
// true = the opt-in "fast" split (converters write K-major BF16 hb / lb tiles, 4 TF32 + 4 BF16 MMAs), see below.
// Promotion: chains of p.chunk k-blocks in tensor memory, then sum = fma(chain, 1 + rz_gain * n, sum) in registers
// (compensates the accumulator's round-toward-zero bias; exact chain sums are passed through unchanged).
constexpr int kPairThreads = 512;
constexpr int kPairConvWarps = 4, kPairEpiWarps = 8;
constexpr int PBN = 128;          // B rows staged per CTA; the UMMA N is 2 * PBN
constexpr int kSkSlotFloats = 2 * kPairEpiWarps * 128 * 32;   // one cluster's 256 x 256 fp32 partial tile

struct PairParams {
  int M, N, K;
  int a_mn_major, b_mn_major;
  float* C;
  int64_t ldc;
  int accumulate;
  int tiles_m, tiles_n;            // in 256 x 256 pair tiles
  int raster, group;               // tile order, see tile_coords_pair
  int dp_tiles;                    // tiles [0, dp_tiles): whole tiles, round-robin
  int sk_clusters, sk_share;       // stream-K part: clusters [0, sk_clusters) take sk_share k-blocks each
  float* sk_partials;              // [clusters][2 CTAs][8 warps][128][32]
  uint32_t* sk_flags;              // [clusters][2 CTAs][8 warps]
  const float* bias;               // epilogue: + bias[col] (nullable)
  int relu;                        // epilogue: max(v, 0) with where(v > 0, v, 0) semantics
  const float* mask_src;           // epilogue: * (mask_src[row, col] > 0) (nullable)
  int64_t ld_mask;
  int hint_a, hint_b, hint_c;      // L2 eviction hints: 0 none, 1 evict_first, 2 evict_last
  int flags;
  float rz_gain;                   // per-MMA-instruction compensation of the accumulator's round-toward-zero bias (see epilogue)
  int chunk;                       // k-blocks per in-TMEM accumulation chain before the promotion to fp32 registers
  unsigned long long* timing;      // MDB_GEMM_TIMING=1: per-CTA stall-cycle counters (16 per CTA), else null
};

#define MDB_T0() (kTiming ? clock64() : 0ll)
#define MDB_TACC(var, t0) do { if (kTiming) var += clock64() - (t0); } while (0)

struct Segment { int tile, kb0, kb1; };

// Every role of the kernel walks the same segment list of its cluster.
struct WorkIter {
  int next_dp, step, dp_tiles, num_k;
  long long lin, lin_end;
  __device__ __forceinline__ WorkIter(const PairParams& p, int cluster, int nclusters, int num_k_) {
    next_dp = cluster; step = nclusters; dp_tiles = p.dp_tiles; num_k = num_k_;
    lin = lin_end = 0;
    if (cluster < p.sk_clusters) {
      const long long total = (long long)(p.tiles_m * p.tiles_n - p.dp_tiles) * num_k;
      lin = (long long)cluster * p.sk_share;
      lin_end = lin + p.sk_share < total ? lin + p.sk_share : total;
    }
  }
  __device__ __forceinline__ bool next(Segment& s) {
    if (next_dp < dp_tiles) {
      s.tile = next_dp; s.kb0 = 0; s.kb1 = num_k;
      next_dp += step;
      return true;
    }
    if (lin < lin_end) {
      const int t = (int)(lin / num_k);
      s.kb0 = (int)(lin - (long long)t * num_k);
      const long long left = lin_end - lin;
      s.kb1 = (long long)s.kb0 + left < num_k ? s.kb0 + (int)left : num_k;
      s.tile = dp_tiles + t;
      lin += s.kb1 - s.kb0;
      return true;
    }
    return false;
  }
};

__device__ __forceinline__ void tile_coords_pair(const PairParams& p, int t, int& m_blk, int& n_blk) {
  if (p.raster == 0) {        // groups of `group` tile-rows, walked column by column (m fastest)
    const int per_group = p.group * p.tiles_n;
    const int g = t / per_group, r = t - g * per_group;
    const int rows = min(p.group, p.tiles_m - g * p.group);
    m_blk = g * p.group + (r % rows);
    n_blk = r / rows;
  } else {                    // groups of `group` tile-columns, walked row by row (n fastest)
    const int per_group = p.group * p.tiles_m;
    const int g = t / per_group, r = t - g * per_group;
    const int cols = min(p.group, p.tiles_n - g * p.group);
    n_blk = g * p.group + (r % cols);
    m_blk = r / cols;
  }
}

__device__ __forceinline__ uint64_t l2_policy(int hint) {
  uint64_t pol = 0;
  if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- TF32 + 2 x BF16 split ("hybrid", kHybrid) ---------------------------------------------------------
// a*b = a_t*b_t + a_l*b + a*b_l - a_l*b_l, a_t = what the TF32 datapath sees (top 19 bits), a_l = a - a_t
// (exact, < 2^-10 |a|).  The big term runs as ONE kind::tf32 MMA on the raw fp32 tiles; the two cross
// terms only need ~8 significant bits per factor to stay below 2^-18 of the product, so they run as
// kind::f16 MMAs on BF16 tiles (K = 16 per instruction: half the instructions and half the shared-memory
// bytes of a TF32 cross term).  Per k-block: 4 + 2 + 2 = 8 tensor-core instructions instead of 12.
// The converters write, per operand, hb = bf16(x) and lb = bf16(x - x_t) as K-MAJOR 128 x 32 BF16 tiles
// (64-byte rows, SWIZZLE_64B: 16-B chunk ^= (row >> 1) & 3, 8-row / 512-B atoms) whatever the layout of
// the raw tile -- an MN-major raw tile is transposed in registers on the way.
__device__ __forceinline__ void lds128(float* v, uint32_t addr) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t* r) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// (x0, x1) = two consecutive k  ->  bf16x2 of the values and of their TF32 remainders (x0 in the low half)
__device__ __forceinline__ void split_pack(float x0, float x1, uint32_t& hb, uint32_t& lb) {
  const float l0 = __fsub_rn(x0, __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u));
  const float l1 = __fsub_rn(x1, __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u));
  asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(hb) : "f"(x1), "f"(x0));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lb) : "f"(l1), "f"(l0));
}
// K-major raw tile (128 rows x 128 B, SWIZZLE_128B).  Converter warp cw, iteration it: rows (it*4+cw)*8 + (lane&7),
// k range 8j .. 8j+7 (j = lane >> 3) = logical 16-B chunks 2j, 2j+1.  A quarter-warp reads / writes 8 distinct
// 16-B bank groups in both directions.
template <int kHalf>
__device__ __forceinline__ void cv_load_k(float (&v)[32], uint32_t src, int cw, int lane) {
  const int r7 = lane & 7, j = lane >> 3;
  const uint32_t p0 = (uint32_t)((2 * j) ^ r7) << 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int it = kHalf * 2 + i;
    const uint32_t row = src + (uint32_t)(((it * 4 + cw) * 8 + r7) * 128);
    lds128(&v[it * 8], row + p0);
    lds128(&v[it * 8 + 4], row + (p0 ^ 16u));
  }
}
template <int kHalf>
__device__ __forceinline__ void cv_proc_k(const float (&v)[32], uint32_t hb, uint32_t lb, int cw, int lane) {
  const int r7 = lane & 7, j = lane >> 3;
  const uint32_t o = (uint32_t)(r7 * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int it = kHalf * 2 + i;
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_pack(v[it * 8 + 2 * e], v[it * 8 + 2 * e + 1], h[e], l[e]);
    const uint32_t off = (uint32_t)((it * 4 + cw) * 512) + o;
    sts128(hb + off, h);
    sts128(lb + off, l);
  }
}
// MN-major raw tile (4 chunks of [32 k][32 m] fp32, 128-B rows, SWIZZLE_128B_ATOM_32B: 32-B unit ^= k & 3).
// Thread: chunk cw, 4 rows m = 8u + 4h + {0..3} (one float4 per k), 8 consecutive k (kg = u ^ q, q = lane >> 3, so
// that the 16-B stores of a quarter-warp fall into distinct bank groups; lanes with h = 1 store their row pairs
// in swapped order for the same reason).
template <int kHalf>
__device__ __forceinline__ void cv_load_mn(float (&v)[32], uint32_t src, int cw, int lane) {
  const int q = lane >> 3, u = (lane >> 1) & 3, h = lane & 1, kg = u ^ q;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = kHalf * 4 + i;
    lds128(&v[s * 4], src + (uint32_t)(cw * 4096 + (kg * 8 + s) * 128 + (((u ^ (s & 3)) << 5) | (h << 4))));
  }
}
template <int kHalf>
__device__ __forceinline__ void cv_proc_mn(const float (&v)[32], uint32_t hb, uint32_t lb, int cw, int lane) {
  const int q = lane >> 3, u = (lane >> 1) & 3, h = lane & 1, kg = u ^ q;
  constexpr int i0 = kHalf * 2;
  uint32_t h0[4], l0[4], h1[4], l1[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    split_pack(v[(2 * e) * 4 + i0], v[(2 * e + 1) * 4 + i0], h0[e], l0[e]);
    split_pack(v[(2 * e) * 4 + i0 + 1], v[(2 * e + 1) * 4 + i0 + 1], h1[e], l1[e]);
  }
  uint32_t fh[4], fl[4], sh[4], sl[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    fh[e] = h ? h1[e] : h0[e]; fl[e] = h ? l1[e] : l0[e];
    sh[e] = h ? h0[e] : h1[e]; sl[e] = h ? l0[e] : l1[e];
  }
  // row within the tile = cw*32 + u*8 + h*4 + i'  ->  atom (cw*4 + u), row-in-atom h*4 + i'
  const uint32_t base = (uint32_t)((cw * 4 + u) * 512 + ((kg ^ (2 * h + kHalf)) << 4));
  const uint32_t first = base + (uint32_t)((h * 4 + i0 + h) * 64), second = base + (uint32_t)((h * 4 + i0 + 1 - h) * 64);
  sts128(hb + first, fh);
  sts128(lb + first, fl);
  sts128(hb + second, sh);
  sts128(lb + second, sl);
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int kHi, int kLo, bool kTiming, bool kHybrid>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_3xtf32_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const PairParams p) {
  using S = Smem<PBN, kHi, kLo>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* lo_ring = smem + S::HI_BYTES;
  uint64_t* hi_full = (uint64_t*)(smem + S::HI_BYTES + S::LO_BYTES);
  uint64_t* hi_empty = hi_full + kHi;
  uint64_t* lo_full = hi_empty + kHi;
  uint64_t* lo_empty = lo_full + kLo;
  uint64_t* tmem_full = lo_empty + kLo;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();              // rank inside the CTA pair; 0 = leader
  const int cluster = (int)cluster_id_x(), nclusters = (int)num_clusters_x();
  constexpr uint32_t kTmemCols = 512;            // two 256-column accumulator stages
  const int num_k = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kHi; ++s) { mbar_init(&hi_full[s], 1); mbar_init(&hi_empty[s], 1); }
    for (int s = 0; s < kLo; ++s) { mbar_init(&lo_full[s], 2 * kPairConvWarps); mbar_init(&lo_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * kPairEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();                            // both CTAs' barriers are initialised, TMEM allocated
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2 || warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      const uint64_t pol_a = l2_policy(p.hint_a), pol_b = l2_policy(p.hint_b);
      Ring hi;
      long long w_empty = 0, t_all = MDB_T0();
      Segment s;
      for (WorkIter w(p, cluster, nclusters, num_k); w.next(s);) {
        int m_blk, n_blk;
        tile_coords_pair(p, s.tile, m_blk, n_blk);
        const int m0 = m_blk * 256 + (int)rank * BM, n0 = n_blk * 256 + (int)rank * PBN;
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          const int k0 = kb * BK;
          const long long tw = MDB_T0();
          mbar_wait(&hi_empty[hi.slot], hi.phase ^ 1);
          MDB_TACC(w_empty, tw);
          const uint32_t a_hi = smem_u32(smem + hi.slot * S::SLOT_BYTES), b_hi = a_hi + S::A_BYTES;
          uint64_t* hbar = &hi_full[hi.slot];
          mbar_expect_tx(hbar, S::SLOT_BYTES);
          if (!p.a_mn_major) {
            tma_load_2d_hint(a_hi, &map_a, hbar, k0, m0, pol_a);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d_hint(a_hi + c * 4096, &map_a, hbar, m0 + 32 * c, k0, pol_a);
          }
          if (!p.b_mn_major) {
            tma_load_2d_hint(b_hi, &map_b, hbar, k0, n0, pol_b);
          } else {
#pragma unroll
            for (int c = 0; c < PBN / 32; ++c) tma_load_2d_hint(b_hi + c * 4096, &map_b, hbar, n0 + 32 * c, k0, pol_b);
          }
          hi.advance(kHi);
        }
      }
      if (kTiming) { p.timing[blockIdx.x * 16 + 0] = w_empty; p.timing[blockIdx.x * 16 + 1] = clock64() - t_all; }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) ======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) |
                             ((uint32_t)p.b_mn_major << 16) | ((uint32_t)((2 * PBN) >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      // kind::f16, BF16 x BF16 -> FP32, both operands K-major, M = 256, N = 256
      const uint32_t idesc_bf = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * PBN) >> 3) << 17) |
                                ((uint32_t)(256 >> 4) << 24);
      const uint32_t a_lbo = p.a_mn_major ? 4096 : 16, b_lbo = p.b_mn_major ? 4096 : 16;
      const uint32_t a_sbo = p.a_mn_major ? 512 : 1024, b_sbo = p.b_mn_major ? 512 : 1024;
      const uint32_t a_lt = p.a_mn_major ? 1 : 2, b_lt = p.b_mn_major ? 1 : 2;
      const uint32_t a_kstep = p.a_mn_major ? 1024 : UMMA_K * 4, b_kstep = p.b_mn_major ? 1024 : UMMA_K * 4;
      Ring hi, lo;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_lo = 0, w_tmem = 0, t_all = MDB_T0();
      Segment s;
      for (WorkIter w(p, cluster, nclusters, num_k); w.next(s);) {
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          const bool chunk_start = ((kb - s.kb0) % p.chunk) == 0;
          const bool chunk_end = ((kb - s.kb0 + 1) % p.chunk) == 0 || kb == s.kb1 - 1;
          if (chunk_start) {
            const long long tw = MDB_T0();
            mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);   // both CTAs drained this accumulator
            MDB_TACC(w_tmem, tw);
            tcgen05_fence_after();
          }
          const uint32_t tmem_d = tmem_base + acc * 256;
          const long long tw2 = MDB_T0();
          mbar_wait_cluster(&lo_full[lo.slot], lo.phase);    // raw + lo tiles ready in both CTAs
          MDB_TACC(w_lo, tw2);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_hi = smem_u32(smem + hi.slot * S::SLOT_BYTES), b_hi = a_hi + S::A_BYTES;
            const uint32_t a_lo = smem_u32(lo_ring + lo.slot * S::SLOT_BYTES), b_lo = a_lo + S::A_BYTES;
            if (kHybrid) {
              // raw x raw in TF32, then the two cross terms on the converters' K-major BF16 tiles:
              // lo slot = [A hb | A lb | B hb | B lb], 8 KB each
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_tf32_pair(tmem_d, make_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt),
                               make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt), idesc, !(chunk_start && k == 0));
              const uint32_t a_hb = a_lo, a_lb = a_lo + 8192, b_hb = a_lo + 16384, b_lb = a_lo + 24576;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_bf16_pair(tmem_d, make_desc(a_lb + k * 32, 16, 512, 4), make_desc(b_hb + k * 32, 16, 512, 4), idesc_bf, 1);
                umma_bf16_pair(tmem_d, make_desc(a_hb + k * 32, 16, 512, 4), make_desc(b_lb + k * 32, 16, 512, 4), idesc_bf, 1);
              }
            } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da_hi = make_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t da_lo = make_desc(a_lo + k * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t db_hi = make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
              const uint64_t db_lo = make_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
              umma_tf32_pair(tmem_d, da_lo, db_hi, idesc, !(chunk_start && k == 0));
              umma_tf32_pair(tmem_d, da_hi, db_lo, idesc, 1);
              umma_tf32_pair(tmem_d, da_hi, db_hi, idesc, 1);
            }
            }
            umma_commit_pair(&lo_empty[lo.slot]);
            umma_commit_pair(&hi_empty[hi.slot]);
            if (chunk_end) umma_commit_pair(&tmem_full[acc]);
          }
          __syncwarp();
          hi.advance(kHi);
          lo.advance(kLo);
          if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (kTiming && lane == 0) {
        p.timing[blockIdx.x * 16 + 2] = w_lo; p.timing[blockIdx.x * 16 + 3] = w_tmem;
        p.timing[blockIdx.x * 16 + 4] = clock64() - t_all;
      }
    }
  } else if (warp >= 4 + kPairEpiWarps) {
    // ===================================== converters (both CTAs) ============================
    // hybrid: 72 (two 32-register tile fragments in flight); the shares still balance exactly:
    // released 4*32*88 + 128*56 = 18432 = requested 256*72
    if (kHybrid) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    constexpr int kConv = kPairConvWarps * 32;                     // 128 threads
    const int t = threadIdx.x - (4 + kPairEpiWarps) * 32;
    Ring hi, lo;
    long long w_hi = 0, w_lo = 0, t_work = 0, t_sig = 0, t_all = MDB_T0();
    Segment s;
    for (WorkIter w(p, cluster, nclusters, num_k); w.next(s);) {
      for (int kb = s.kb0; kb < s.kb1; ++kb) {
        const long long ta = MDB_T0();
        mbar_wait(&hi_full[hi.slot], hi.phase);
        MDB_TACC(w_hi, ta);
        const long long tb = MDB_T0();
        mbar_wait(&lo_empty[lo.slot], lo.phase ^ 1);
        MDB_TACC(w_lo, tb);
        const long long tc0 = MDB_T0();
        const uint32_t src = smem_u32(smem + hi.slot * S::SLOT_BYTES);
        const uint32_t dst = smem_u32(lo_ring + lo.slot * S::SLOT_BYTES);
        if (kHybrid) {
          const int cw = t >> 5;
          const uint32_t src_b = src + S::A_BYTES;
          const uint32_t a_hb = dst, a_lb = dst + 8192, b_hb = dst + 16384, b_lb = dst + 24576;
          float va[32], vb[32];
          // loads of the B tile are issued between the two halves of the A tile's conversion
          if (p.a_mn_major) { cv_load_mn<0>(va, src, cw, lane); cv_load_mn<1>(va, src, cw, lane); }
          else { cv_load_k<0>(va, src, cw, lane); cv_load_k<1>(va, src, cw, lane); }
          if (p.a_mn_major) cv_proc_mn<0>(va, a_hb, a_lb, cw, lane); else cv_proc_k<0>(va, a_hb, a_lb, cw, lane);
          if (p.b_mn_major) cv_load_mn<0>(vb, src_b, cw, lane); else cv_load_k<0>(vb, src_b, cw, lane);
          if (p.a_mn_major) cv_proc_mn<1>(va, a_hb, a_lb, cw, lane); else cv_proc_k<1>(va, a_hb, a_lb, cw, lane);
          if (p.b_mn_major) cv_load_mn<1>(vb, src_b, cw, lane); else cv_load_k<1>(vb, src_b, cw, lane);
          if (p.b_mn_major) { cv_proc_mn<0>(vb, b_hb, b_lb, cw, lane); cv_proc_mn<1>(vb, b_hb, b_lb, cw, lane); }
          else { cv_proc_k<0>(vb, b_hb, b_lb, cw, lane); cv_proc_k<1>(vb, b_hb, b_lb, cw, lane); }
        } else {
        constexpr int kVecs = S::SLOT_BYTES / 16 / kConv;          // 16 float4 per thread
        constexpr int kBatch = 4, kBatches = kVecs / kBatch;
        // software pipeline: the loads of batch b+1 are in flight while batch b is converted and stored
        // (the LDS latency under tensor-core shared-memory traffic is several hundred cycles)
        float4 v[2][kBatch];
        auto load_batch = [&](int bt, float4 (&dstv)[kBatch]) {
#pragma unroll
          for (int j = 0; j < kBatch; ++j)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(dstv[j].x), "=f"(dstv[j].y), "=f"(dstv[j].z), "=f"(dstv[j].w)
                         : "r"(src + (t + (bt * kBatch + j) * kConv) * 16));
        };
        load_batch(0, v[0]);
#pragma unroll
        for (int bt = 0; bt < kBatches; ++bt) {
          if (bt + 1 < kBatches) load_batch(bt + 1, v[(bt + 1) & 1]);
#pragma unroll
          for (int j = 0; j < kBatch; ++j) {
            const float4 x = v[bt & 1][j];
            float e[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float h = __uint_as_float(__float_as_uint(e[i]) & 0xFFFFE000u);   // what the MMA sees
              if (p.flags & 4) e[i] = __fsub_rn(e[i], h);
              else e[i] = __uint_as_float((__float_as_uint(__fsub_rn(e[i], h)) + 0x1000u) & 0xFFFFE000u);
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (t + (bt * kBatch + j) * kConv) * 16),
                         "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3])
                         : "memory");
          }
        }
        }
        MDB_TACC(t_work, tc0);
        const long long td = MDB_T0();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to UMMA
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&lo_full[lo.slot], 0);       // tell the leader's MMA warp
        MDB_TACC(t_sig, td);
        hi.advance(kHi);
        lo.advance(kLo);
      }
    }
    if (kTiming && t == 0) {
      unsigned long long* d = p.timing + blockIdx.x * 16;
      d[6] = w_hi; d[7] = w_lo; d[8] = t_work; d[9] = t_sig; d[10] = clock64() - t_all;
    }
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs) ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int q = warp & 3;                               // TMEM lane quarter this warp may touch
    const int eh = (warp - 4) >> 2;                       // which 128-column half of the accumulator
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (p.ldc % 4 == 0) && (((uintptr_t)p.C & 15) == 0) &&
                        (!p.bias || ((uintptr_t)p.bias & 15) == 0) &&
                        (!p.mask_src || (p.ld_mask % 4 == 0 && ((uintptr_t)p.mask_src & 15) == 0));
    const bool vec8_ok = vec_ok && (p.ldc % 8 == 0) && (((uintptr_t)p.C & 31) == 0);
    const uint64_t pol_c = l2_policy(p.hint_c);
    const int slot_warp = (int)rank * kPairEpiWarps + (warp - 4);          // 0..15 inside the cluster
    long long w_full = 0, w_sk = 0, t_all = MDB_T0();
    Segment s;
    for (WorkIter w(p, cluster, nclusters, num_k); w.next(s);) {
      int m_blk, n_blk;
      tile_coords_pair(p, s.tile, m_blk, n_blk);
      const int row = m_blk * 256 + (int)rank * BM + q * 32 + lane;
      const int n0 = n_blk * 256 + eh * 128;
      float sum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = 0.f;
      const int num_chunks = (s.kb1 - s.kb0 + p.chunk - 1) / p.chunk;
      for (int ch = 0; ch < num_chunks; ++ch) {
        // The tensor core TRUNCATES the fp32 accumulator after every instruction: a chain of n instructions loses
        // ~0.5 ulp per step, always toward zero, and the running sum grows towards its final value r, so the
        // expected loss is proportional to r itself: E[loss | r] ~ rz_gain * n * r.  The promotion adds it back
        // with the multiply of an FMA; what remains is the zero-mean part of the error.  A batch of chain sums whose
        // low four mantissa bits are all zero is taken as exact (nothing was shifted out: products of small integers,
        // powers of two ...) and promoted unchanged, so exactly representable results stay exact.
        const int kbs = min(p.chunk, s.kb1 - s.kb0 - ch * p.chunk);
        const float comp = 1.f + p.rz_gain * (float)(kbs * (kHybrid ? 8 : 12));
        const long long tw = MDB_T0();
        mbar_wait(&tmem_full[acc], acc_phase);
        MDB_TACC(w_full, tw);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + eh * 128 + c * 32);
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
        if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // ---- stream-K hand-over ------------------------------------------------------------------
      if (s.kb0 > 0) {
        // this cluster continued a tile another cluster owns: deposit the register sums, raise the flag
        float* part = p.sk_partials + ((size_t)cluster * (2 * kPairEpiWarps) + slot_warp) * 4096;
#pragma unroll
        for (int j = 0; j < 128; ++j) __stcg(part + j * 32 + lane, sum[j]);
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_gpu(p.sk_flags + cluster * (2 * kPairEpiWarps) + slot_warp, 1u);
        continue;
      }
      if (s.kb1 < num_k) {
        // owner of a tile whose k-range continues in the following clusters: add their sums in order
        const long long tile_end = (long long)(s.tile - p.dp_tiles + 1) * num_k;
        const long long tw = MDB_T0();
        for (int c2 = cluster + 1; c2 < p.sk_clusters && (long long)c2 * p.sk_share < tile_end; ++c2) {
          uint32_t* flag = p.sk_flags + c2 * (2 * kPairEpiWarps) + slot_warp;
          if (lane == 0) {
            long long t0 = 0;
            for (uint32_t spins = 0; ld_acquire_gpu(flag) == 0u; ++spins) {
              if (spins == 64) t0 = clock64();
              if (spins > 64 && (spins & 1023) == 0 && clock64() - t0 > 4000000000ll) __trap();
            }
          }
          __syncwarp();
          const float* part = p.sk_partials + ((size_t)c2 * (2 * kPairEpiWarps) + slot_warp) * 4096;
#pragma unroll
          for (int j = 0; j < 128; ++j) sum[j] = __fadd_rn(sum[j], __ldcg(part + j * 32 + lane));
          __syncwarp();
          if (lane == 0) *(volatile uint32_t*)flag = 0u;      // slot is free for the next launch
        }
        MDB_TACC(w_sk, tw);
      }
      // ---- store (with the fused epilogue) -----------------------------------------------------
      if (row < p.M) {
        float* crow = p.C + (int64_t)row * p.ldc;
        const float* mrow = p.mask_src ? p.mask_src + (int64_t)row * p.ld_mask : nullptr;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col0 = n0 + c * 32;
          if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = sum[c * 32 + j + i];
              if (p.bias) {
                const float4 b0 = __ldg((const float4*)(p.bias + col0 + j)), b1 = __ldg((const float4*)(p.bias + col0 + j + 4));
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __fadd_rn(o[i], b[i]);
              }
              if (p.relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = o[i] > 0.f ? o[i] : 0.f;
              }
              if (mrow) {
                const float4 m0 = __ldg((const float4*)(mrow + col0 + j)), m1 = __ldg((const float4*)(mrow + col0 + j + 4));
                const float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __fmul_rn(o[i], m[i] > 0.f ? 1.f : 0.f);
              }
              if (p.accumulate) {
                const float4 c0 = *(const float4*)(crow + col0 + j), c1 = *(const float4*)(crow + col0 + j + 4);
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __fadd_rn(cc[i], o[i]);
              }
              if (vec8_ok) {
                // 256-bit stores (sm_100): every instruction writes one whole 32-B sector of this thread's
                // row; with 128-bit stores each sector took two half-writes and the tile store outlasted
                // the two chunks of TMEM lookahead
                asm volatile("st.global.L2::cache_hint.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(crow + col0 + j),
                             "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]),
                             "l"(pol_c)
                             : "memory");
              } else {
                *(float4*)(crow + col0 + j) = make_float4(o[0], o[1], o[2], o[3]);
                *(float4*)(crow + col0 + j + 4) = make_float4(o[4], o[5], o[6], o[7]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) {
                float v = sum[c * 32 + j];
                if (p.bias) v = __fadd_rn(v, p.bias[col0 + j]);
                if (p.relu) v = v > 0.f ? v : 0.f;
                if (mrow) v = __fmul_rn(v, mrow[col0 + j] > 0.f ? 1.f : 0.f);
                if (p.accumulate) v = __fadd_rn(crow[col0 + j], v);
                crow[col0 + j] = v;
              }
          }
        }
      }
    }
    if (kTiming && warp == 4 && lane == 0) {
      p.timing[blockIdx.x * 16 + 11] = w_full; p.timing[blockIdx.x * 16 + 12] = clock64() - t_all;
      p.timing[blockIdx.x * 16 + 13] = w_sk;
    }
  }

  // ------------------------------------------ teardown ---------------------------------------
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();        // no CTA leaves (or frees TMEM) while its peer may still signal / read it
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}
