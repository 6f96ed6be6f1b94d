// CTA-pair (cta_group::2) 3xTF32 GEMM kernel.  Included by gemm_tcgen05.cu inside namespace mdb::tc
// (shares Params / Smem / Ring / tile_coords with the single-CTA kernel defined there).
#pragma once

// =====================================================================================================
// CTA-PAIR kernel (cta_group::2): one 256 x 256 output tile per pair of SMs, RAW mode only.
//
// Why: at M=128, N=128 a kind::tf32 MMA reads 8 KB of shared memory per 64 tensor-core cycles -- the
// whole 128 B/clk of an SM -- and the TMA writes (32 KB per k-block) and the converter warps (64 KB per
// k-block) compete for the same port: ncu showed the single-CTA kernel at 47-54 % tensor-pipe
// utilisation with the LSU shared-memory share at 43 % (profiles/r01_gemm_raw_ncu.md).  As a pair,
// each SM still stages a 128-row A tile and a 128-row B tile per k-block (same 32 KB raw + 32 KB lo),
// but every MMA is 256 x 256 x 8: each SM's tensor core works 128 cycles on the same 8 KB, so the
// operand traffic per flop halves (MMA 64 B/clk + converters 42 + TMA 21 = ~125 B/clk).
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
// launch value of 128.  setmaxnreg.inc only draws on registers that other warps RELEASED with .dec
// (a first version launched 640 threads at 96 and asked for more than was released: it hung in .inc),
// so the shares balance: released 88*128 + 64*128 = 19456 >= requested 72*256 = 18432.
//
// Where the time goes (MDB_GEMM_TIMING=1 build, 8192^3, cycles per k-block; the MMAs of one k-block
// need 12 x 128 = 1536 tensor-core cycles):   period 1805 = 85 % tensor-pipe utilisation
//   converter: work 1216 + fence/arrive 253 + waits 185      MMA warp: waits for lo_full 612
//   TMA producer: waits for hi_empty 1135 (never the bottleneck)
// The converter is throughput-bound on the shared-memory port, which the three clients share:
// tensor core 768 wavefronts of 128 B per k-block, converter 256 (LDS) + 256 (STS), TMA writes 256
// = 1536 wavefronts per k-block at 1 wavefront/clk -- the same 1536 cycles the MMAs need.  The
// kernel sits at that shared-memory roofline; software-pipelining the converter loads, deeper
// rings (<5,2>, <6,1>) and dropping two thirds of the MMAs all leave the period unchanged.
constexpr int kPairThreads = 512;
constexpr int kPairConvWarps = 4, kPairEpiWarps = 8;
constexpr int PBN = 128;          // B rows staged per CTA; the UMMA N is 2 * PBN

struct PairParams {
  int M, N, K;
  int a_mn_major, b_mn_major;
  float* C;
  int64_t ldc;
  int accumulate;
  int tiles_m, tiles_n, group_m;   // in 256 x 256 pair tiles
  int flags;
  unsigned long long* timing;      // MDB_GEMM_TIMING=1: per-CTA stall-cycle counters (16 per CTA), else null
};

#define MDB_T0() (kTiming ? clock64() : 0ll)
#define MDB_TACC(var, t0) do { if (kTiming) var += clock64() - (t0); } while (0)

// kMc = 1: cluster of 2 (one pair).  kMc = 2: cluster of 4 = two pairs working on horizontally
// adjacent tiles (same A rows): every A tile is fetched from L2 ONCE and TMA-multicast into both pairs
// (each of the two CTAs that need it loads half of it for both), which cuts the L2 -> SM traffic by
// a quarter.  A slot may only be refilled when BOTH pairs are done with it: hi_empty counts one
// commit per pair, multicast to all four CTAs.
template <int kHi, int kLo, bool kTiming, int kMc = 1>
__global__ void __cluster_dims__(2 * kMc, 1, 1) __launch_bounds__(kPairThreads, 1)
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
  const uint32_t crank = cluster_ctarank();            // rank in the cluster (0..2*kMc-1)
  const uint32_t rank = crank & 1;                      // rank inside the CTA pair
  const uint32_t pair = crank >> 1;                     // which pair of the cluster (kMc = 2)
  const uint32_t leader = crank & ~1u;                  // cluster rank of this pair's leader CTA
  const int first_tile = (int)cluster_id_x(), tile_step = (int)num_clusters_x();
  constexpr uint32_t kTmemCols = 512;            // two 256-column accumulator stages
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_k = (p.K + BK - 1) / BK;
  Params tp;                                     // tile_coords() only reads these three
  tp.tiles_m = p.tiles_m; tp.tiles_n = p.tiles_n; tp.group_m = p.group_m;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kHi; ++s) { mbar_init(&hi_full[s], 1); mbar_init(&hi_empty[s], kMc); }
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
    if (lane == 0) {
      Ring hi;
      long long w_empty = 0, t_all = MDB_T0();
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        int m_blk, n_blk;
        tile_coords(tp, t, m_blk, n_blk);
        const int m0 = m_blk * 256 + (int)rank * BM, n0 = (n_blk * kMc + (int)pair) * 256 + (int)rank * PBN;
        for (int kb = 0; kb < num_k; ++kb) {
          const int k0 = (kTiming && (p.flags & 4096)) ? (kb & 7) * BK : kb * BK;   // 4096: diagnostic, L2-resident k range
          const long long tw = MDB_T0();
          mbar_wait(&hi_empty[hi.slot], hi.phase ^ 1);
          MDB_TACC(w_empty, tw);
          const uint32_t a_hi = smem_u32(smem + hi.slot * S::SLOT_BYTES), b_hi = a_hi + S::A_BYTES;
          uint64_t* hbar = &hi_full[hi.slot];
          mbar_expect_tx(hbar, S::SLOT_BYTES);
          if (!p.a_mn_major) {
            if constexpr (kMc == 1) {
              tma_load_2d(a_hi, &map_a, hbar, k0, m0);
            } else {               // this CTA fetches rows [64 pair, +64) of the tile for both pairs (box = 32 x 64)
              tma_load_2d_mc(a_hi + pair * (64 * 128), &map_a, hbar, k0, m0 + 64 * (int)pair,
                             (uint16_t)((1u << rank) | (1u << (rank + 2))));
            }
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) {
              if constexpr (kMc == 1) tma_load_2d(a_hi + c * 4096, &map_a, hbar, m0 + 32 * c, k0);
              else if ((c >> 1) == (int)pair)     // two of the four 32-column boxes each, for both pairs
                tma_load_2d_mc(a_hi + c * 4096, &map_a, hbar, m0 + 32 * c, k0, (uint16_t)((1u << rank) | (1u << (rank + 2))));
            }
          }
          if (!p.b_mn_major) {
            tma_load_2d(b_hi, &map_b, hbar, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < PBN / 32; ++c) tma_load_2d(b_hi + c * 4096, &map_b, hbar, n0 + 32 * c, k0);
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
      const uint16_t pair_mask = (uint16_t)(3u << (2 * pair)), all_mask = (uint16_t)((1u << (2 * kMc)) - 1);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) |
                             ((uint32_t)p.b_mn_major << 16) | ((uint32_t)((2 * PBN) >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      const uint32_t a_lbo = p.a_mn_major ? 4096 : 16, b_lbo = p.b_mn_major ? 4096 : 16;
      const uint32_t a_sbo = p.a_mn_major ? 512 : 1024, b_sbo = p.b_mn_major ? 512 : 1024;
      const uint32_t a_lt = p.a_mn_major ? 1 : 2, b_lt = p.b_mn_major ? 1 : 2;
      const uint32_t a_kstep = p.a_mn_major ? 1024 : UMMA_K * 4, b_kstep = p.b_mn_major ? 1024 : UMMA_K * 4;
      Ring hi, lo;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_lo = 0, w_tmem = 0, t_all = MDB_T0();
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        for (int kb = 0; kb < num_k; ++kb) {
          const bool chunk_start = (kb % kChunk) == 0;
          const bool chunk_end = ((kb + 1) % kChunk) == 0 || kb == num_k - 1;
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
          if (lane == 0) {
            const uint32_t a_hi = smem_u32(smem + hi.slot * S::SLOT_BYTES), b_hi = a_hi + S::A_BYTES;
            const uint32_t a_lo = smem_u32(lo_ring + lo.slot * S::SLOT_BYTES), b_lo = a_lo + S::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da_hi = make_desc(a_hi + k * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t da_lo = make_desc(a_lo + k * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t db_hi = make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
              const uint64_t db_lo = make_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
              if (kTiming && (p.flags & 65536)) {
                // diagnostic (wrong results): A operand from tensor memory -- what would the period be
                // if the tensor core did not read the A tiles from shared memory?
                const uint32_t ta = tmem_base + (acc ^ 1) * 256 + k * 8;
                umma_tf32_pair_ts(tmem_d, ta, db_hi, idesc, 1);
                umma_tf32_pair_ts(tmem_d, ta, db_lo, idesc, 1);
                umma_tf32_pair_ts(tmem_d, ta + 32, db_hi, idesc, 1);
                continue;
              }
              if (kTiming && (p.flags & (128 | 256))) {
                // diagnostic build only (results are wrong): 128 = alternate the two TMEM buffers
                // between consecutive MMAs (no back-to-back dependency on one accumulator),
                // 256 = issue only hi*hi (one MMA per k-step)
                const uint32_t alt = (p.flags & 128) ? tmem_base + (acc ^ 1) * 256 : tmem_d;
                if (!(p.flags & 256)) {
                  umma_tf32_pair(tmem_d, da_lo, db_hi, idesc, 1);
                  umma_tf32_pair(alt, da_hi, db_lo, idesc, 1);
                }
                umma_tf32_pair((k & 1) ? alt : tmem_d, da_hi, db_hi, idesc, 1);
                continue;
              }
              umma_tf32_pair(tmem_d, da_lo, db_hi, idesc, !(chunk_start && k == 0));
              umma_tf32_pair(tmem_d, da_hi, db_lo, idesc, 1);
              umma_tf32_pair(tmem_d, da_hi, db_hi, idesc, 1);
            }
            umma_commit_pair(&lo_empty[lo.slot], pair_mask);
            umma_commit_pair(&hi_empty[hi.slot], all_mask);       // kMc = 2: both pairs must release an A slot
            if (chunk_end) umma_commit_pair(&tmem_full[acc], pair_mask);
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    constexpr int kConv = kPairConvWarps * 32;                     // 128 threads
    const int t = threadIdx.x - (4 + kPairEpiWarps) * 32;
    Ring hi, lo;
    long long w_hi = 0, w_lo = 0, t_work = 0, t_sig = 0, t_all = MDB_T0();
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      for (int kb = 0; kb < num_k; ++kb) {
        const long long ta = MDB_T0();
        mbar_wait(&hi_full[hi.slot], hi.phase);
        MDB_TACC(w_hi, ta);
        const long long tb = MDB_T0();
        mbar_wait(&lo_empty[lo.slot], lo.phase ^ 1);
        MDB_TACC(w_lo, tb);
        const long long tc0 = MDB_T0();
        const uint32_t src = smem_u32(smem + hi.slot * S::SLOT_BYTES);
        const uint32_t dst = smem_u32(lo_ring + lo.slot * S::SLOT_BYTES);
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
        if (!(kTiming && (p.flags & 512))) {                       // 512: diagnostic, no conversion
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
        if (lane == 0) {                                               // tell the leader's MMA warp
          if (kTiming && (p.flags & 16384)) mbar_arrive_cluster_release(&lo_full[lo.slot], leader);   // diagnostic: 28 % slower
          else mbar_arrive_cluster(&lo_full[lo.slot], leader);
        }
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
    const bool vec_ok = (p.ldc % 4 == 0) && (((uintptr_t)p.C & 15) == 0);
    const bool vec8_ok = (p.ldc % 8 == 0) && (((uintptr_t)p.C & 31) == 0) && !(p.flags & 1048576);   // 1048576: A/B, 128-bit stores
    const int num_chunks = (num_k + kChunk - 1) / kChunk;
    long long w_full = 0, t_all = MDB_T0();
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      int m_blk, n_blk;
      tile_coords(tp, t, m_blk, n_blk);
      const int row = m_blk * 256 + (int)rank * BM + q * 32 + lane;
      const int n0 = (n_blk * kMc + (int)pair) * 256 + eh * 128;
      float sum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < num_chunks; ++ch) {
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
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] = __fadd_rn(sum[c * 32 + j], __uint_as_float(r[j]));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], leader);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (row < p.M && !(kTiming && (p.flags & 524288))) {   // 524288: diagnostic, skip the C store
        float* crow = p.C + (int64_t)row * p.ldc;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col0 = n0 + c * 32;
          if (vec8_ok && !p.accumulate && col0 + 32 <= p.N) {
            // 256-bit stores (sm_100): every instruction writes one whole 32-B sector of this thread's
            // row.  With 128-bit stores each sector took two half-writes and the tile store
            // (128 KB per CTA, all CTAs at once) outlasted the two chunks of TMEM lookahead: ~50 us
            // of exposed store time per 128 MB of output.
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(crow + col0 + j),
                           "f"(sum[c * 32 + j]), "f"(sum[c * 32 + j + 1]), "f"(sum[c * 32 + j + 2]),
                           "f"(sum[c * 32 + j + 3]), "f"(sum[c * 32 + j + 4]), "f"(sum[c * 32 + j + 5]),
                           "f"(sum[c * 32 + j + 6]), "f"(sum[c * 32 + j + 7])
                           : "memory");
          } else if (vec_ok && col0 + 32 <= p.N) {
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
    if (kTiming && warp == 4 && lane == 0) {
      p.timing[blockIdx.x * 16 + 11] = w_full; p.timing[blockIdx.x * 16 + 12] = clock64() - t_all;
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

