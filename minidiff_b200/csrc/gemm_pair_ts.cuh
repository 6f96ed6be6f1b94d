// Experimental CTA-pair kernel with the A operand in tensor memory.  Included by gemm_tcgen05.cu
// inside namespace mdb::tc, after gemm_pair.cuh (shares PairParams and the timing macros).
#pragma once

// =====================================================================================================
// CTA-PAIR kernel with the A operand in TENSOR MEMORY (experimental, mdb_gemm_tune bit 17).
//
// The pair kernel above sits on the shared-memory port (tensor core 768 + converters 512 + TMA 256
// wavefronts per k-block = the 1536 cycles its MMAs need).  tcgen05.mma can take A from tensor memory
// instead: then the tensor core reads only the B tiles from shared memory (384 wavefronts) and the
// converters write A_hi / A_lo with tcgen05.st instead of st.shared (another -128).  Price: the A
// stages need TMEM columns, so the 256-column accumulator is SINGLE buffered (MMAs pause while the
// epilogue drains a chunk):   TMEM = D [0,256) + 4 stages x (A_hi 32 | A_lo 32) columns.
// Converter thread r owns row r of the 128-row A tile (its TMEM lane): it reads the row from the raw
// (swizzled) tile, stores the raw bits as A_hi (the tensor core truncates), and x - trunc(x) as A_lo.
//
// RESULT (round 1): bit-identical to the other kernels on every layout, but SLOWER -- 4.02 vs 3.85 ms
// per 8192^3.  The stall counters show why: the MMA warp waits 1313 cycles per k-block for the
// accumulator (tmem_empty).  Draining 128 lanes x 256 columns through tcgen05.ld is bound by the
// TMEM read port (64 B/clk: >= 2048 cycles per 128 KB, ~3000 measured) plus two barrier hops, and
// with one accumulator none of it overlaps the 6144 MMA cycles of a chunk.  The promotion scheme
// (kChunk = 4, needed because the tensor core truncates when it accumulates) therefore REQUIRES a
// double-buffered accumulator, and 2 x 256 columns leave no TMEM for A.  Kept as a working reference
// for a future variant with a different TMEM budget (e.g. 256 x 128 tiles with A multicast); the
// dispatcher never selects it.
constexpr int kTsStages = 4;        // B_lo slots == TMEM A stages

template <int kRaw>
struct SmemTs {
  static constexpr int A_BYTES = BM * BK * 4, B_BYTES = PBN * BK * 4;
  static constexpr int RAW_SLOT = A_BYTES + B_BYTES;
  static constexpr int RAW_BYTES = kRaw * RAW_SLOT, LO_BYTES = kTsStages * B_BYTES;
  static constexpr int TOTAL = RAW_BYTES + LO_BYTES + 512 + 1024;
};

template <int kRaw, bool kTiming>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_3xtf32_pair_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                           const PairParams p) {
  using S = SmemTs<kRaw>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* lo_ring = smem + S::RAW_BYTES;
  uint64_t* raw_full = (uint64_t*)(smem + S::RAW_BYTES + S::LO_BYTES);
  uint64_t* raw_empty = raw_full + kRaw;
  uint64_t* lo_full = raw_empty + kRaw;
  uint64_t* lo_empty = lo_full + kTsStages;
  uint64_t* tmem_full = lo_empty + kTsStages;    // [1]
  uint64_t* tmem_empty = tmem_full + 1;          // [1]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int first_tile = (int)cluster_id_x(), tile_step = (int)num_clusters_x();
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kAcol = 256;                // first A-stage column
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_k = (p.K + BK - 1) / BK;
  Params tp;
  tp.tiles_m = p.tiles_m; tp.tiles_n = p.tiles_n; tp.group_m = p.group_m;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRaw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 1); }
    for (int s = 0; s < kTsStages; ++s) { mbar_init(&lo_full[s], 2 * kPairConvWarps); mbar_init(&lo_empty[s], 1); }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_empty[0], 2 * kPairEpiWarps);
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
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2 || warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      Ring raw;
      long long w_empty = 0, t_all = MDB_T0();
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        int m_blk, n_blk;
        tile_coords(tp, t, m_blk, n_blk);
        const int m0 = m_blk * 256 + (int)rank * BM, n0 = n_blk * 256 + (int)rank * PBN;
        for (int kb = 0; kb < num_k; ++kb) {
          const int k0 = kb * BK;
          const long long tw = MDB_T0();
          mbar_wait(&raw_empty[raw.slot], raw.phase ^ 1);
          MDB_TACC(w_empty, tw);
          const uint32_t a_raw = smem_u32(smem + raw.slot * S::RAW_SLOT), b_raw = a_raw + S::A_BYTES;
          uint64_t* bar = &raw_full[raw.slot];
          mbar_expect_tx(bar, S::RAW_SLOT);
          if (!p.a_mn_major) {
            tma_load_2d(a_raw, &map_a, bar, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(a_raw + c * 4096, &map_a, bar, m0 + 32 * c, k0);
          }
          if (!p.b_mn_major) {
            tma_load_2d(b_raw, &map_b, bar, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < PBN / 32; ++c) tma_load_2d(b_raw + c * 4096, &map_b, bar, n0 + 32 * c, k0);
          }
          raw.advance(kRaw);
        }
      }
      if (kTiming) { p.timing[blockIdx.x * 16 + 0] = w_empty; p.timing[blockIdx.x * 16 + 1] = clock64() - t_all; }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) ======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0) {
      // A comes from tensor memory (K-major by construction); B from shared memory as before
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn_major << 16) |
                             ((uint32_t)((2 * PBN) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t b_lbo = p.b_mn_major ? 4096 : 16, b_sbo = p.b_mn_major ? 512 : 1024;
      const uint32_t b_lt = p.b_mn_major ? 1 : 2, b_kstep = p.b_mn_major ? 1024 : UMMA_K * 4;
      Ring raw, lo;
      uint32_t d_phase = 0;
      long long w_lo = 0, w_tmem = 0, t_all = MDB_T0();
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        for (int kb = 0; kb < num_k; ++kb) {
          const bool chunk_start = (kb % kChunk) == 0;
          const bool chunk_end = ((kb + 1) % kChunk) == 0 || kb == num_k - 1;
          if (chunk_start) {
            const long long tw = MDB_T0();
            mbar_wait_cluster(&tmem_empty[0], d_phase ^ 1);       // both CTAs drained the accumulator
            MDB_TACC(w_tmem, tw);
            tcgen05_fence_after();
          }
          const long long tw2 = MDB_T0();
          mbar_wait_cluster(&lo_full[lo.slot], lo.phase);         // A stage + B_lo ready in both CTAs
          MDB_TACC(w_lo, tw2);
          tcgen05_fence_after();
          if (lane == 0) {
            const uint32_t b_hi = smem_u32(smem + raw.slot * S::RAW_SLOT) + S::A_BYTES;
            const uint32_t b_lo = smem_u32(lo_ring + lo.slot * S::B_BYTES);
            const uint32_t a_hi_t = tmem_base + kAcol + lo.slot * 64, a_lo_t = a_hi_t + 32;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t db_hi = make_desc(b_hi + k * b_kstep, b_lbo, b_sbo, b_lt);
              const uint64_t db_lo = make_desc(b_lo + k * b_kstep, b_lbo, b_sbo, b_lt);
              umma_tf32_pair_ts(tmem_base, a_lo_t + k * UMMA_K, db_hi, idesc, !(chunk_start && k == 0));
              umma_tf32_pair_ts(tmem_base, a_hi_t + k * UMMA_K, db_lo, idesc, 1);
              umma_tf32_pair_ts(tmem_base, a_hi_t + k * UMMA_K, db_hi, idesc, 1);
            }
            umma_commit_pair(&lo_empty[lo.slot]);
            umma_commit_pair(&raw_empty[raw.slot]);
            if (chunk_end) umma_commit_pair(&tmem_full[0]);
          }
          __syncwarp();
          raw.advance(kRaw);
          lo.advance(kTsStages);
          if (chunk_end) d_phase ^= 1;
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
    constexpr int kConv = kPairConvWarps * 32;                     // 128 threads: thread r <-> A row r
    const int t = threadIdx.x - (4 + kPairEpiWarps) * 32;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
    Ring raw, lo;
    long long w_hi = 0, w_lo = 0, t_work = 0, t_sig = 0, t_all = MDB_T0();
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      for (int kb = 0; kb < num_k; ++kb) {
        const long long ta = MDB_T0();
        mbar_wait(&raw_full[raw.slot], raw.phase);
        MDB_TACC(w_hi, ta);
        const long long tb = MDB_T0();
        mbar_wait(&lo_empty[lo.slot], lo.phase ^ 1);
        MDB_TACC(w_lo, tb);
        tcgen05_fence_after();
        const long long tc0 = MDB_T0();
        const uint32_t a_src = smem_u32(smem + raw.slot * S::RAW_SLOT), b_src = a_src + S::A_BYTES;
        const uint32_t b_dst = smem_u32(lo_ring + lo.slot * S::B_BYTES);
        const uint32_t ta_hi = tmem_base + lane_base + kAcol + lo.slot * 64;
        // ---- A: this thread's row, 32 k-values in two halves of 16
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[16];
          if (!p.a_mn_major) {           // K-major, SWIZZLE_128B: 16-B chunk c of row r sits at chunk c ^ (r % 8)
            const uint32_t rowbase = a_src + (uint32_t)t * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(v[4 * c]), "=r"(v[4 * c + 1]), "=r"(v[4 * c + 2]), "=r"(v[4 * c + 3])
                           : "r"(rowbase + ((uint32_t)((half * 4 + c) ^ (t & 7)) << 4)));
          } else {                       // MN-major, 32-B units XOR (k % 4): element (k, m) of chunk m / 32
            const uint32_t colbase = a_src + (uint32_t)(t >> 5) * 4096 + (uint32_t)(t & 7) * 4;
            const uint32_t unit = (uint32_t)((t & 31) >> 3);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int k = half * 16 + i;
              asm volatile("ld.shared.b32 %0, [%1];"
                           : "=r"(v[i])
                           : "r"(colbase + (uint32_t)k * 128 + ((unit ^ (uint32_t)(k & 3)) << 5)));
            }
          }
          MDB_TMEM_ST16(ta_hi + half * 16, v);                       // raw bits: the tensor core truncates
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = __uint_as_float(v[i]);
            const float h = __uint_as_float(v[i] & 0xFFFFE000u);
            v[i] = __float_as_uint(__fsub_rn(x, h));
          }
          MDB_TMEM_ST16(ta_hi + 32 + half * 16, v);
        }
        // ---- B: raw tile -> lo tile, shared memory -> shared memory (same swizzled offsets)
        constexpr int kVecs = S::B_BYTES / 16 / kConv;             // 8 float4 per thread
#pragma unroll
        for (int bt = 0; bt < 2; ++bt) {
          float4 x[kVecs / 2];
#pragma unroll
          for (int j = 0; j < kVecs / 2; ++j)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x[j].x), "=f"(x[j].y), "=f"(x[j].z), "=f"(x[j].w)
                         : "r"(b_src + (t + (bt * (kVecs / 2) + j) * kConv) * 16));
#pragma unroll
          for (int j = 0; j < kVecs / 2; ++j) {
            float e[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
              e[i] = __fsub_rn(e[i], __uint_as_float(__float_as_uint(e[i]) & 0xFFFFE000u));
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(b_dst + (t + (bt * (kVecs / 2) + j) * kConv) * 16),
                         "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3])
                         : "memory");
          }
        }
        MDB_TACC(t_work, tc0);
        const long long td = MDB_T0();
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&lo_full[lo.slot], 0);
        MDB_TACC(t_sig, td);
        raw.advance(kRaw);
        lo.advance(kTsStages);
      }
    }
    if (kTiming && t == 0) {
      unsigned long long* d = p.timing + blockIdx.x * 16;
      d[6] = w_hi; d[7] = w_lo; d[8] = t_work; d[9] = t_sig; d[10] = clock64() - t_all;
    }
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs) ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int q = warp & 3;
    const int eh = (warp - 4) >> 2;
    uint32_t e_phase = 0;
    const bool vec_ok = (p.ldc % 4 == 0) && (((uintptr_t)p.C & 15) == 0);
    const int num_chunks = (num_k + kChunk - 1) / kChunk;
    long long w_full = 0, t_all = MDB_T0();
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      int m_blk, n_blk;
      tile_coords(tp, t, m_blk, n_blk);
      const int row = m_blk * 256 + (int)rank * BM + q * 32 + lane;
      const int n0 = n_blk * 256 + eh * 128;
      float sum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < num_chunks; ++ch) {
        const long long tw = MDB_T0();
        mbar_wait(&tmem_full[0], e_phase);
        MDB_TACC(w_full, tw);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eh * 128 + c * 32);
          MDB_TMEM_LD32(taddr, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] = __fadd_rn(sum[c * 32 + j], __uint_as_float(r[j]));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tmem_empty[0], 0);
        e_phase ^= 1;
      }
      if (row < p.M) {
        float* crow = p.C + (int64_t)row * p.ldc;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
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
    if (kTiming && warp == 4 && lane == 0) {
      p.timing[blockIdx.x * 16 + 11] = w_full; p.timing[blockIdx.x * 16 + 12] = clock64() - t_all;
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

