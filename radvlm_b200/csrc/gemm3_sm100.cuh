// Scheduled, variable-width variant of the CTA-pair GEMM (gemm2_sm100.cuh).
//
// Why.  N = 1152 (out_proj, fc2, the patch embedding and three of the four data-gradient GEMMs) is 4.5 x 256: with
// 256-wide tiles 10 % of the tensor work is padding, with 192-wide tiles a 256 x 192 pair MMA streams 10 KB of operands
// per SM in 96 cycles (104 B/clk, above the ~95 B/clk shared memory sustains), so both forms ran at ~0.78 of the
// sustained peak against 0.90 for the 256-wide fc1 (profiles/r02b_bench_head.json).  Here a row block is cut exactly:
// full 256-wide tiles plus ONE 128-wide tile when the remainder is <= 128 columns (1152 = 4 x 256 + 128,
// 3456 = 13 x 256 + 128).  8/9 of the work runs on the tile shape that reaches 0.90; nothing is padded.
//
// Tiles of two widths cannot be dealt round-robin (a 128-wide tile costs ~2/3 of a full one).  The host simulates
// list scheduling once per shape (tiles in row-block-major order, each to the least loaded CTA pair, gemm_capi.cu) and
// passes every pair its tile list in the kernel parameters (__grid_constant__, 8 KB): the kernel stays free of
// atomics and its tile order - hence its L2 reuse pattern - stays the row-block-major one of the round-robin kernel.
//
// Also new here: the fp32 residual epilogue (out = acc + bias + resid) prefetches the residual lines of the NEXT tile
// into L2 before it drains the current one.  The epilogue handles a tile in 32-column chunks, each a dependent chain
// TMEM load -> transpose -> residual load -> store; the residual was written a whole layer ago and misses L2, so with
// K = 1152 (out_proj: 6900 cycles of MMA per tile) three to four serialised HBM latencies per tile made the epilogue
// the critical path (out_proj at 0.62 of the sustained peak).
#pragma once

#include "gemm2_sm100.cuh"

namespace rv {

constexpr int kSchedMaxEntries = 4000;  // tiles per launch (80 image tiles: 228 row blocks x 17 column tiles for fc1)
constexpr int kSchedMaxClusters = 75;
constexpr int kSchedBN = 256;

struct GemmSched {
  uint16_t off[kSchedMaxClusters + 1];  // tiles of CTA pair c: ent[off[c]] .. ent[off[c + 1] - 1], in issue order
  uint16_t ent[kSchedMaxEntries];       // row block * 32 + column tile index
};

struct Gemm3Cfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;         // 16 KB: this CTA's 128 rows of A
  static constexpr int kBBytes = (kSchedBN / 2) * kGemmBK * 2;  // 16 KB: this CTA's half of a full-width B tile
  static constexpr int kBBoxBytes = 64 * kGemmBK * 2;           // B is fetched in boxes of 64 rows (or 64 columns)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 6;
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kGemmStagingBytes + 1024 + 256;
};

// Residual prefetch of the bf16-delta epilogue (out_proj).  Its drain was a chain of four dependent chunks per warp, each
// TMEM load -> transpose -> residual load from global memory -> store: 17.7 K cycles per 256 x 256 tile against a
// 9.2 K-cycle main loop (profiles/r02q_ncu_full_tower_summary.csv: 53 % tensor pipe, 41 % DRAM - neither roofline).
// With RV_DELTA_PF every epilogue warp owns a ring of four 2 KB chunk slots ([32 rows][64 B] of the bf16 residual) filled
// by cp.async: chunk q of the NEXT tile is requested as soon as chunk q of the current tile has been consumed, so the
// loads fly under the other chunks' work and across the tile boundary, and the drain no longer waits for DRAM.  The
// 64 KB come out of the operand ring (4 stages instead of 6: K = 1152 is 18 slabs, the kernel is not load-bound).
#ifndef RV_DELTA_PF
#define RV_DELTA_PF 1
#endif
constexpr int kResidChunkBytes = 32 * 64;                  // 32 rows x 32 bf16 columns
constexpr int kResidWarpBytes = 4 * kResidChunkBytes;      // chunks 0..3 of a warp's 32 x 128 block
template <int EPI>
struct Gemm3CfgT : Gemm3Cfg {
  static constexpr bool kResidPf = (RV_DELTA_PF != 0) && (EPI == EPI_DELTA_BF16);
  static constexpr int kStages = kResidPf ? 4 : Gemm3Cfg::kStages;
  static constexpr int kResidBytes = kResidPf ? kGemmEpiWarps * kResidWarpBytes : 0;
  static constexpr int kSmemBytes = kStages * Gemm3Cfg::kStageBytes + kGemmStagingBytes + kResidBytes + 1024 + 256;
};
static_assert(Gemm3CfgT<EPI_DELTA_BF16>::kSmemBytes <= 232448, "shared memory of the delta epilogue variant");

// One chunk ([32 rows][32 columns] bf16) of the residual block of a warp: 128 x 16 B, four per lane.
__device__ __forceinline__ void delta_prefetch_chunk(const GemmArgs& a, int row0, int col0, uint32_t slot, int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = lane + 32 * j;
    const int grow = row0 + (idx >> 2), gcol = col0 + (idx & 3) * 8;
    const bool ok = grow < a.M && gcol < a.N;
    const __nv_bfloat16* src = ok ? a.aux16 + static_cast<size_t>(grow) * a.ldo + gcol : a.aux16;
    cp_async_16(slot + static_cast<uint32_t>(idx) * 16u, src, ok ? 16u : 0u);
  }
}

// Drain of one tile of the bf16-delta epilogue with the residual ring (see above).  cur / nxt = 32-column chunks of this
// warp's block in this / the next tile (4 for a full tile, 2 for the 128-wide one, 0 = no next tile).  Four cp.async
// groups are committed per tile whatever its width, so exactly three groups are younger than the chunk being consumed.
__device__ __forceinline__ void gemm_epilogue_drain_delta_pf(const GemmArgs& args, int row, int col_base, int cur,
                                                             int nxt_row0, int nxt_col_base, int nxt, uint32_t t_row,
                                                             uint32_t stage, uint32_t ring, int lane, int ln_slot) {
  const int row0 = row - lane;
  const bool stats = args.ln_part != nullptr && ln_slot >= 0;
  f32x2 ps[8], pq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0ull;
  uint32_t r0[32], r1[32];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t slot = ring + static_cast<uint32_t>(q) * kResidChunkBytes;
    if (q < cur) {
      if ((q & 1) == 0) {
        tmem_ld_x32(t_row + 32 * q, r0);
        if (q + 1 < cur) tmem_ld_x32(t_row + 32 * (q + 1), r1);
        tmem_wait_ld();
      }
      cp_async_wait_group<3>();
      __syncwarp();
      const uint32_t* acc = (q & 1) ? r1 : r0;
      if (stats) gemm_epilogue_f32_staged<EPI_DELTA_BF16, true, true>(args, row0, col_base + 32 * q, acc, stage, lane, ps, pq, slot);
      else gemm_epilogue_f32_staged<EPI_DELTA_BF16, false, true>(args, row0, col_base + 32 * q, acc, stage, lane, nullptr, nullptr, slot);
      // (ends with __syncwarp: every lane has read the slot)
    }
    if (q < nxt) delta_prefetch_chunk(args, nxt_row0, nxt_col_base + 32 * q, slot, lane);
    cp_async_commit();
  }
  if (stats) {  // as in gemm_epilogue_drain: the 8 lanes that share a row add up, lane & 7 == 0 writes the slot
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s0, s1, q0, q1;
      f2_get(ps[i], s0, s1);
      f2_get(pq[i], q0, q1);
      float sum = s0 + s1, sq = q0 + q1;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      const int grow = row0 + i * 4 + (lane >> 3);
      if ((lane & 7) == 0 && grow < args.M)
        args.ln_part[static_cast<size_t>(grow) * args.ln_slots + ln_slot] = make_float2(sum, sq);
    }
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// (the bulk form, cp.async.bulk.prefetch.L2 with one row segment per lane, was measured too: no shorter drain and
// ~2400 cycles of issue per tile against ~770; tools/gemm_timeline.cu)

#ifdef RV_GEMM_TIMELINE
#define RV_GTL(slot)                                                                                      \
  do {                                                                                                    \
    if (args.timeline != nullptr && cluster_id == RV_GEMM_TIMELINE && leader && lane == 0 && (e - e_begin) < 24) \
      args.timeline[(e - e_begin) * 8 + (slot)] = clock64();                                              \
  } while (0)
#else
#define RV_GTL(slot) do { } while (0)
#endif

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_2cta_sched_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                               const __grid_constant__ CUtensorMap tmap_b64, const GemmArgs args,
                               const __grid_constant__ GemmSched sched) {
  using Cfg = Gemm3CfgT<EPI>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileM = 2 * kGemmBM;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base + kStages * Cfg::kStageBytes;  // epilogue transpose buffers
  const uint32_t resid_base = stage_base + kGemmStagingBytes;          // residual rings of the delta epilogue (or empty)
  const uint32_t bar_base = resid_base + Cfg::kResidBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x >> 1;

  const int num_n = (args.N + kSchedBN - 1) / kSchedBN;
  const int last_w = (args.N - (num_n - 1) * kSchedBN <= 128) ? 128 : kSchedBN;
  const int num_k = (args.K + kGemmBK - 1) / kGemmBK;
  const int e_begin = sched.off[cluster_id], e_end = sched.off[cluster_id + 1];
  // tile e of this pair: row block, first column, width (256, or 128 for the last column tile of a row block)
  auto tile_of = [&](int e, int& m_blk, int& n0, int& w) {
    const int v = sched.ent[e];
    m_blk = v >> 5;
    const int n_idx = v & 31;
    n0 = n_idx * kSchedBN;
    w = (n_idx == num_n - 1) ? last_w : kSchedBN;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_b64);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 2);   // leader: own arrive.expect_tx + the peer's remote arrive
      mbar_init(empty_bar(s), 1);  // one multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * kGemmEpiWarps);  // epilogue warps of both CTAs (leader's copy is the one used)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cta2(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish_cta2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();

  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  // PDL: the prologue above ran under the previous kernel's tail; from here on global memory is touched
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int e = e_begin; e < e_end; ++e) {
        int m_blk, n0, w;
        tile_of(e, m_blk, n0, w);
        const int row_a = m_blk * kTileM + static_cast<int>(rank) * kGemmBM;
        const int row_b = n0 + static_cast<int>(rank) * (w / 2);
        const int boxes = w / 128;  // 64-row (64-column) boxes of B per CTA: 2 for a full tile, 1 for a 128-wide one
        const uint32_t tx = 2u * static_cast<uint32_t>(Cfg::kABytes + boxes * Cfg::kBBoxBytes);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), tx);
          else mbar_arrive_remote(full_bar(stage), 0);
          tma_load_2d_cta2(sa, &tmap_a, full_bar(stage), kb * kGemmBK, row_a);
          if (!args.b_mn) {  // W [N, K]: one box of {64 K-elements, 128 rows} (tmap_b) or {64, 64} (tmap_b64)
            if (boxes == 2) tma_load_2d_cta2(sb, &tmap_b, full_bar(stage), kb * kGemmBK, row_b);
            else tma_load_2d_cta2(sb, &tmap_b64, full_bar(stage), kb * kGemmBK, row_b);
          } else {           // W stored [K, N]: boxes of {64 N-elements, 64 K-rows}
            for (int i = 0; i < boxes; ++i)
              tma_load_2d_cta2(sb + i * Cfg::kBBoxBytes, &tmap_b, full_bar(stage), row_b + 64 * i, kb * kGemmBK);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp converged, elect.sync inside) =====================
    if (leader) {
      const uint32_t idesc_full = make_idesc_bf16(kTileM, kSchedBN) | (args.b_mn ? (1u << 16) : 0u);
      const uint32_t idesc_half = make_idesc_bf16(kTileM, 128) | (args.b_mn ? (1u << 16) : 0u);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t desc_k = make_smem_desc(smem_base, 1024, kLayoutSw128);
      const uint64_t desc_mn = (desc_k & ~(static_cast<uint64_t>(0x3FFF) << 16)) | (static_cast<uint64_t>(8192 >> 4) << 16);
      const uint64_t b0 = args.b_mn ? desc_mn : desc_k;
      const uint64_t b_step = args.b_mn ? (2048 >> 4) : 2;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int e = e_begin; e < e_end; ++e) {
        int m_blk, n0, w;
        tile_of(e, m_blk, n0, w);
        const uint32_t idesc = (w == kSchedBN) ? idesc_full : idesc_half;
        RV_GTL(0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        RV_GTL(1);
        const uint32_t d_tmem = tmem_u + static_cast<uint32_t>(acc * kSchedBN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (kb == 0) RV_GTL(2);
          const uint64_t soff = static_cast<uint64_t>((stage * Cfg::kStageBytes) >> 4);
          const uint64_t adesc = desc_k + soff;
          const uint64_t bdesc = b0 + soff + static_cast<uint64_t>(Cfg::kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_bf16_ss_cta2_elect(d_tmem, adesc + 2 * k, bdesc + b_step * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_cta2_mc_elect(empty_bar(stage), 3);
          if (kb == num_k - 1) umma_commit_cta2_mc_elect(tfull_bar(acc), 3);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        RV_GTL(3);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue (8 warps in each CTA) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    // fp32 residual epilogue: pull this warp's 32 x (w / 2) block of the residual into L2 ahead of its use
    auto prefetch_resid = [&](int e) {
      if constexpr (EPI == EPI_RESID_F32) {
        int m_blk, n0, w;
        tile_of(e, m_blk, n0, w);
        const int row0 = m_blk * kTileM + static_cast<int>(rank) * kGemmBM + quad * 32;
        const int col0 = n0 + half * (w / 2);
        const int lines = w / 64;  // 128-byte lines per row of the block
        for (int i = lane; i < 32 * lines; i += 32) {
          const int r = row0 + i / lines, c = col0 + (i % lines) * 32;
          if (r < args.M && c < args.N) prefetch_l2(args.aux + static_cast<size_t>(r) * args.ldo + c);
        }
      }
    };
    if (e_begin < e_end) prefetch_resid(e_begin);
    // delta epilogue with the residual ring: works on full 16-byte vectors only (N, ldo multiples of 8: always, in the tower)
    const bool delta_pf = Cfg::kResidPf && ((args.N & 7) == 0) && ((args.ldo & 7) == 0);
    const uint32_t ring = resid_base + static_cast<uint32_t>(warp - 2) * kResidWarpBytes;
    if constexpr (Cfg::kResidPf) {
      if (delta_pf) {   // the first tile's four chunk groups
        int m_blk = 0, n0 = 0, w = 0;
        if (e_begin < e_end) tile_of(e_begin, m_blk, n0, w);
        const int r0w = m_blk * kTileM + static_cast<int>(rank) * kGemmBM + quad * 32, c0w = n0 + half * (w / 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (e_begin < e_end && q < w / 64) delta_prefetch_chunk(args, r0w, c0w + 32 * q, ring + q * kResidChunkBytes, lane);
          cp_async_commit();
        }
      }
    }
    for (int e = e_begin; e < e_end; ++e) {
      int m_blk, n0, w;
      tile_of(e, m_blk, n0, w);
      if (e + 1 < e_end) prefetch_resid(e + 1);
      const int row = m_blk * kTileM + static_cast<int>(rank) * kGemmBM + quad * 32 + lane;
      float ln_nmean, ln_rstd;  // LayerNorm fold: this row's statistics, fetched while the tile's MMAs still run
      gemm_ln_row_stats(args, row, ln_nmean, ln_rstd);
      if (warp == 2) RV_GTL(4);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (warp == 2) RV_GTL(5);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * kSchedBN + half * (w / 2));
      const uint32_t stg = stage_base + static_cast<uint32_t>(warp - 2) * kGemmStageWarpBytes;
      const int ln_slot = 2 * (n0 / kSchedBN) + half;  // one statistics slot per (column tile, half): gemm_args.h ln_part
      bool drained = false;
      if constexpr (Cfg::kResidPf) {
        if (delta_pf) {
          int m2 = 0, n2 = 0, w2 = 0;
          if (e + 1 < e_end) tile_of(e + 1, m2, n2, w2);
          gemm_epilogue_drain_delta_pf(args, row, n0 + half * (w / 2), w / 64,
                                       m2 * kTileM + static_cast<int>(rank) * kGemmBM + quad * 32, n2 + half * (w2 / 2), w2 / 64,
                                       t_row, stg, ring, lane, ln_slot);
          drained = true;
        }
      }
      if (!drained) {
        if (w == kSchedBN) gemm_epilogue_drain<EPI, kSchedBN / 2>(args, row, n0 + half * (kSchedBN / 2), t_row, stg, lane, ln_slot,
                                                                  ln_nmean, ln_rstd);
        else gemm_epilogue_drain<EPI, 64>(args, row, n0 + half * 64, t_row, stg, lane, ln_slot, ln_nmean, ln_rstd);
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 2) RV_GTL(6);
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_remote(tempty_bar(acc), 0);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cta2(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace rv
