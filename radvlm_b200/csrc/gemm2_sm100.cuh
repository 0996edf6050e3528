// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a CTA pair (one cluster of 2, same TPC)
// computes a 256 x BN output tile with UMMA 256 x BN x 16.  Each CTA loads its own 128 rows of A and HALF of
// the B tile (BN/2 rows); the tensor core reads the other half from the peer's shared memory, so the
// L2 -> SMEM traffic per FLOP drops by 1/3 against the 128 x BN single-CTA kernel (which is L2-bandwidth
// bound at ~1 PFLOP/s) and the same shared memory buffers 1.5x more MMA time.
//
// Protocol (per pipeline stage):
//   * both producers wait on their OWN empty barrier, then issue TMA with .cta_group::2 so the transaction
//     bytes land on the LEADER's full barrier (peer bit of the barrier address cleared);
//     the leader arrives with expect_tx(bytes of both CTAs), the peer arrives remotely (count = 2)
//   * the leader's MMA thread waits on its full barrier, issues tcgen05.mma.cta_group::2 and commits with
//     multicast to the empty barrier of BOTH CTAs; the last K slab also commits to both tmem_full barriers
//   * epilogue warps of both CTAs drain their own TMEM half (rows rank*128 ..) and arrive on the LEADER's
//     tmem_empty barrier (count = 16 warps)
#pragma once

#include "gemm_sm100.cuh"

namespace rv {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster addresses of a CTA pair differ in bit 24

__device__ __forceinline__ void tma_load_2d_cta2(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cta2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cta2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cta2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_cta2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Whole-warp issue variants (see umma_bf16_ss_elect in common.cuh): the converged warp executes the block, elect.sync
// picks the issuing lane, ptxas emits back-to-back UTCHMMA on uniform registers (no ELECT / R2UR loop per MMA).
__device__ __forceinline__ void umma_bf16_ss_cta2_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred P1, p;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@P1 tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cta2_mc_elect(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "@P1 tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cta2_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;   // 16 KB: this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * kGemmBK * 2;  // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 128) ? 8 : 6;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kGemmStagingBytes + 1024 + 256;
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmArgs args) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileM = 2 * kGemmBM;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base + kStages * Cfg::kStageBytes;  // epilogue transpose buffers
  const uint32_t bar_base = stage_base + kGemmStagingBytes;
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
  const int num_clusters = gridDim.x >> 1;

  const int num_m = (args.M + kTileM - 1) / kTileM;
  const int num_n = (args.N + BN - 1) / BN;
  const int num_mn = num_m * num_n;
  const int num_k_all = (args.K + kGemmBK - 1) / kGemmBK;
  // Work items.  Whole tiles (k_splits <= 1): cluster c takes tiles c, c + C, ...  Split-K (k_splits > 1, atomic fp32
  // epilogue only): item t = split * num_mn + tile, split s covers K slabs [s * k_per, min(num_k_all, (s+1) * k_per)).
  // Items are dealt round-robin, so the clusters of a wave work on the SAME K range of different tiles: every operand
  // slab is fetched from HBM once and shared through L2.  (A "stream-K" deal - equal contiguous (tile, slab) ranges per
  // cluster - was measured: perfectly balanced but 21 % slower, the clusters then sit at different K offsets and each
  // re-reads its operand rows from HBM.)  for_each_segment is evaluated identically by the producer, the MMA issuer and
  // the epilogue warps.
  const int k_splits = args.k_splits > 1 ? args.k_splits : 1;
  const int k_per = (num_k_all + k_splits - 1) / k_splits;
  auto for_each_segment = [&](auto&& body) {
    const int num_items = num_mn * k_splits;
    for (int t = cluster_id; t < num_items; t += num_clusters) {
      const int split = t / num_mn;
      const int kb0 = split * k_per;
      body(t - split * num_mn, kb0, min(num_k_all, kb0 + k_per));
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
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

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for_each_segment([&](int tile, int kb0, int kb1) {
        const int m_blk = tile / num_n;
        const int n_blk = tile - m_blk * num_n;
        const int row_a = m_blk * kTileM + static_cast<int>(rank) * kGemmBM;
        const int row_b = n_blk * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          else mbar_arrive_remote(full_bar(stage), 0);
          if (!args.a_mn) {
            tma_load_2d_cta2(sa, &tmap_a, full_bar(stage), kb * kGemmBK, row_a);
          } else {  // A stored [K, M]: boxes of {64 M-elements, 64 K-rows}, one per 64 rows of the tile
#pragma unroll
            for (int i = 0; i < kGemmBM / 64; ++i)
              tma_load_2d_cta2(sa + i * 8192, &tmap_a, full_bar(stage), row_a + 64 * i, kb * kGemmBK);
          }
          if (!args.b_mn) {
            tma_load_2d_cta2(sb, &tmap_b, full_bar(stage), kb * kGemmBK, row_b);
          } else {  // W stored [K, N]: boxes of {64 N-elements, 64 K-rows}
#pragma unroll
            for (int i = 0; i < BN / 2 / 64; ++i)
              tma_load_2d_cta2(sb + i * 8192, &tmap_b, full_bar(stage), row_b + 64 * i, kb * kGemmBK);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      });
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the loop converged (all operands are warp-uniform); elect.sync inside the asm blocks picks
    // the lane that issues.
    if (leader) {
      // instruction descriptor: bits 15 / 16 select MN-major A / B
      const uint32_t idesc = make_idesc_bf16(kTileM, BN) | (args.a_mn ? (1u << 15) : 0u) | (args.b_mn ? (1u << 16) : 0u);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      // K-major operand: 8-row groups 1024 B apart, 32 B per 16-element K step inside the 128B swizzle atom.
      // MN-major operand (TMA boxes of 64 MN-elements x 64 K-rows, 8 KB each): 8 K-rows per 1024 B atom (SBO),
      // 64-element MN groups 8192 B apart (LBO), 2048 B per 16-row K step.
      const uint64_t desc_k = make_smem_desc(smem_base, 1024, kLayoutSw128);
      const uint64_t desc_mn = (desc_k & ~(static_cast<uint64_t>(0x3FFF) << 16)) | (static_cast<uint64_t>(8192 >> 4) << 16);
      const uint64_t a0 = args.a_mn ? desc_mn : desc_k, b0 = args.b_mn ? desc_mn : desc_k;
      const uint64_t a_step = args.a_mn ? (2048 >> 4) : 2, b_step = args.b_mn ? (2048 >> 4) : 2;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for_each_segment([&](int /*tile*/, int kb0, int kb1) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + static_cast<uint32_t>(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t soff = static_cast<uint64_t>((stage * Cfg::kStageBytes) >> 4);
          const uint64_t adesc = a0 + soff;
          const uint64_t bdesc = b0 + soff + static_cast<uint64_t>(Cfg::kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_bf16_ss_cta2_elect(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          umma_commit_cta2_mc_elect(empty_bar(stage), 3);
          if (kb == kb1 - 1) umma_commit_cta2_mc_elect(tfull_bar(acc), 3);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      });
    }
  } else {
    // ===================== epilogue (4 warps in each CTA) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for_each_segment([&](int tile, int /*kb0*/, int /*kb1*/) {
      const int m_blk = tile / num_n;
      const int n_blk = tile - m_blk * num_n;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * kTileM + static_cast<int>(rank) * kGemmBM + quad * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + half * (BN / 2));
      float ln_nmean, ln_rstd;
      gemm_ln_row_stats(args, row, ln_nmean, ln_rstd);
      gemm_epilogue_drain<EPI, BN / 2>(args, row, n_blk * BN + half * (BN / 2), t_row,
                                       stage_base + static_cast<uint32_t>(warp - 2) * kGemmStageWarpBytes, lane, -1,
                                       ln_nmean, ln_rstd);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_remote(tempty_bar(acc), 0);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    });
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cta2(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace rv
