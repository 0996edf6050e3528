// mlp2x_gelu projector as ONE persistent kernel (llava_arch.py:192-196, multimodal_projector/builder.py:41-48):
//     H   = gelu_erf(X W1^T + b1)        phase 0   X [rows, 1152] bf16,  W1 [3584, 1152]
//     out = H W2^T + b2                  phase 1   H [rows, 3584] bf16,  W2 [3584, 3584]
// Both contractions run on the CTA-pair tcgen05 pipeline of gemm3_sm100.cuh (256 x 256 tiles, TMA-fed 6-stage ring,
// double-buffered TMEM accumulators that keep alternating across the two phases).  The intermediate H of a 128-row
// block is 917 KB - beyond the shared memory + TMEM of an SM pair - so it goes through global memory, but in a tile
// order that keeps it in the 126 MB L2: the host deals the tiles in chunks of `chunk` row blocks,
//     phase-0 tiles of chunk 0 | phase-0 of chunk 1 | phase-1 of chunk 0 | phase-0 of chunk 2 | phase-1 of chunk 1 | ...
// (list scheduling over the CTA pairs, a phase-1 tile charged K2 / K1 of a phase-0 tile), so a chunk of H (29 MB for
// 16 row blocks) is produced, read back by the second GEMM and dead before it would be evicted.
//
// Dependency.  A phase-1 tile of row block m reads all of H[m]: every epilogue warp that has stored its part of a
// phase-0 tile of row block m publishes it (fence.proxy.async + __threadfence + one atomic add on ready[m]); the TMA
// producer of a phase-1 tile spins (bounded, ld.acquire.gpu) until ready[m] has reached warps_per_block, then crosses
// to the async proxy (fence.proxy.async) before its first bulk-tensor load of H.  In the global tile order every
// phase-0 tile of a row block precedes its phase-1 tiles and each CTA pair works through its list in that order, so the
// oldest unfinished tile can always run: no deadlock as long as all pairs are resident (grid <= SM pairs, one CTA per SM).
//
// The same kernel chains the two GEMMs of a ViT MLP block (siglip_encoder.py:252-254, 296-298): phase 0 = fc1 with the
// LayerNorm fold and GELU-tanh, phase 1 = fc2 with the fp32 residual epilogue (+ bf16 stream copy + row statistics); the
// [rows, 4304] activation is then read back from L2 instead of HBM (1 GB less DRAM traffic per layer and tower call).
// Phase-1 rows may end in a 128-wide tile like in gemm3_sm100.cuh (N = 1152 = 4 x 256 + 128).
#pragma once

#include "gemm3_sm100.cuh"

namespace rv {

constexpr int kChainMaxEntries = 8000;  // 80 image tiles: 228 row blocks x (14 + 14) column tiles = 6384

struct ChainSched {
  uint16_t off[kSchedMaxClusters + 1];
  uint16_t ent[kChainMaxEntries];  // row block * 64 + phase * 32 + column tile index
};

struct ChainArgs {
  GemmArgs g[2];         // phase 0 (GELU-erf, bf16 H) and phase 1 (bias, bf16 / f16 / f32 out)
  unsigned int* ready;   // [row blocks], zero on entry: epilogue warps that have published phase-0 tiles
  unsigned int ready_target;  // phase-0 column tiles * 16 warps
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// tmap_w2h: W2 with 64-row boxes (128-wide phase-1 tiles); equal to tmap_w2 when phase 1 has none
template <int EPI1, int EPI2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_chain_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                  const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w2,
                  const __grid_constant__ CUtensorMap tmap_w2h, const __grid_constant__ ChainArgs ca,
                  const __grid_constant__ ChainSched sched) {
  using Cfg = Gemm3Cfg;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileM = 2 * kGemmBM;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base + kStages * Cfg::kStageBytes;
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
  const int e_begin = sched.off[cluster_id], e_end = sched.off[cluster_id + 1];
  const int num_k0 = (ca.g[0].K + kGemmBK - 1) / kGemmBK, num_k1 = (ca.g[1].K + kGemmBK - 1) / kGemmBK;
  const int num_n1 = (ca.g[1].N + kSchedBN - 1) / kSchedBN;
  const int last_w1 = (ca.g[1].N - (num_n1 - 1) * kSchedBN <= 128) ? 128 : kSchedBN;
  auto tile_of = [&](int e, int& ph, int& m_blk, int& n0, int& w) {
    const int v = sched.ent[e];
    m_blk = v >> 6;
    ph = (v >> 5) & 1;
    const int n_idx = v & 31;
    n0 = n_idx * kSchedBN;
    w = (ph == 1 && n_idx == num_n1 - 1) ? last_w1 : kSchedBN;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_w2);
    tma_prefetch_desc(&tmap_w2h);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * kGemmEpiWarps);
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
      for (int e = e_begin; e < e_end; ++e) {
        int ph, m_blk, n0, w;
        tile_of(e, ph, m_blk, n0, w);
        const CUtensorMap* ta = ph ? &tmap_h : &tmap_x;
        const CUtensorMap* tb = ph ? (w == kSchedBN ? &tmap_w2 : &tmap_w2h) : &tmap_w1;
        const int nk = ph ? num_k1 : num_k0;
        const int row_a = m_blk * kTileM + static_cast<int>(rank) * kGemmBM;
        const int row_b = n0 + static_cast<int>(rank) * (w / 2);
        const uint32_t tx = 2u * static_cast<uint32_t>(Cfg::kABytes + (w / 2) * kGemmBK * 2);
        if (ph) {  // H[m_blk] must be complete (and visible to the async proxy) before the first load of it
          const long long t0 = clock64();
          while (ld_acquire_gpu(ca.ready + m_blk) < ca.ready_target) {
            __nanosleep(64);
            if (clock64() - t0 > 6000000000LL) __trap();
          }
          fence_proxy_async_all();
        }
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), tx);
          else mbar_arrive_remote(full_bar(stage), 0);
          tma_load_2d_cta2(sa, ta, full_bar(stage), kb * kGemmBK, row_a);
          tma_load_2d_cta2(sb, tb, full_bar(stage), kb * kGemmBK, row_b);
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
      constexpr uint32_t idesc_full = make_idesc_bf16(kTileM, kSchedBN);
      constexpr uint32_t idesc_half = make_idesc_bf16(kTileM, 128);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t desc_k = make_smem_desc(smem_base, 1024, kLayoutSw128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int e = e_begin; e < e_end; ++e) {
        int ph, m_blk, n0, w;
        tile_of(e, ph, m_blk, n0, w);
        const uint32_t idesc = (w == kSchedBN) ? idesc_full : idesc_half;
        const int nk = ph ? num_k1 : num_k0;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + static_cast<uint32_t>(acc * kSchedBN);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t soff = static_cast<uint64_t>((stage * Cfg::kStageBytes) >> 4);
          const uint64_t adesc = desc_k + soff;
          const uint64_t bdesc = adesc + static_cast<uint64_t>(Cfg::kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_bf16_ss_cta2_elect(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_cta2_mc_elect(empty_bar(stage), 3);
          if (kb == nk - 1) umma_commit_cta2_mc_elect(tfull_bar(acc), 3);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
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
    // phase-1 fp32 residual epilogue: pull this warp's block of the residual into L2 ahead of its use (gemm3_sm100.cuh)
    auto prefetch_resid = [&](int e) {
      if constexpr (EPI2 == EPI_RESID_F32) {
        int ph, m_blk, n0, w;
        tile_of(e, ph, m_blk, n0, w);
        if (ph == 0) return;
        const int row0 = m_blk * kTileM + static_cast<int>(rank) * kGemmBM + quad * 32;
        const int col0 = n0 + half * (w / 2);
        const int lines = w / 64;
        for (int i = lane; i < 32 * lines; i += 32) {
          const int r = row0 + i / lines, c = col0 + (i % lines) * 32;
          if (r < ca.g[1].M && c < ca.g[1].N) prefetch_l2(ca.g[1].aux + static_cast<size_t>(r) * ca.g[1].ldo + c);
        }
      }
    };
    for (int e = e_begin; e < e_end; ++e) {
      int ph, m_blk, n0, w;
      tile_of(e, ph, m_blk, n0, w);
      if (e + 1 < e_end) prefetch_resid(e + 1);
      const int row = m_blk * kTileM + static_cast<int>(rank) * kGemmBM + quad * 32 + lane;
      float ln_nmean = 0.f, ln_rstd = 1.f;   // LayerNorm folded into phase 0: row statistics fetched before the wait
      if (ph == 0) gemm_ln_row_stats(ca.g[0], row, ln_nmean, ln_rstd);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * kSchedBN + half * (w / 2));
      const uint32_t stg = stage_base + static_cast<uint32_t>(warp - 2) * kGemmStageWarpBytes;
      const int col = n0 + half * (w / 2);
      if (ph == 0) {
        gemm_epilogue_drain<EPI1, kSchedBN / 2>(ca.g[0], row, col, t_row, stg, lane, -1, ln_nmean, ln_rstd);
      } else {
        const int ln_slot = 2 * (n0 / kSchedBN) + half;   // row-statistics slot (gemm_args.h ln_part)
        if (w == kSchedBN) gemm_epilogue_drain<EPI2, kSchedBN / 2>(ca.g[1], row, col, t_row, stg, lane, ln_slot);
        else gemm_epilogue_drain<EPI2, 64>(ca.g[1], row, col, t_row, stg, lane, ln_slot);
      }
      tc_fence_before();
      __syncwarp();
      if (ph == 0) {  // publish this warp's part of H[m_blk]: generic-proxy stores -> async-proxy (TMA) readers elsewhere
        fence_proxy_async_all();
        __threadfence();
        if (lane == 0) atomicAdd(ca.ready + m_blk, 1u);
      }
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
