// Backward of the SigLIP attention (siglip_encoder.py:216-235 under autograd) on sm_100a tensor cores.
//
//   P  = softmax(Q K^T * scale)            recomputed from the forward's log-sum-exp:  P = 2^(S * scale * log2e - lse)
//   dV = P^T dO        dP = dO V^T         dS = P o (dP - delta) * scale,  delta = rowsum(dO o O)
//   dQ = dS K          dK = dS^T Q
//
// One CTA owns one 128-key block j of one (tile, head): dK_j and dV_j accumulate in TMEM over the query blocks,
// dQ_i partial products are added to an fp32 buffer (every key block contributes to every query block).  The
// reduction goes through the TMA unit: the 128 x 80 fp32 partial tile is staged in shared memory and ONE bulk
// reduce-add (cp.reduce.async.bulk ... add.f32, UBLKRED) per query block adds it to global memory as whole lines -
// round 1 issued 10 red.global.add.v4.f32 per thread and was bound by them (943 MB of LSU-issued reductions per
// 40-tile launch, lg_throttle 2.0 / long_scoreboard 7.8 stalls per issue, tensor pipe 20 %).  The six key-block CTAs
// of a (tile, head) walk the query blocks in rotated order (i = (j + t) mod 6), so at any time they reduce into six
// different dQ blocks instead of contending for the same lines.
// All five products run on tcgen05; the operands are read in place, without transposed copies:
//   * Q_i, K_j, dO_i sit in shared memory as five [128 rows x 32 B] 32B-swizzled chunks (dO_i is fetched by TMA
//     straight out of the token-major [tokens, heads*hd] gradient of the attention output).  The same bytes are a
//     K-major operand when the contraction runs over head_dim (S, dP) and an MN-major operand when it runs over
//     the rows (dV = P^T dO, dK = dS^T Q, dQ = dS K).
//   * V_j is loaded like K_j (five chunks) and is a plain K-major B operand of dP = dO V^T.
//   * P and dS are written by the softmax warps as bf16 [q][key] tiles (128B swizzle): K-major A for dQ = dS K,
//     MN-major A for P^T dO and dS^T Q.
// Warp roles: warp0 TMA, warp1 MMA issue (whole warp, elect.sync), warps 2..9 softmax gradient (two threads per
// query row), warps 10..13 dQ read-out (one thread per row: TMEM -> staging tile -> bulk reduce-add), off the
// softmax warps' critical path.  Per step the MMA warp issues S_{t+1} / dP_{t+1} as soon as S_t / dP_t have been read
// out of TMEM (half-way through the softmax of step t), then dQ_t / dV_t / dK_t once P_t / dS_t are published: the
// softmax of step t+1 runs while those execute (it waits for them only before overwriting P / dS).
// TMEM: S [0,128)  dP [128,256)  dV [256,336)  dK [336,416)  dQ [416,496).
//
// Padding contract: Q / K / V pad rows and head-dim pad columns are zero (V WITHOUT the forward's ones column: use a
// buffer prepared with plain zeros); rows of dO beyond this tile (the next tile's tokens) are neutralised by zeroing
// the P / dS rows of invalid queries.
#pragma once

#include "common.cuh"

namespace rv {

struct AttnBwdArgs {
  const float* lse;      // [tiles*heads, seq_pad] base-2 log-sum-exp from the forward
  const float* delta;    // [tiles*heads, seq_pad] rowsum(dO o O)
  float* dq_acc;         // [tiles*heads, seq_pad, kAbDqPitch] fp32, zero-initialised
  __nv_bfloat16* dqkv;   // [tiles*seq, 3*heads*hd]: this kernel writes the dK and dV column blocks
  int seq, seq_pad, heads, hd;
  float scale, scale_log2e;
#ifdef RV_ABWD_TIMELINE
  long long* timeline;   // tuning builds only (tools/attn_bwd_timeline.cu): clock64 stamps of the CTAs of one (tile, head)
#endif
};

#ifdef RV_ABWD_TIMELINE
#define RV_ABTL(step, slot)                                                                                   \
  do {                                                                                                        \
    if (args.timeline != nullptr && tile == RV_ABWD_TIMELINE && head == 0 && lane == 0)                       \
      args.timeline[jblk * 128 + (step) * 16 + (slot)] = clock64();                                            \
  } while (0)
#else
#define RV_ABTL(step, slot) do { } while (0)
#endif

constexpr int kAbThreads = 448;
constexpr int kAbDqPitch = 84;           // floats per dq_acc row: 80 + 4 so that rows 336 B apart spread over all banks
constexpr int kAbTile = 128 * 80 * 2;   // 20480: Q / K / V / dO tiles (5 SW32 chunks each)
constexpr int kAbPTile = 128 * 128 * 2; // 32768: P and dS
constexpr int kAbDqTile = 64 * kAbDqPitch * 4;   // 21504: fp32 staging of HALF a dQ block (64 rows) per bulk reduction
constexpr int kAbQBufs = 2, kAbDoBufs = 3;       // Q is released right after dK (its last reader); dO (strided rows of
                                                 // the token-major gradient: the slow load) gets a third buffer
constexpr int kAbSmemBytes = (2 + kAbQBufs + kAbDoBufs) * kAbTile + 2 * kAbPTile + kAbDqTile + 256;
static_assert(kAbSmemBytes <= 227 * 1024, "shared memory budget");
constexpr int kAbTmemCols = 512;

__device__ __forceinline__ float ex2_approx_bwd(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAbThreads, 1)
siglip_attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q,   // Q  [th*seq_pad, 80]   box {16, 128} SW32
                            const __grid_constant__ CUtensorMap tmap_k,   // K  same
                            const __grid_constant__ CUtensorMap tmap_v,   // V  same
                            const __grid_constant__ CUtensorMap tmap_do,  // dO [tiles*seq, heads*hd] box {16, 128} SW32
                            const AttnBwdArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t sK = smem_base;
  const uint32_t sV = sK + kAbTile;
  const uint32_t sQ = sV + kAbTile;                   // Q_t in sQ + (t % 2) * kAbTile
  const uint32_t sdO = sQ + kAbQBufs * kAbTile;       // dO_t in sdO + (t % 3) * kAbTile
  const uint32_t sP = sdO + kAbDoBufs * kAbTile;
  const uint32_t sdS = sP + kAbPTile;
  const uint32_t sDQ = sdS + kAbPTile;       // [64 q][kAbDqPitch] fp32, the layout of the rows in dq_acc
  const uint32_t bar_base = sDQ + kAbDqTile;
  const uint32_t bar_kv = bar_base + 0;       // K_j, V_j landed
  const uint32_t bar_q = bar_base + 8;        // [2] Q_t landed
  const uint32_t bar_do = bar_base + 24;      // [3] dO_t landed
  const uint32_t bar_sdp = bar_base + 48;     // S, dP complete in TMEM
  const uint32_t bar_pds = bar_base + 56;     // P, dS written to shared memory (256 arrivals)
  const uint32_t bar_mma2 = bar_base + 64;    // dK, dQ, dV products of this query block complete
  const uint32_t bar_dqfree = bar_base + 72;  // dQ read out of TMEM (128 arrivals: the dQ warps)
  const uint32_t bar_qfree = bar_base + 80;   // [2] dK_t complete: Q buffer t % 2 may be reloaded
  const uint32_t bar_dofree = bar_base + 96;  // [3] dV_t complete: dO buffer t % 3 may be reloaded
  const uint32_t bar_dq = bar_base + 120;     // dQ product of this query block complete
  const uint32_t bar_sdpfree = bar_base + 128; // S, dP of this step are in registers (256 arrivals): TMEM columns reusable
  const uint32_t tmem_ptr_smem = bar_base + 136;

  const int warp = threadIdx.x >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int lane = threadIdx.x & 31;
  const int jblk = blockIdx.x;
  const int head = blockIdx.y;
  const int tile = blockIdx.z;
  const int th = tile * args.heads + head;
  const int num_q = (args.seq + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    mbar_init(bar_kv, 1);
    for (uint32_t b = 0; b < kAbQBufs; ++b) { mbar_init(bar_q + 8 * b, 1); mbar_init(bar_qfree + 8 * b, 1); }
    for (uint32_t b = 0; b < kAbDoBufs; ++b) { mbar_init(bar_do + 8 * b, 1); mbar_init(bar_dofree + 8 * b, 1); }
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 256);
    mbar_init(bar_mma2, 1);
    mbar_init(bar_dq, 1);
    mbar_init(bar_dqfree, 128);
    mbar_init(bar_sdpfree, 256);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, kAbTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 336,
                 tdQ = tmem_base + 416;
  if (warp == 1) RV_ABTL(6, 0);   // CTA set up (barriers, TMEM)

  if (warp_u == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int k_row0 = th * args.seq_pad + jblk * 128;
      mbar_arrive_expect_tx(bar_kv, 2 * kAbTile);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        tma_load_2d(sK + c * 4096, &tmap_k, bar_kv, c * 16, k_row0);
        tma_load_2d(sV + c * 4096, &tmap_v, bar_kv, c * 16, k_row0);
      }
      for (int t = 0; t < num_q; ++t) {   // dO_t two blocks, Q_t one block ahead of the compute
        const int i = (jblk + t) % num_q; // rotated order: the key-block CTAs of a head work on different query blocks
        const uint32_t bd = static_cast<uint32_t>(t % kAbDoBufs), bq = static_cast<uint32_t>(t % kAbQBufs);
        const int q_row0 = th * args.seq_pad + i * 128;
        const int do_row0 = tile * args.seq + i * 128;
        if (t >= kAbDoBufs) mbar_wait(bar_dofree + 8 * bd, static_cast<uint32_t>((t / kAbDoBufs - 1) & 1));
        mbar_arrive_expect_tx(bar_do + 8 * bd, kAbTile);
#pragma unroll
        for (int c = 0; c < 5; ++c)
          tma_load_2d(sdO + bd * kAbTile + c * 4096, &tmap_do, bar_do + 8 * bd, head * args.hd + c * 16, do_row0);
        if (t >= kAbQBufs) mbar_wait(bar_qfree + 8 * bq, static_cast<uint32_t>((t / kAbQBufs - 1) & 1));
        mbar_arrive_expect_tx(bar_q + 8 * bq, kAbTile);
#pragma unroll
        for (int c = 0; c < 5; ++c) tma_load_2d(sQ + bq * kAbTile + c * 4096, &tmap_q, bar_q + 8 * bq, c * 16, q_row0);
      }
    }
  } else if (warp_u == 1) {
    // ===================== MMA issue (whole warp, elect.sync inside the asm blocks) =====================
    constexpr uint32_t kAMn = 1u << 15, kBMn = 1u << 16;
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);                  // S  = Q K^T      (K-major, K-major)
    constexpr uint32_t idesc_dp = make_idesc_bf16(128, 128);                 // dP = dO V^T     (K-major, K-major)
    constexpr uint32_t idesc_kv = make_idesc_bf16(128, 80) | kAMn | kBMn;    // dV = P^T dO, dK = dS^T Q
    constexpr uint32_t idesc_dq = make_idesc_bf16(128, 80) | kBMn;           // dQ = dS K
    const uint32_t tS_u = __shfl_sync(0xffffffffu, tS, 0);
    const uint32_t tdP_u = tS_u + 128, tdV_u = tS_u + 256, tdK_u = tS_u + 336, tdQ_u = tS_u + 416;
    // K-major views (contraction over head_dim): chunk c = K step c
    const uint64_t q_k = make_smem_desc(sQ, 256, kLayoutSw32), k_k = make_smem_desc(sK, 256, kLayoutSw32);
    const uint64_t do_k = make_smem_desc(sdO, 256, kLayoutSw32);
    // MN-major views (contraction over the 128 rows): 16-element head-dim groups 4096 B apart, 8 rows per 256 B atom
    const uint64_t q_mn = make_smem_desc_lbo(sQ, 4096, 256, kLayoutSw32);
    const uint64_t k_mn = make_smem_desc_lbo(sK, 4096, 256, kLayoutSw32);
    const uint64_t do_mn = make_smem_desc_lbo(sdO, 4096, 256, kLayoutSw32);
    const uint64_t v_k = make_smem_desc(sV, 256, kLayoutSw32);
    // P / dS tiles [128 q x 128 keys]: K-major A (contraction over keys) and MN-major A (contraction over queries)
    const uint64_t ds_k = make_smem_desc(sdS, 1024, kLayoutSw128);
    const uint64_t p_mn = make_smem_desc_lbo(sP, 16384, 1024, kLayoutSw128);
    const uint64_t ds_mn = make_smem_desc_lbo(sdS, 16384, 1024, kLayoutSw128);

    // S_{i+1} / dP_{i+1} are issued as soon as the softmax-gradient warps hold S_i / dP_i in registers (bar_sdpfree,
    // half-way through their work on block i): they complete while the second half of block i is still being computed,
    // so block i+1 starts without waiting for the tensor core; dQ_i / dV_i / dK_i then run under block i+1's first half.
    auto issue_sdp = [&](int i) {
      const uint32_t bq = static_cast<uint32_t>(i % kAbQBufs), bd = static_cast<uint32_t>(i % kAbDoBufs);
      mbar_wait(bar_q + 8 * bq, static_cast<uint32_t>((i / kAbQBufs) & 1));
      mbar_wait(bar_do + 8 * bd, static_cast<uint32_t>((i / kAbDoBufs) & 1));
      tc_fence_after();
      const uint64_t qoff = static_cast<uint64_t>(bq * (kAbTile >> 4)), dooff = static_cast<uint64_t>(bd * (kAbTile >> 4));
#pragma unroll
      for (int c = 0; c < 5; ++c)   // S: K step c = chunk c (4096 B)
        umma_bf16_ss_elect(tS_u, q_k + qoff + c * 256, k_k + c * 256, idesc_s, c != 0 ? 1u : 0u);
#pragma unroll
      for (int c = 0; c < 5; ++c)   // dP: K step c = chunk c
        umma_bf16_ss_elect(tdP_u, do_k + dooff + c * 256, v_k + c * 256, idesc_dp, c != 0 ? 1u : 0u);
      umma_commit_elect(bar_sdp);
    };
    mbar_wait(bar_kv, 0);
    RV_ABTL(6, 1);                // K / V landed
    issue_sdp(0);
    for (int i = 0; i < num_q; ++i) {   // i = step (the t of the producer / softmax loops): buffers and parities only
      const uint32_t par = static_cast<uint32_t>(i & 1);
      const uint32_t bq = static_cast<uint32_t>(i % kAbQBufs), bd = static_cast<uint32_t>(i % kAbDoBufs);
      const uint64_t qoff = static_cast<uint64_t>(bq * (kAbTile >> 4)), dooff = static_cast<uint64_t>(bd * (kAbTile >> 4));
      if (i + 1 < num_q) {
        mbar_wait(bar_sdpfree, par);                            // S_i / dP_i are in registers
        tc_fence_after();
        RV_ABTL(i, 0);
        issue_sdp(i + 1);
        RV_ABTL(i, 1);
      }
      mbar_wait(bar_pds, par);                                  // P, dS in shared memory
      tc_fence_after();
      RV_ABTL(i, 2);
#pragma unroll
      for (int s = 0; s < 8; ++s)   // dK += dS^T Q first: it is the last reader of Q_i, whose buffer is reloaded next
        umma_bf16_ss_elect(tdK_u, ds_mn + s * 128, q_mn + qoff + s * 32, idesc_kv, (i | s) != 0 ? 1u : 0u);
      umma_commit_elect(bar_qfree + 8 * bq);
      if (i > 0) {
        mbar_wait(bar_dqfree, par ^ 1u);                        // previous dQ read out
        tc_fence_after();
      }
#pragma unroll
      for (int s = 0; s < 8; ++s)   // dQ = dS K (fresh per step; read out and reduced by the dQ warps)
        umma_bf16_ss_elect(tdQ_u, ds_k + (s >> 2) * 1024 + (s & 3) * 2, k_mn + s * 32, idesc_dq, s != 0 ? 1u : 0u);
      umma_commit_elect(bar_dq);
#pragma unroll
      for (int s = 0; s < 8; ++s)   // dV += P^T dO: K step s = 16 query rows (P: 2048 B, dO chunks: 512 B)
        umma_bf16_ss_elect(tdV_u, p_mn + s * 128, do_mn + dooff + s * 32, idesc_kv, (i | s) != 0 ? 1u : 0u);
      umma_commit_elect(bar_dofree + 8 * bd);
      umma_commit_elect(bar_mma2);
      RV_ABTL(i, 3);
    }
  } else if (warp_u < 10) {
    // ===================== softmax-gradient warps: two threads per query row =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const uint32_t row_off = static_cast<uint32_t>(half) * 16384u + static_cast<uint32_t>(r) * 128u;
    const int D = args.heads * args.hd;

    // lse / delta of the first query block; the next block's pair is fetched one iteration ahead
    float lse_n = 0.f, delta_n = 0.f;
    {
      const int q0 = (jblk % num_q) * 128 + r;
      if (q0 < args.seq) {
        lse_n = __ldg(args.lse + static_cast<size_t>(th) * args.seq_pad + q0);
        delta_n = __ldg(args.delta + static_cast<size_t>(th) * args.seq_pad + q0);
      }
    }
    for (int t = 0; t < num_q; ++t) {
      const int i = (jblk + t) % num_q;           // query block of this step (rotated order, see the header)
      const uint32_t par = static_cast<uint32_t>(t & 1);
      const int qrow = i * 128 + r;
      const bool q_ok = qrow < args.seq;
      const float lse = lse_n, delta = delta_n;
      if (t + 1 < num_q) {
        const int qn = ((jblk + t + 1) % num_q) * 128 + r;
        if (qn < args.seq) {
          lse_n = __ldg(args.lse + static_cast<size_t>(th) * args.seq_pad + qn);
          delta_n = __ldg(args.delta + static_cast<size_t>(th) * args.seq_pad + qn);
        }
      }
      mbar_wait(bar_sdp, par);
      tc_fence_after();
      if (warp == 2) RV_ABTL(t, 4);
      const float nds = -delta * args.scale;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[32], dp[32];
        tmem_ld_x32(tS + lane_off + static_cast<uint32_t>(half * 64 + c * 32), s);
        tmem_ld_x32(tdP + lane_off + static_cast<uint32_t>(half * 64 + c * 32), dp);
        tmem_wait_ld();
        if (c == 1) {   // all of S_t / dP_t is in registers: the tensor core may overwrite them with block t+1
          tc_fence_before();
          mbar_arrive(bar_sdpfree);
          if (warp == 2) RV_ABTL(t, 5);
        }
        const int key0 = jblk * 128 + half * 64 + c * 32;
        uint32_t pk[16], dk[16];
        if (q_ok && key0 + 32 <= args.seq) {   // common case: nothing to mask
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = ex2_approx_bwd(fmaf(__uint_as_float(s[e]), args.scale_log2e, -lse));
            const float p1 = ex2_approx_bwd(fmaf(__uint_as_float(s[e + 1]), args.scale_log2e, -lse));
            pk[e >> 1] = pack_bf16x2(p0, p1);
            dk[e >> 1] = pack_bf16x2(p0 * fmaf(__uint_as_float(dp[e]), args.scale, nds),
                                     p1 * fmaf(__uint_as_float(dp[e + 1]), args.scale, nds));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float p0 = ex2_approx_bwd(fmaf(__uint_as_float(s[e]), args.scale_log2e, -lse));
            float p1 = ex2_approx_bwd(fmaf(__uint_as_float(s[e + 1]), args.scale_log2e, -lse));
            if (!q_ok || key0 + e >= args.seq) p0 = 0.f;
            if (!q_ok || key0 + e + 1 >= args.seq) p1 = 0.f;
            pk[e >> 1] = pack_bf16x2(p0, p1);
            dk[e >> 1] = pack_bf16x2(p0 * fmaf(__uint_as_float(dp[e]), args.scale, nds),
                                     p1 * fmaf(__uint_as_float(dp[e + 1]), args.scale, nds));
          }
        }
        // dV_{t-1} / dK_{t-1} read the previous P / dS: they run while this block's exponentials are computed
        if (c == 0 && t > 0) mbar_wait(bar_mma2, par ^ 1u);
        if (c == 0 && warp == 2) RV_ABTL(t, 6);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t chunk = static_cast<uint32_t>(4 * c + u) ^ sw;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + row_off + chunk * 16u), "r"(pk[4 * u]),
                       "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS + row_off + chunk * 16u), "r"(dk[4 * u]),
                       "r"(dk[4 * u + 1]), "r"(dk[4 * u + 2]), "r"(dk[4 * u + 3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pds);
      if (warp == 2) RV_ABTL(t, 7);

    }

    // ---- dK_j, dV_j -> bf16 -> dqkv[(tile*seq + key), D + head*hd + d] and [.., 2D + head*hd + d]
    //      (lanes are keys now)
    mbar_wait(bar_mma2, static_cast<uint32_t>((num_q - 1) & 1));
    tc_fence_after();
    const int key = jblk * 128 + r;
    uint32_t kk[40], vv[40];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      tmem_ld_x8(tdK + lane_off + static_cast<uint32_t>(half * 40 + c * 8), kk + c * 8);
      tmem_ld_x8(tdV + lane_off + static_cast<uint32_t>(half * 40 + c * 8), vv + c * 8);
    }
    tmem_wait_ld();
    if (key < args.seq) {
      __nv_bfloat16* base = args.dqkv + (static_cast<size_t>(tile) * args.seq + key) * (3 * D) + head * args.hd + half * 40;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (half * 40 + c * 8 < args.hd) {
          uint4 a, b;
          a.x = pack_bf16x2(__uint_as_float(kk[8 * c + 0]), __uint_as_float(kk[8 * c + 1]));
          a.y = pack_bf16x2(__uint_as_float(kk[8 * c + 2]), __uint_as_float(kk[8 * c + 3]));
          a.z = pack_bf16x2(__uint_as_float(kk[8 * c + 4]), __uint_as_float(kk[8 * c + 5]));
          a.w = pack_bf16x2(__uint_as_float(kk[8 * c + 6]), __uint_as_float(kk[8 * c + 7]));
          b.x = pack_bf16x2(__uint_as_float(vv[8 * c + 0]), __uint_as_float(vv[8 * c + 1]));
          b.y = pack_bf16x2(__uint_as_float(vv[8 * c + 2]), __uint_as_float(vv[8 * c + 3]));
          b.z = pack_bf16x2(__uint_as_float(vv[8 * c + 4]), __uint_as_float(vv[8 * c + 5]));
          b.w = pack_bf16x2(__uint_as_float(vv[8 * c + 6]), __uint_as_float(vv[8 * c + 7]));
          reinterpret_cast<uint4*>(base + D)[c] = a;
          reinterpret_cast<uint4*>(base + 2 * D)[c] = b;
        }
      }
    }
    if (warp == 2) RV_ABTL(6, 3);  // dK / dV stored
  }

  if (warp_u >= 10) {
    // ===================== dQ warps: one thread per query row =====================
    // dQ_i partial product of this key block: TMEM -> registers -> fp32 staging tile -> ONE bulk reduce-add into
    // dq_acc[th, i*128 .. +128, :] (contiguous, same row pitch as the staging tile).  Rows of invalid queries and the
    // columns in [hd, 80) hold zeros (their dS rows / K columns are zero), the 4 pad columns are zeroed once: all of it
    // lands in padding of the accumulator.
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t row_smem = sDQ + static_cast<uint32_t>(r & 63) * (kAbDqPitch * 4u);
    const bool dq_issuer = (threadIdx.x == 320);
    const int my_half = r >> 6;   // rows [0, 64) are staged and reduced first, then rows [64, 128), through the same tile
    if (my_half == 0) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(row_smem + 320u), "r"(0u) : "memory");
    for (int t = 0; t < num_q; ++t) {
      const int i = (jblk + t) % num_q;
      mbar_wait(bar_dq, static_cast<uint32_t>(t & 1));
      tc_fence_after();
      if (warp == 10) RV_ABTL(t, 8);
      uint32_t dq[80];
#pragma unroll
      for (int c = 0; c < 10; ++c) tmem_ld_x8(tdQ + lane_off + static_cast<uint32_t>(c * 8), dq + c * 8);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_dqfree);
      if (warp == 10) RV_ABTL(t, 9);
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        if (dq_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous reduction has read the tile
        named_bar_sync(1, 128);
        if (my_half == hh) {
#pragma unroll
          for (int e = 0; e < 20; ++e)   // rows 336 B apart: the 8 threads of a quarter warp hit 8 different bank groups
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_smem + static_cast<uint32_t>(e) * 16u),
                         "r"(dq[4 * e]), "r"(dq[4 * e + 1]), "r"(dq[4 * e + 2]), "r"(dq[4 * e + 3]) : "memory");
          fence_proxy_async_smem();
        }
        named_bar_sync(2, 128);
        if (dq_issuer) {
          float* dst = args.dq_acc + (static_cast<size_t>(th) * args.seq_pad + static_cast<size_t>(i) * 128 + hh * 64) * kAbDqPitch;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(sDQ),
                       "r"(static_cast<uint32_t>(kAbDqTile)) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (warp == 10) RV_ABTL(t, 10);
    }
    if (dq_issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all reductions performed before exit
    if (warp == 10) RV_ABTL(6, 2);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAbTmemCols);
  }
}

// delta[th, q] = sum_d dO[token, head*hd + d] * O[token, head*hd + d].  One warp per token: the row is read with
// coalesced 16-byte loads, every 16-byte piece belongs to one head (hd % 8 == 0), pieces are summed per head through
// shared memory.  heads * hd / 8 <= 8 * 32 pieces per row.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O, float* __restrict__ delta,
                  int tokens, int seq, int seq_pad, int heads, int hd) {
  __shared__ float part[8][256];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int token = blockIdx.x * 8 + wib;
  if (token >= tokens) return;
  const int per_head = hd / 8;
  const int pieces = heads * per_head;
  const uint4* a = reinterpret_cast<const uint4*>(dO + static_cast<size_t>(token) * heads * hd);
  const uint4* b = reinterpret_cast<const uint4*>(O + static_cast<size_t>(token) * heads * hd);
  for (int v = lane; v < pieces; v += 32) {
    const uint4 x = a[v], y = b[v];
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc += __uint_as_float(xw[j] << 16) * __uint_as_float(yw[j] << 16) +
             __uint_as_float(xw[j] & 0xFFFF0000u) * __uint_as_float(yw[j] & 0xFFFF0000u);
    part[wib][v] = acc;
  }
  __syncwarp();
  const int tile = token / seq, t = token - tile * seq;
  for (int h = lane; h < heads; h += 32) {
    float acc = 0.f;
    for (int j = 0; j < per_head; ++j) acc += part[wib][h * per_head + j];
    delta[(static_cast<size_t>(tile) * heads + h) * seq_pad + t] = acc;
  }
}

// dq_acc fp32 [th, seq_pad, kAbDqPitch] -> bf16 dqkv[(tile*seq + q), head*hd + d]
__global__ void __launch_bounds__(256)
attn_dq_store_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, int tokens, int seq,
                     int seq_pad, int heads, int hd) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one thread per (token, head, d/8)
  const int per_head = hd / 8;
  const size_t total = static_cast<size_t>(tokens) * heads * per_head;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % per_head);
  const int head = static_cast<int>((idx / per_head) % heads);
  const int token = static_cast<int>(idx / (static_cast<size_t>(per_head) * heads));
  const int tile = token / seq, t = token - tile * seq;
  const float* src = dq_acc + ((static_cast<size_t>(tile) * heads + head) * seq_pad + t) * kAbDqPitch + v * 8;
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
  *reinterpret_cast<uint4*>(dqkv + static_cast<size_t>(token) * (3 * heads * hd) + head * hd + v * 8) = o;
}

}  // namespace rv
