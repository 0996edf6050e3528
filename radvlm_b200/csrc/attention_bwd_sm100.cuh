// Backward of the SigLIP attention (siglip_encoder.py:216-235 under autograd) on sm_100a tensor cores.
//
//   P  = softmax(Q K^T * scale)            recomputed from the forward's log-sum-exp:  P = 2^(S * scale * log2e - lse)
//   dV = P^T dO        dP = dO V^T         dS = P o (dP - delta) * scale,  delta = rowsum(dO o O)
//   dQ = dS K          dK = dS^T Q
//
// One CTA owns one 128-key block j of one (tile, head): dK_j and dV_j accumulate in TMEM over the query blocks
// i = 0..5, dQ_i partial products are added to an fp32 buffer with atomics (every key block contributes to every
// query block).  All five products run on tcgen05; the operands are read in place, without transposed copies:
//   * Q_i, K_j, dO_i sit in shared memory as five [128 rows x 32 B] 32B-swizzled chunks (dO_i is fetched by TMA
//     straight out of the token-major [tokens, heads*hd] gradient of the attention output).  The same bytes are a
//     K-major operand when the contraction runs over head_dim (S, dP) and an MN-major operand when it runs over
//     the rows (dV = P^T dO, dK = dS^T Q, dQ = dS K).
//   * V_j is loaded like K_j (five chunks) and is a plain K-major B operand of dP = dO V^T.
//   * P and dS are written by the softmax warps as bf16 [q][key] tiles (128B swizzle): K-major A for dQ = dS K,
//     MN-major A for P^T dO and dS^T Q.
// Warp roles: warp0 TMA, warp1 MMA issue (whole warp, elect.sync), warps 2..9 two threads per query row.
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
  float* dq_acc;         // [tiles*heads, seq_pad, 80] fp32, zero-initialised
  __nv_bfloat16* dqkv;   // [tiles*seq, 3*heads*hd]: this kernel writes the dK and dV column blocks
  int seq, seq_pad, heads, hd;
  float scale, scale_log2e;
};

constexpr int kAbThreads = 320;
constexpr int kAbTile = 128 * 80 * 2;   // 20480: Q / K / V / dO tiles (5 SW32 chunks each)
constexpr int kAbPTile = 128 * 128 * 2; // 32768: P and dS
constexpr int kAbSmemBytes = 6 * kAbTile + 2 * kAbPTile + 128;  // K, V, 2 x (Q, dO), P, dS
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
  const uint32_t sQ = sV + kAbTile;          // two buffers: Q_i in sQ + (i & 1) * 2 * kAbTile
  const uint32_t sdO = sQ + kAbTile;         // ... dO_i right behind its Q_i
  const uint32_t sP = sQ + 4 * kAbTile;
  const uint32_t sdS = sP + kAbPTile;
  const uint32_t bar_base = sdS + kAbPTile;
  const uint32_t bar_kv = bar_base + 0;       // K_j, Vt_j landed
  const uint32_t bar_qdo = bar_base + 8;      // [2] Q_i, dO_i landed in buffer i & 1
  const uint32_t bar_sdp = bar_base + 24;     // S, dP complete in TMEM
  const uint32_t bar_pds = bar_base + 32;     // P, dS written to shared memory (256 arrivals)
  const uint32_t bar_mma2 = bar_base + 40;    // dV, dK, dQ products of this query block complete
  const uint32_t bar_dqfree = bar_base + 48;  // dQ read out of TMEM (256 arrivals)
  const uint32_t bar_qdofree = bar_base + 56; // [2] the products reading Q / dO buffer i & 1 are complete
  const uint32_t bar_dq = bar_base + 72;      // dQ product of this query block complete (read out while dV / dK run)
  const uint32_t tmem_ptr_smem = bar_base + 80;

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
    mbar_init(bar_qdo, 1);
    mbar_init(bar_qdo + 8, 1);
    mbar_init(bar_qdofree, 1);
    mbar_init(bar_qdofree + 8, 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 256);
    mbar_init(bar_mma2, 1);
    mbar_init(bar_dq, 1);
    mbar_init(bar_dqfree, 256);
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
      for (int i = 0; i < num_q; ++i) {   // Q_i / dO_i -> buffer i & 1, one query block ahead of the compute
        const uint32_t b = static_cast<uint32_t>(i & 1);
        if (i >= 2) mbar_wait(bar_qdofree + 8 * b, static_cast<uint32_t>(((i >> 1) - 1) & 1));
        mbar_arrive_expect_tx(bar_qdo + 8 * b, 2 * kAbTile);
        const int q_row0 = th * args.seq_pad + i * 128;
        const int do_row0 = tile * args.seq + i * 128;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          tma_load_2d(sQ + b * 2 * kAbTile + c * 4096, &tmap_q, bar_qdo + 8 * b, c * 16, q_row0);
          tma_load_2d(sdO + b * 2 * kAbTile + c * 4096, &tmap_do, bar_qdo + 8 * b, head * args.hd + c * 16, do_row0);
        }
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

    // S_{i+1} / dP_{i+1} are issued right behind the dV / dK / dQ products of block i (their TMEM columns are free once
    // the softmax-gradient warps have published P_i / dS_i), so they run while block i's dQ is read out.
    auto issue_sdp = [&](int i) {
      const uint32_t b = static_cast<uint32_t>(i & 1);
      mbar_wait(bar_qdo + 8 * b, static_cast<uint32_t>((i >> 1) & 1));
      tc_fence_after();
      const uint64_t qoff = static_cast<uint64_t>(b * ((2 * kAbTile) >> 4));
#pragma unroll
      for (int c = 0; c < 5; ++c)   // S: K step c = chunk c (4096 B)
        umma_bf16_ss_elect(tS_u, q_k + qoff + c * 256, k_k + c * 256, idesc_s, c != 0 ? 1u : 0u);
#pragma unroll
      for (int c = 0; c < 5; ++c)   // dP: K step c = chunk c
        umma_bf16_ss_elect(tdP_u, do_k + qoff + c * 256, v_k + c * 256, idesc_dp, c != 0 ? 1u : 0u);
      umma_commit_elect(bar_sdp);
    };
    mbar_wait(bar_kv, 0);
    issue_sdp(0);
    for (int i = 0; i < num_q; ++i) {
      const uint32_t par = static_cast<uint32_t>(i & 1);
      const uint64_t qoff = static_cast<uint64_t>(par * ((2 * kAbTile) >> 4));
      mbar_wait(bar_pds, par);                                  // P, dS in shared memory; S / dP consumed
      if (i > 0) mbar_wait(bar_dqfree, par ^ 1u);               // previous dQ read out
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < 8; ++s)   // dQ = dS K first: its read-out (atomics) then overlaps the dV / dK products
        umma_bf16_ss_elect(tdQ_u, ds_k + (s >> 2) * 1024 + (s & 3) * 2, k_mn + s * 32, idesc_dq, s != 0 ? 1u : 0u);
      umma_commit_elect(bar_dq);
#pragma unroll
      for (int s = 0; s < 8; ++s)   // dV += P^T dO: K step s = 16 query rows (P: 2048 B, dO chunks: 512 B)
        umma_bf16_ss_elect(tdV_u, p_mn + s * 128, do_mn + qoff + s * 32, idesc_kv, (i | s) != 0 ? 1u : 0u);
#pragma unroll
      for (int s = 0; s < 8; ++s)   // dK += dS^T Q
        umma_bf16_ss_elect(tdK_u, ds_mn + s * 128, q_mn + qoff + s * 32, idesc_kv, (i | s) != 0 ? 1u : 0u);
      umma_commit_elect(bar_mma2);
      umma_commit_elect(bar_qdofree + 8 * par);
      if (i + 1 < num_q) issue_sdp(i + 1);
    }
  } else {
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
    if (r < args.seq) {
      lse_n = __ldg(args.lse + static_cast<size_t>(th) * args.seq_pad + r);
      delta_n = __ldg(args.delta + static_cast<size_t>(th) * args.seq_pad + r);
    }
    for (int i = 0; i < num_q; ++i) {
      const uint32_t par = static_cast<uint32_t>(i & 1);
      const int qrow = i * 128 + r;
      const bool q_ok = qrow < args.seq;
      const float lse = lse_n, delta = delta_n;
      if (i + 1 < num_q && qrow + 128 < args.seq) {
        lse_n = __ldg(args.lse + static_cast<size_t>(th) * args.seq_pad + qrow + 128);
        delta_n = __ldg(args.delta + static_cast<size_t>(th) * args.seq_pad + qrow + 128);
      }
      mbar_wait(bar_sdp, par);
      tc_fence_after();
      if (i > 0) mbar_wait(bar_mma2, par ^ 1u);  // the products that read the previous P / dS are complete
      const float nds = -delta * args.scale;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[32], dp[32];
        tmem_ld_x32(tS + lane_off + static_cast<uint32_t>(half * 64 + c * 32), s);
        tmem_ld_x32(tdP + lane_off + static_cast<uint32_t>(half * 64 + c * 32), dp);
        tmem_wait_ld();
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

      // ---- dQ_i partial product of this key block -> fp32 atomics (columns [40*half, +40) < hd)
      mbar_wait(bar_dq, par);
      tc_fence_after();
      uint32_t dq[40];
#pragma unroll
      for (int c = 0; c < 5; ++c) tmem_ld_x8(tdQ + lane_off + static_cast<uint32_t>(half * 40 + c * 8), dq + c * 8);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_dqfree);
#ifndef RV_ABWD_NO_ATOMICS
      if (q_ok) {
        // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the L2 atomic operations of scalar adds.
        // Columns >= hd are padding of the accumulator row: adding to them is harmless and keeps the vectors whole.
        float* dst = args.dq_acc + (static_cast<size_t>(th) * args.seq_pad + qrow) * 80 + half * 40;
#pragma unroll
        for (int e = 0; e < 40; e += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + e), "f"(__uint_as_float(dq[e])),
                       "f"(__uint_as_float(dq[e + 1])), "f"(__uint_as_float(dq[e + 2])), "f"(__uint_as_float(dq[e + 3]))
                       : "memory");
      }
#endif
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

// dq_acc fp32 [th, seq_pad, 80] -> bf16 dqkv[(tile*seq + q), head*hd + d]
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
  const float* src = dq_acc + ((static_cast<size_t>(tile) * heads + head) * seq_pad + t) * 80 + v * 8;
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
  *reinterpret_cast<uint4*>(dqkv + static_cast<size_t>(token) * (3 * heads * hd) + head * hd + v * 8) = o;
}

}  // namespace rv
