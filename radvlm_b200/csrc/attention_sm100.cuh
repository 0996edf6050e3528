// Non-causal multi-head attention for the SigLIP tower on sm_100a (729 tokens, head_dim 72).
//
// Replaces the materialised-score attention of SigLipAttention.forward
//   (finetuning/llava/model/multimodal_encoder/siglip_encoder.py:216-235:
//    matmul(q,k^T)*scale -> softmax(fp32) -> cast -> matmul(p,v) -> transpose/reshape)
// with one fused kernel.  Layout contract (produced by the QKV GEMM epilogue, gemm_sm100.cuh):
//    Q, K : [tiles*heads, seq_pad, hd_pad] bf16, zero padded (hd 72 -> 80, seq 729 -> 768)
//    Vt   : [tiles*heads, hd_pad, seq_pad] bf16  (V transposed so the PV B-operand is K-major)
//    out  : [tiles*seq, heads*hd] bf16 token-major (the A operand of out_proj)
//
// One CTA = 128 query rows of one (tile, head).  Warp roles: warp0 TMA producer, warp1 tcgen05.mma
// issuer, warps 2..5 softmax (one query row per thread).  S = Q K_j^T (128x128 fp32) and the running
// O (128x80 fp32) live in TMEM; P_j is written to shared memory as bf16 in the 128B-swizzled K-major
// layout and fed back as the A operand of O += P_j V_j.  Online softmax with lazy, deferred rescale: the
// reference maximum only moves when the row maximum grows by more than 2^8, so the TMEM O rescale
// (tcgen05.ld -> mul -> tcgen05.st) is rare.  Two CTAs are resident per SM (256 TMEM columns and
// ~93 KB shared memory each) so one CTA's softmax overlaps the other's MMAs.
#pragma once

#include "common.cuh"

namespace rv {

struct AttnArgs {
  __nv_bfloat16* out;  // [tiles*seq, heads*hd]
  int seq;             // 729
  int seq_pad;         // 768 (multiple of 128)
  int heads;           // 16
  int hd;              // 72
  float scale_log2e;   // hd^-0.5 * log2(e)
};

constexpr int kAttnBQ = 128;       // query rows per CTA
constexpr int kAttnBKV = 128;      // keys per inner step
constexpr int kAttnHdPad = 80;     // padded head dim (5 x UMMA_K)
constexpr int kAttnThreads = 192;
constexpr int kAttnQBytes = kAttnBQ * kAttnHdPad * 2;    // 20480 : 5 chunks of [128 x 32 B] (SW32)
constexpr int kAttnKBytes = kAttnBKV * kAttnHdPad * 2;   // 20480
constexpr int kAttnVBytes = kAttnHdPad * kAttnBKV * 2;   // 20480 : 2 atoms of [80 x 128 B] (SW128)
constexpr int kAttnPBytes = kAttnBQ * kAttnBKV * 2;      // 32768 : 2 atoms of [128 x 128 B] (SW128)
constexpr int kAttnSmemBytes = kAttnQBytes + kAttnKBytes + kAttnVBytes + kAttnPBytes + 1024 + 128;
constexpr int kAttnTmemCols = 256;  // S: [0,128)  O: [128,208)
constexpr float kAttnRescaleThreshold = 8.0f;

__global__ void __launch_bounds__(kAttnThreads, 2)
siglip_attention_kernel(const __grid_constant__ CUtensorMap tmap_q,
                        const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_vt, const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + kAttnQBytes;
  const uint32_t sV = sK + kAttnKBytes;
  const uint32_t sP = sV + kAttnVBytes;
  const uint32_t bar_base = sP + kAttnPBytes;
  const uint32_t bar_k = bar_base + 0;       // K_j (+Q for j == 0) landed
  const uint32_t bar_v = bar_base + 8;       // Vt_j landed
  const uint32_t bar_s = bar_base + 16;      // S_j complete in TMEM (also: K buffer free)
  const uint32_t bar_sfree = bar_base + 24;  // softmax has read S_j out of TMEM
  const uint32_t bar_p = bar_base + 32;      // P_j in smem, O rescaled
  const uint32_t bar_o = bar_base + 40;      // O += P_j V_j complete (P and V buffers free)
  const uint32_t tmem_ptr_smem = bar_base + 48;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qblk = blockIdx.x;
  const int head = blockIdx.y;
  const int tile = blockIdx.z;
  const int th = tile * args.heads + head;
  const int num_kv = args.seq_pad / kAttnBKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_sfree, 128);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, kAttnTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int q_row0 = th * args.seq_pad + qblk * kAttnBQ;
      for (int j = 0; j < num_kv; ++j) {
        const uint32_t par = static_cast<uint32_t>(j & 1);
        // K buffer is free once S_{j-1} has been computed
        if (j > 0) mbar_wait(bar_s, par ^ 1u);
        mbar_arrive_expect_tx(bar_k, j == 0 ? (kAttnQBytes + kAttnKBytes) : kAttnKBytes);
        if (j == 0) {
#pragma unroll
          for (int c = 0; c < 5; ++c) tma_load_2d(sQ + c * 4096, &tmap_q, bar_k, c * 16, q_row0);
        }
        const int k_row0 = th * args.seq_pad + j * kAttnBKV;
#pragma unroll
        for (int c = 0; c < 5; ++c) tma_load_2d(sK + c * 4096, &tmap_k, bar_k, c * 16, k_row0);
        // V buffer is free once PV_{j-1} has completed
        if (j > 0) mbar_wait(bar_o, par ^ 1u);
        mbar_arrive_expect_tx(bar_v, kAttnVBytes);
        tma_load_2d(sV, &tmap_vt, bar_v, j * kAttnBKV, th * kAttnHdPad);
        tma_load_2d(sV + 10240, &tmap_vt, bar_v, j * kAttnBKV + 64, th * kAttnHdPad);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kAttnBQ, kAttnBKV);
      constexpr uint32_t idesc_o = make_idesc_bf16(kAttnBQ, kAttnHdPad);
      auto issue_s = [&]() {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const uint64_t ad = make_smem_desc(sQ + c * 4096, 256, kLayoutSw32);
          const uint64_t bd = make_smem_desc(sK + c * 4096, 256, kLayoutSw32);
          umma_bf16_ss(tS, ad, bd, idesc_s, c != 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
      };
      mbar_wait(bar_k, 0);
      tc_fence_after();
      issue_s();
      for (int j = 0; j < num_kv; ++j) {
        const uint32_t par = static_cast<uint32_t>(j & 1);
        if (j + 1 < num_kv) {
          mbar_wait(bar_k, par ^ 1u);   // K_{j+1} landed
          mbar_wait(bar_sfree, par);    // softmax drained S_j
          tc_fence_after();
          issue_s();
        }
        mbar_wait(bar_p, par);  // P_j written, O rescaled
        mbar_wait(bar_v, par);  // Vt_j landed
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const uint64_t ad = make_smem_desc(sP + (s >> 2) * 16384 + (s & 3) * 32, 1024, kLayoutSw128);
          const uint64_t bd = make_smem_desc(sV + (s >> 2) * 10240 + (s & 3) * 32, 1024, kLayoutSw128);
          umma_bf16_ss(tO, ad, bd, idesc_o, (j | s) != 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
    }
  } else {
    // ===================== softmax / correction / output (4 warps, one row per thread) ==========
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // row within the query block
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    // Online softmax, single pass over each S block in 32-column chunks with the next tcgen05.ld in flight:
    // exponentials use the RUNNING reference m_ref (log2 domain); when a block's maximum exceeds it by more than
    // 2^8 the reference moves and the O / l rescale is deferred to the start of the next block (after that
    // block's PV has drained).  The result is exact: O and l always share one reference.
    float m_ref = 0.f;
    float l_sum = 0.f;
    float pend_alpha = 1.f;
    bool pend = false;
    const float sc = args.scale_log2e;
    const uint32_t p_row = sP + static_cast<uint32_t>(r) * 128u;
    const uint32_t sw = static_cast<uint32_t>(r & 7);

    for (int j = 0; j < num_kv; ++j) {
      const uint32_t par = static_cast<uint32_t>(j & 1);
      mbar_wait(bar_s, par);
      tc_fence_after();
      uint32_t sbuf[2][32];
      tmem_ld_x32(tS + lane_off, sbuf[0]);
      bool waited_o = (j == 0);
      if (__any_sync(0xffffffffu, pend)) {  // deferred rescale (rare): needs PV_{j-1} complete
        mbar_wait(bar_o, par ^ 1u);
        waited_o = true;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kAttnHdPad / 16; ++c) {
          uint32_t o[16];
          tmem_ld_x16(tO + lane_off + c * 16, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * pend_alpha);
          tmem_st_x16(tO + lane_off + c * 16, o);
        }
        tmem_wait_st();
        l_sum *= pend_alpha;
        pend_alpha = 1.f;
        pend = false;
      }
      const int valid = args.seq - j * kAttnBKV;  // keys >= valid are padding (last block only)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t* cur = sbuf[c & 1];
        tmem_wait_ld();
        if (c < 3) {
          tmem_ld_x32(tS + lane_off + 32 * (c + 1), sbuf[(c + 1) & 1]);
        } else {
          tc_fence_before();
          mbar_arrive(bar_sfree);  // S_j is entirely in registers: the MMA warp may overwrite it
        }
        if (valid < kAttnBKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) cur[i] = 0xFF800000u;  // -inf
        }
        if (j == 0 && c == 0) {  // initial reference: maximum of the first 32 keys
          float m0[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) m0[i] = __uint_as_float(cur[i]);
#pragma unroll
          for (int i = 4; i < 32; ++i) m0[i & 3] = fmaxf(m0[i & 3], __uint_as_float(cur[i]));
          m_ref = fmaxf(fmaxf(m0[0], m0[1]), fmaxf(m0[2], m0[3])) * sc;
        }
        const float neg_m = -m_ref;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float v0 = __uint_as_float(cur[i]), v1 = __uint_as_float(cur[i + 1]);
          mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(v0, v1));
          // p = 2^(s * scale*log2e - m_ref): one FFMA + one MUFU.EX2 (argument clamped against overflow)
          const float a0 = fminf(fmaf(v0, sc, neg_m), 100.f), a1 = fminf(fmaf(v1, sc, neg_m), 100.f);
          float p0, p1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(a0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(a1));
          sum4[(i >> 1) & 3] += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        if (c == 0 && !waited_o) mbar_wait(bar_o, par ^ 1u);  // PV_{j-1} done: the P buffer may be overwritten
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t unit = static_cast<uint32_t>((c & 1) * 4 + u) ^ sw;
          const uint32_t addr = p_row + static_cast<uint32_t>(c >> 1) * 16384u + unit * 16u;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * u + 0]),
                       "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                       : "memory");
        }
      }
      l_sum += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sc;
      if (mx > m_ref + kAttnRescaleThreshold) {
        pend_alpha = exp2f(m_ref - mx);
        m_ref = mx;
        pend = true;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }

    // ---- final: O / l -> bf16 -> out[(tile*seq + t), head*hd + d]
    mbar_wait(bar_o, static_cast<uint32_t>((num_kv - 1) & 1));
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
    const int t = qblk * kAttnBQ + r;
    uint32_t o[kAttnHdPad];
#pragma unroll
    for (int c = 0; c < kAttnHdPad / 16; ++c) tmem_ld_x16(tO + lane_off + c * 16, o + c * 16);
    tmem_wait_ld();
    if (t < args.seq) {
      __nv_bfloat16* dst = args.out +
                           (static_cast<size_t>(tile) * args.seq + t) * (args.heads * args.hd) +
                           head * args.hd;
      const int nvec = args.hd / 8;
#pragma unroll
      for (int c = 0; c < kAttnHdPad / 8; ++c) {
        if (c < nvec) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv_l, __uint_as_float(o[8 * c + 1]) * inv_l);
          pk.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv_l, __uint_as_float(o[8 * c + 3]) * inv_l);
          pk.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv_l, __uint_as_float(o[8 * c + 5]) * inv_l);
          pk.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv_l, __uint_as_float(o[8 * c + 7]) * inv_l);
          reinterpret_cast<uint4*>(dst)[c] = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}

}  // namespace rv
