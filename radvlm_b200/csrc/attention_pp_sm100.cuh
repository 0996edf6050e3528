// Non-causal multi-head attention for the SigLIP tower on sm_100a (729 tokens, head_dim 72): the "ping-pong" kernel.
//
// Same contract as attention_sm100.cuh (replaces SigLipAttention.forward,
// finetuning/llava/model/multimodal_encoder/siglip_encoder.py:216-235; Q / K / V layouts produced by the QKV GEMM
// epilogue, V with a ones column at `hd` so the tensor core accumulates the softmax row sums).
//
// Why a second organisation.  The timeline of the two-CTAs-per-SM kernel (tools/attn_timeline.cu) showed the
// softmax warps of both resident CTAs drifting into phase: all of them burst their MUFU.EX2 instructions at the same
// time (the XU pipe takes 8 cycles per warp instruction, 16 ex2 / clk / SM) and then all of them sit in the
// MUFU-free part of a block (TMEM load, row maximum, exchange, scaling, P store) together, so the XU pipe idled for
// more than half of the time (ncu: XU 49 %, issue 46 %, tensor 32 %) and nothing was saturated.  Processor sharing
// keeps two such loops locked in phase; only an explicit hand-over un-locks them.
//
// Organisation.  ONE CTA per SM owns 256 query rows of a (tile, head): two groups of 128 rows (A, B), each with its
// own S (96 columns) / P (48) / O (80) TMEM regions (2 x 256 = 512) and four softmax warps, ONE thread per query
// row (96 score columns in registers: exact row maximum without any exchange).  The groups take turns on the XU
// pipe through two named barriers: a group does its MUFU-free work (S load, maximum, lazy-rescale decision, scale
// FFMAs, the polynomial share of the exponentials), waits for its turn, issues its MUFU burst, hands the turn to the
// other group, then packs and stores P.  The MMA warp issues S_A, S_B one key block ahead and O_A += P_A V,
// O_B += P_B V as the P tiles are published; K / V blocks are loaded once for 256 query rows.
#pragma once

#include <type_traits>

#include "attention_sm100.cuh"

namespace rv {

constexpr int kPpGroups = 2;
constexpr int kPpItemRows = kPpGroups * kAttnBQ;             // 256 query rows per work item
constexpr int kPpSoftmaxWarps = 4 * kPpGroups;
constexpr int kPpGroupThreads = 128;
constexpr int kPpEpilogueWarps = 4;                          // read O out of TMEM and store it, off the softmax path
constexpr int kPpEpilogueThreads = kPpEpilogueWarps * 32;
constexpr int kPpFirstSoftmaxWarp = 4;                       // warpgroup 0: TMA producer, MMA issuer, two idle warps
constexpr int kPpFirstEpilogueWarp = kPpFirstSoftmaxWarp + kPpSoftmaxWarps;
constexpr int kPpThreads = (kPpFirstEpilogueWarp + kPpEpilogueWarps) * 32;  // 512 = 4 warpgroups
// Register budget per warpgroup (setmaxnreg; 512 threads start at 128): 56 + 2 x 160 + 120 = 496 <= 512
constexpr int kPpRegsIo = 56, kPpRegsSoftmax = 160, kPpRegsEpilogue = 120;
constexpr int kPpStages = 4;                                 // K / V ring depth
constexpr int kPpQBufs = 2;                                  // Q of the next item is prefetched during the current one
constexpr int kPpMrefBytes = 2 * kPpGroups * kAttnBQ * 4;           // reference maxima handed to the epilogue warps (lse)
constexpr int kPpOutBytes = kAttnBQ * (kAttnHdPad - 8) * 2;   // 18432 : [128 rows][hd <= 72] bf16 staging tile of the TMA store
constexpr int kPpSmemBytes = kPpQBufs * kPpGroups * kAttnQBytes + kPpStages * (kAttnKBytes + kAttnVBytes) + kPpOutBytes +
                             kPpMrefBytes + 512;
static_assert(kPpSmemBytes <= 227 * 1024, "shared memory budget");
constexpr int kPpTmemCols = 512;
constexpr int kPpTmemGroup = 256;  // per group: S [0,96)  P (bf16 pairs) [96,144)  O [160,240) (32-column aligned)
constexpr int kPpTmemP = kAttnBKV, kPpTmemO = 160;
static_assert(kPpTmemP + kAttnBKV / 2 <= kPpTmemO, "P must not overlap O");
static_assert(kPpTmemO + kAttnHdPad <= kPpTmemGroup && kPpGroups * kPpTmemGroup <= kPpTmemCols, "TMEM budget");
// Exponentials per 16 computed on the FMA pipe (degree-3 polynomial) instead of MUFU.EX2.
#ifndef RV_PP_POLY_PER_16
#define RV_PP_POLY_PER_16 0
#endif

__device__ __forceinline__ float ex2_approx_v(float x) {  // volatile: stays behind the turn barrier
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16x2_v(float lo, float hi) {  // volatile: keeps its place in the burst
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// Turn barriers (hardware named barriers 1 and 2, 128 waiting + 128 arriving threads).
template <int kId>
__device__ __forceinline__ void turn_wait(uint32_t zero) {  // `zero` ties the wait behind its producers
  asm volatile("bar.sync %0, %1;" ::"n"(kId), "r"(2 * kPpGroupThreads + zero) : "memory");
}
template <int kId>
__device__ __forceinline__ void turn_pass(uint32_t zero) {  // `zero` ties the arrival behind its producers
  asm volatile("bar.arrive %0, %1;" ::"n"(kId), "r"(2 * kPpGroupThreads + zero) : "memory");
}

#ifdef RV_ATTN_TIMELINE
#define RV_PPTL(slot)                                                                                                \
  do {                                                                                                                 \
    if (args.lse != nullptr && it == RV_ATTN_TIMELINE && lane == 0 && quad == 2 && (slot) < 32)                         \
      reinterpret_cast<long long*>(args.lse)[static_cast<size_t>(blockIdx.x) * 64 + grp * 32 + (slot)] = clock64();   \
  } while (0)
#else
#define RV_PPTL(slot) do { } while (0)
#endif
#define RV_PPTL_J(slot) do { if (j < 3) RV_PPTL(slot); } while (0)
#ifdef RV_ATTN_TIMELINE
#define RV_PPTL_NEXT()                                                                                               \
  do {                                                                                                                 \
    if (args.lse != nullptr && it == RV_ATTN_TIMELINE + 1 && j < 2 && lane == 0 && quad == 2)                          \
      reinterpret_cast<long long*>(args.lse)[static_cast<size_t>(blockIdx.x) * 64 + grp * 32 + 28 + j] = clock64();   \
  } while (0)
#else
#define RV_PPTL_NEXT() do { } while (0)
#endif

// Work item w = (tile * heads + head) * num_qblk + qblk with 256-row query blocks; CTA c processes w = c,
// c + gridDim.x, ...  `g` counts key blocks over all of a CTA's items: rings and barriers run across items.
__global__ void __launch_bounds__(kPpThreads, 1)
siglip_attention_pp_kernel(const __grid_constant__ CUtensorMap tmap_q,    // Q  columns [0,64)  : SW128 box {64, 128}
                           const __grid_constant__ CUtensorMap tmap_q2,   // Q  columns [64,80) : SW32  box {16, 128}
                           const __grid_constant__ CUtensorMap tmap_k,    // K  columns [0,64)  : SW128 box {64, 96}
                           const __grid_constant__ CUtensorMap tmap_k2,   // K  columns [64,80) : SW32  box {16, 96}
                           const __grid_constant__ CUtensorMap tmap_v,    // V  16-column chunks : SW32  box {16, 96}
                           const __grid_constant__ CUtensorMap tmap_out,  // out [tiles][seq][heads*hd] : box {hd, 128, 1}
                           const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();  // swizzled operand tiles need 1024-byte alignment
  const uint32_t sQ = smem_base;                          // [item parity][group][Q block]
  const uint32_t sK = sQ + kPpQBufs * kPpGroups * kAttnQBytes;  // ring
  const uint32_t sV = sK + kPpStages * kAttnKBytes;       // ring
  const uint32_t sO = sV + kPpStages * kAttnVBytes;       // output staging tile (epilogue warps)
  const uint32_t sM = sO + kPpOutBytes;       // [item parity][group][row] fp32 reference maxima of a finished item
  const uint32_t bar_base = sM + kPpMrefBytes;
  const uint32_t bar_k = bar_base + 0;        // [stages] K_g landed
  const uint32_t bar_v = bar_base + 32;       // [stages] V_g landed
  const uint32_t bar_kfree = bar_base + 64;   // [stages] S_A,g and S_B,g complete
  const uint32_t bar_vfree = bar_base + 96;   // [stages] PV_A,g and PV_B,g complete
  const uint32_t bar_s = bar_base + 128;      // [group] S_g complete in TMEM
  const uint32_t bar_sfree = bar_base + 144;  // [group] S_g is in registers
  const uint32_t bar_p = bar_base + 160;      // [group] P_g in TMEM, O rescaled
  const uint32_t bar_o = bar_base + 176;      // [group] O += P_g V_g complete
  const uint32_t bar_ofree = bar_base + 192;  // [group] the item's O is in registers (epilogue warps)
  const uint32_t bar_q = bar_base + 208;      // [item parity] Q (both groups) of item `it` landed
  const uint32_t bar_qfree = bar_base + 224;  // [item parity] last S of the item complete: Q may be overwritten
  const uint32_t bar_ofin = bar_base + 240;   // [group] last PV of the item complete: O is final
  const uint32_t tmem_ptr_smem = bar_base + 256;
  const uint32_t zero_smem = bar_base + 264;  // a zero word (see the turn barriers)
  static_assert(kPpStages <= 4, "barrier layout");

  const int warp = threadIdx.x >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int lane = threadIdx.x & 31;
  const int num_kv = (args.seq + kAttnBKV - 1) / kAttnBKV;
  const int num_items = (args.total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_q2);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_k2);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_out);
    for (uint32_t i = 0; i < kPpStages; ++i) {
      mbar_init(bar_k + 8 * i, 1);
      mbar_init(bar_v + 8 * i, 1);
      mbar_init(bar_kfree + 8 * i, 1);
      mbar_init(bar_vfree + 8 * i, 1);
    }
    for (uint32_t i = 0; i < kPpGroups; ++i) {
      mbar_init(bar_s + 8 * i, 1);
      mbar_init(bar_sfree + 8 * i, kPpGroupThreads);
      mbar_init(bar_p + 8 * i, kPpGroupThreads);
      mbar_init(bar_o + 8 * i, 1);
      mbar_init(bar_ofree + 8 * i, kPpEpilogueThreads);
      mbar_init(bar_ofin + 8 * i, 1);
    }
    for (uint32_t i = 0; i < kPpQBufs; ++i) {
      mbar_init(bar_q + 8 * i, 1);
      mbar_init(bar_qfree + 8 * i, 1);
    }
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(zero_smem), "r"(0u) : "memory");
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, kPpTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  // PDL: the prologue above ran under the previous kernel's tail; from here on global memory is touched
  griddep_wait();
  griddep_launch_dependents();

  if (warp_u < kPpFirstSoftmaxWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPpRegsIo));
  }
  if (warp_u == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int g = 0;
      uint32_t slot = 0, ring_par = 0;
      for (int it = 0; it < num_items; ++it) {
        const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        const int th = w / args.num_qblk, qblk = w - th * args.num_qblk;
        const uint32_t qb = static_cast<uint32_t>(it & 1), quse = static_cast<uint32_t>(it >> 1);
        if (quse > 0) mbar_wait(bar_qfree + 8 * qb, (quse - 1) & 1u);  // item it-2's last S complete
        const uint32_t bq = bar_q + 8 * qb, sQi = sQ + qb * (kPpGroups * kAttnQBytes);
        mbar_arrive_expect_tx(bq, kPpGroups * kAttnQBytes);
        const int q_row0 = th * args.seq_pad + qblk * kPpItemRows;
#pragma unroll
        for (int gq = 0; gq < kPpGroups; ++gq) {
          tma_load_2d(sQi + gq * kAttnQBytes, &tmap_q, bq, 0, q_row0 + gq * kAttnBQ);
          tma_load_2d(sQi + gq * kAttnQBytes + kAttnQ2Off, &tmap_q2, bq, 64, q_row0 + gq * kAttnBQ);
        }
        for (int j = 0; j < num_kv; ++j, ++g) {
          if (g >= kPpStages) mbar_wait(bar_kfree + 8 * slot, ring_par ^ 1u);
          mbar_arrive_expect_tx(bar_k + 8 * slot, kAttnKBytes);
          const int k_row0 = th * args.seq_pad + j * kAttnBKV;
          tma_load_2d(sK + slot * kAttnKBytes, &tmap_k, bar_k + 8 * slot, 0, k_row0);
          tma_load_2d(sK + slot * kAttnKBytes + kAttnK2Off, &tmap_k2, bar_k + 8 * slot, 64, k_row0);
          if (g >= kPpStages) mbar_wait(bar_vfree + 8 * slot, ring_par ^ 1u);
          mbar_arrive_expect_tx(bar_v + 8 * slot, kAttnVBytes);
#pragma unroll
          for (int c = 0; c < 5; ++c)
            tma_load_2d(sV + slot * kAttnVBytes + c * kAttnVChunk, &tmap_v, bar_v + 8 * slot, c * 16, k_row0);
          if (++slot == kPpStages) { slot = 0; ring_par ^= 1u; }
        }
      }
    }
  } else if (warp_u == 1) {
    // ===================== MMA issuer (whole warp converged, elect.sync inside the asm blocks) =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(kAttnBQ, kAttnBKV);
    constexpr uint32_t idesc_s64 = make_idesc_bf16(kAttnBQ, 64);
    constexpr uint32_t idesc_o = make_idesc_bf16(kAttnBQ, kAttnHdPad) | (1u << 16);  // B (= V) is MN-major
    // The last key block holds seq - (num_kv - 1) * 96 valid keys (57 for 729 tokens): when they fit 64 columns its S is
    // computed 64 wide and its P V product 64 deep - the other 32 columns are padding (4 % of all score columns)
    const bool short_last = args.trim_last && (args.seq - (num_kv - 1) * kAttnBKV) <= 64;
    const uint32_t tbase_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t qd128 = make_smem_desc(sQ, 1024, kLayoutSw128);
    const uint64_t qd32 = make_smem_desc(sQ + kAttnQ2Off, 256, kLayoutSw32);
    const uint64_t kd128 = make_smem_desc(sK, 1024, kLayoutSw128);
    const uint64_t kd32 = make_smem_desc(sK + kAttnK2Off, 256, kLayoutSw32);
    const uint64_t vd = make_smem_desc_lbo(sV, kAttnVChunk, 256, kLayoutSw32);
    const int total_blocks = num_items * num_kv;
    uint32_t s_slot = 0, s_par = 0;  // ring position of the S issue (one block ahead of the PV issue)
    // S_{grp,g} = Q_grp K_g^T
    auto issue_s = [&](int grp, int g, int it, int j) {
      const uint32_t qb = static_cast<uint32_t>(it & 1);
      if (j == 0) mbar_wait(bar_q + 8 * qb, static_cast<uint32_t>(it >> 1) & 1u);
      mbar_wait(bar_k + 8 * s_slot, s_par);
      if (g > 0) mbar_wait(bar_sfree + 8 * grp, static_cast<uint32_t>((g - 1) & 1));
      tc_fence_after();
      const uint32_t tS = tbase_u + static_cast<uint32_t>(grp * kPpTmemGroup);
      const uint64_t qoff = static_cast<uint64_t>((qb * kPpGroups + grp) * (kAttnQBytes >> 4));
      const uint64_t koff = static_cast<uint64_t>(s_slot * (kAttnKBytes >> 4));
      const uint32_t ids = (short_last && j == num_kv - 1) ? idesc_s64 : idesc_s;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        umma_bf16_ss_elect(tS, qd128 + qoff + 2 * c, kd128 + koff + 2 * c, ids, c != 0 ? 1u : 0u);
      umma_bf16_ss_elect(tS, qd32 + qoff, kd32 + koff, ids, 1u);
      umma_commit_elect(bar_s + 8 * grp);
      if (grp == kPpGroups - 1) {
        umma_commit_elect(bar_kfree + 8 * s_slot);
        if (j == num_kv - 1) umma_commit_elect(bar_qfree + 8 * qb);
        if (++s_slot == kPpStages) { s_slot = 0; s_par ^= 1u; }
      }
    };
    if (total_blocks > 0) {
      issue_s(0, 0, 0, 0);
      issue_s(1, 0, 0, 0);
    }
    int g = 0;
    uint32_t slot = 0, ring_par = 0;
    for (int it = 0; it < num_items; ++it) {
      for (int j = 0; j < num_kv; ++j, ++g) {
        const bool wrap = (j + 1 == num_kv);
#pragma unroll
        for (int grp = 0; grp < kPpGroups; ++grp) {
          if (g + 1 < total_blocks) issue_s(grp, g + 1, wrap ? it + 1 : it, wrap ? 0 : j + 1);
          mbar_wait(bar_p + 8 * grp, static_cast<uint32_t>(g & 1));
          mbar_wait(bar_v + 8 * slot, ring_par);
          if (j == 0 && it > 0) mbar_wait(bar_ofree + 8 * grp, static_cast<uint32_t>((it - 1) & 1));
          tc_fence_after();
          const uint32_t tP = tbase_u + static_cast<uint32_t>(grp * kPpTmemGroup + kPpTmemP);
          const uint32_t tO = tbase_u + static_cast<uint32_t>(grp * kPpTmemGroup + kPpTmemO);
          const uint64_t voff = static_cast<uint64_t>(slot * (kAttnVBytes >> 4));
          const int pv_steps = (short_last && wrap) ? 4 : kAttnBKV / 16;
#pragma unroll
          for (int s = 0; s < kAttnBKV / 16; ++s)
            if (s < pv_steps)
              umma_bf16_ts_elect(tO, tP + static_cast<uint32_t>(s * 8), vd + voff + 32 * s, idesc_o, (j | s) != 0 ? 1u : 0u);
          umma_commit_elect(bar_o + 8 * grp);
          if (wrap) umma_commit_elect(bar_ofin + 8 * grp);
          if (grp == kPpGroups - 1) umma_commit_elect(bar_vfree + 8 * slot);
        }
        if (++slot == kPpStages) { slot = 0; ring_par ^= 1u; }
      }
    }
  } else if (warp_u >= kPpFirstSoftmaxWarp && warp_u < kPpFirstEpilogueWarp) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPpRegsSoftmax));
    // ===================== softmax / correction: 2 groups x 4 warps, one thread per query row ==========
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int grp = (warp - kPpFirstSoftmaxWarp) >> 2;
    const int r = quad * 32 + lane;          // row within the group's query block
    const uint32_t tG = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(grp * kPpTmemGroup);
    const uint32_t tS = tG, tP = tG + kPpTmemP, tO = tG + kPpTmemO;
    const uint32_t b_s = bar_s + 8 * grp, b_sfree = bar_sfree + 8 * grp, b_p = bar_p + 8 * grp;
    const uint32_t b_o = bar_o + 8 * grp;
    const float sc = args.scale_log2e;
    const int total_blocks = num_items * num_kv;
    const bool short_last = args.trim_last && (args.seq - (num_kv - 1) * kAttnBKV) <= 64;  // see the MMA issuer

    if (grp == 1 && total_blocks > 0) turn_pass<1>(0u);  // group A owns the first turn

    int g = 0;
    for (int it = 0; it < num_items; ++it) {
      float m_ref = -INFINITY;  // reference maximum (scaled, log2 domain)
      for (int j = 0; j < num_kv; ++j, ++g) {
        mbar_wait(b_s, static_cast<uint32_t>(g & 1));
        tc_fence_after();
        RV_PPTL_J(8 * j + 0);
        RV_PPTL_NEXT();
        uint32_t s[kAttnBKV];
        tmem_ld_x32(tS + 0, s + 0);
        tmem_ld_x32(tS + 32, s + 32);
        tmem_ld_x32(tS + 64, s + 64);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(b_sfree);  // S_{g+1} may be computed while this block's exponentials run
        RV_PPTL_J(8 * j + 1);
        const int nvalid = args.seq - j * kAttnBKV;  // keys beyond it are padding (last block only)
        if (nvalid < kAttnBKV) {
#pragma unroll
          for (int i = 0; i < kAttnBKV; ++i)
            if (i >= nvalid) s[i] = 0xFF800000u;  // -inf -> P = 0
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < kAttnBKV; i += 8) {
          mx0 = fmax3(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
          mx2 = fmax3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
          mx3 = fmax3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
        }
        const float mb = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
        float alpha = 1.f;
        bool need = false;
        if (mb > m_ref + kAttnRescaleThreshold) {
          alpha = exp2f(m_ref - mb);  // 0 on the first block (m_ref = -inf)
          m_ref = mb;
          need = (j > 0);
        }
        RV_PPTL_J(8 * j + 2);
        RV_PPTL_J(8 * j + 3);
        if (__any_sync(0xffffffffu, need)) {  // rare: the reference moved, rescale this row of O
          mbar_wait(b_o, static_cast<uint32_t>((g - 1) & 1));  // PV_{g-1} complete
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kAttnHdPad / 8; ++c) {
            uint32_t o[8];
            tmem_ld_x8(tO + c * 8, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x8(tO + c * 8, o);
          }
        }
        if (j == num_kv - 1 && args.lse != nullptr)  // final reference of the item, for the epilogue warps' lse
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sM + static_cast<uint32_t>(((it & 1) * kPpGroups + grp) * kAttnBQ + r) * 4u), "f"(m_ref) : "memory");
        const float neg_m = -m_ref;
        // ---- MUFU-free part: the polynomial share of the exponentials
        uint32_t psign = 0;  // polynomial results are never negative: (psign >> 31) == 0
#pragma unroll
        for (int i = 0; i < kAttnBKV; ++i) {
          if ((i & 15) >= 16 - RV_PP_POLY_PER_16) {
            s[i] = __float_as_uint(exp2_poly3(fmaf(__uint_as_float(s[i]), sc, neg_m)));
            psign |= s[i];
          }
        }
        RV_PPTL_J(8 * j + 4);
        // ---- this group's turn on the XU pipe.  ptxas schedules arithmetic freely across BAR instructions, so the
        //      burst is tied to the barriers by data: its exponents depend on a (zero) word loaded from shared
        //      memory behind the bar.sync, and the thread count of the bar.arrive depends on the last results.
        if (grp == 0) turn_wait<1>(psign >> 31); else turn_wait<2>(psign >> 31);
        float zero;
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(zero) : "r"(zero_smem) : "memory");
        const float neg_m2 = neg_m + zero;
        RV_PPTL_J(8 * j + 5);
        // The packs of a 16-column group are issued one group behind its exponentials: a pack right behind its
        // two MUFU.EX2 stalls this (only) MUFU-issuing warp of the scheduler until they complete and the XU queue
        // drains meanwhile (measured: the burst ran at 65 % of the XU rate).
        uint32_t pk[kAttnBKV / 2];
        uint32_t sign = 0;  // no P is negative: (sign >> 31) == 0
        const bool short_blk = short_last && j == num_kv - 1;  // 64 score columns only (block-uniform)
        auto burst = [&](auto nch) {
          constexpr int kCh = decltype(nch)::value;   // 16-column groups of this block
#pragma unroll
          for (int c = 0; c <= kCh; ++c) {
            if (c < kCh) {
#pragma unroll
              for (int i = 16 * c; i < 16 * c + 16 - RV_PP_POLY_PER_16; ++i)
                s[i] = __float_as_uint(ex2_approx_v(fmaf(__uint_as_float(s[i]), sc, neg_m2)));
            }
            if (c > 0) {
#pragma unroll
              for (int i = 16 * (c - 1); i < 16 * c; i += 2) {
                pk[i >> 1] = pack_bf16x2_v(__uint_as_float(s[i]), __uint_as_float(s[i + 1]));
                sign |= pk[i >> 1];
              }
            }
          }
        };
        if (short_blk) burst(std::integral_constant<int, 4>{});
        else burst(std::integral_constant<int, kAttnBKV / 16>{});
        if (grp == 0) {
          turn_pass<2>(sign >> 31);
        } else if (g + 1 < total_blocks) {
          turn_pass<1>(sign >> 31);
        }
        RV_PPTL_J(8 * j + 6);
        // ---- this row's 48 packed columns of the P region (PV_{g-1} must have read it: it normally has, it was
        //      issued a whole block ago)
        if (g > 0) mbar_wait(b_o, static_cast<uint32_t>((g - 1) & 1));
        tmem_st_x32(tP, pk);
        if (!short_blk) tmem_st_x16(tP + 32, pk + 32);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(b_p);
        RV_PPTL_J(8 * j + 7);
      }

    }
  } else if (warp_u >= kPpFirstEpilogueWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPpRegsEpilogue));
    // ===================== epilogue warps: O / l of a finished item -> registers (then the next item's PV_0 may
    // overwrite O) -> bf16 -> out[(tile*seq + t), head*hd + d], off the softmax warps' critical path ===========
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    for (int it = 0; it < num_items; ++it) {
      const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int th = w / args.num_qblk, qblk = w - th * args.num_qblk;
      const int tile = th / args.heads, head = th - tile * args.heads;
#pragma unroll 1
      for (int grp = 0; grp < kPpGroups; ++grp) {
        const uint32_t tO = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                            static_cast<uint32_t>(grp * kPpTmemGroup + kPpTmemO);
        mbar_wait(bar_ofin + 8 * grp, static_cast<uint32_t>(it & 1));  // last PV of the item complete
        tc_fence_after();
        RV_PPTL(24);
        uint32_t o[kAttnHdPad];
#pragma unroll
        for (int c = 0; c < kAttnHdPad / 8; ++c) tmem_ld_x8(tO + c * 8, o + c * 8);
        const float l = __uint_as_float(tmem_ld_x1(tO + static_cast<uint32_t>(args.hd)));  // ones column of V
        tmem_wait_ld();
        tc_fence_before();
        float m_ref = 0.f;
        if (args.lse != nullptr)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m_ref)
                       : "r"(sM + static_cast<uint32_t>(((it & 1) * kPpGroups + grp) * kAttnBQ + r) * 4u) : "memory");
        mbar_arrive(bar_ofree + 8 * grp);
        RV_PPTL(25);
        const int t = qblk * kPpItemRows + grp * kAttnBQ + r;
        const float inv_l = 1.0f / l;
#ifndef RV_ATTN_TIMELINE
        if (args.lse != nullptr && t < args.seq)  // softmax = 2^(s * scale * log2e - lse)
          args.lse[static_cast<size_t>(th) * args.seq_pad + t] = m_ref + log2f(l);
#endif
        // rows -> staging tile [128][hd] (16-byte stores, conflict-free per quarter warp at a 144-byte pitch), then
        // ONE TMA store of the tile: per-thread global stores of 144-byte rows 2304 bytes apart cost 288 sector
        // writes per warp and stalled the softmax warps' TMEM loads behind them in the LSU / MIO queue.
        named_bar_sync(3, kPpEpilogueThreads);  // the previous tile's store has finished reading the staging tile
        const uint32_t row_smem = sO + static_cast<uint32_t>(r * args.hd) * 2u;
#pragma unroll
        for (int c = 0; c < kAttnHdPad / 8; ++c) {
          if (c * 8 < args.hd) {
            const uint32_t v0 = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv_l, __uint_as_float(o[8 * c + 1]) * inv_l);
            const uint32_t v1 = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv_l, __uint_as_float(o[8 * c + 3]) * inv_l);
            const uint32_t v2 = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv_l, __uint_as_float(o[8 * c + 5]) * inv_l);
            const uint32_t v3 = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv_l, __uint_as_float(o[8 * c + 7]) * inv_l);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_smem + c * 16), "r"(v0), "r"(v1), "r"(v2),
                         "r"(v3) : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(4, kPpEpilogueThreads);
        if (warp == kPpFirstEpilogueWarp && lane == 0) {
          tma_store_3d(&tmap_out, sO, head * args.hd, qblk * kPpItemRows + grp * kAttnBQ, tile);  // rows >= seq dropped
          tma_store_commit();
          tma_store_wait_read();
        }
        RV_PPTL(26);
      }
    }
    if (warp == kPpFirstEpilogueWarp && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kPpTmemCols);
  }
}

}  // namespace rv
