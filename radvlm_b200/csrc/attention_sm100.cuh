// Non-causal multi-head attention for the SigLIP tower on sm_100a (729 tokens, head_dim 72).
//
// Replaces the materialised-score attention of SigLipAttention.forward
//   (finetuning/llava/model/multimodal_encoder/siglip_encoder.py:216-235:
//    matmul(q,k^T)*scale -> softmax(fp32) -> cast -> matmul(p,v) -> transpose/reshape)
// with one fused kernel.  Layout contract (produced by the QKV GEMM epilogue, gemm_sm100.cuh):
//    Q, K : [tiles*heads, seq_pad, hd_pad] bf16, zero padded (hd 72 -> 80, seq 729 -> 768)
//    V    : [tiles*heads, seq_pad, hd_pad] bf16, same layout as K (the QKV epilogue writes all three with 16-byte
//           row stores); it is the B operand of O += P V as an MN-major tcgen05 operand (head_dim contiguous), so no
//           transposed copy exists anywhere.  Column `hd` of every valid key holds a one, so the tensor core also
//           accumulates the softmax row sum (O[:, hd] = sum_j P_j 1) from exactly the bf16 P values it multiplies
//    out  : [tiles*seq, heads*hd] bf16 token-major (the A operand of out_proj)
//
// A work item = 128 query rows of one (tile, head).  The kernel is persistent: two CTAs are resident per SM (256
// TMEM columns and ~100 KB shared memory each) and each walks a strided list of items, so barrier set-up, the
// TMEM allocation and above all the first Q / K / V load latency are paid once per CTA, not once per item.  Warp roles: warp0 TMA producer, warp1 tcgen05.mma issuer, warps 2..9 softmax.
// S = Q K_j^T (128x128 fp32) and the running O (128x80 fp32) live in TMEM.  P_j never touches shared memory:
// it is written back into TMEM over S_j (bf16 pairs, 64 columns) and read from there as the A operand of
// O += P_j V_j.  ncu showed the earlier smem-P version bound by shared-memory bandwidth (both CTAs together
// moved ~113 B/clk of the 128 B/clk an SM has: Q/K/P/V operand reads of N<=128 MMAs + P stores + TMA writes);
// P in TMEM removes 64 of the 164 KB a key block moved.
//
// What is left is bound by the MUFU unit (16 ex2 / clk / SM, measured: tools/bench_mufu.cu), so the softmax is
// organised to keep that pipe fed: TWO threads per query row (each owns 64 of the 128 key columns of a block),
// ~3 issue slots per element (FFMA + EX2 + half a pack + half a 3-input max), no row-sum adds (ones row above).
// A thread pulls its 64 columns into registers with one TMEM round trip; the block maximum is exchanged between
// the two half-row threads through shared memory, so every exponential uses an exact, current reference: no
// clamps, nothing stale.  Within a CTA the chain softmax_j -> PV_j -> S_{j+1} is serial (P aliases S); the
// second resident CTA fills the tensor pipe / MUFU while the first is in the other phase.
// Online softmax with lazy rescale: the reference maximum only moves when the row maximum grows by more than
// 2^8, so the TMEM O rescale (tcgen05.ld -> mul -> tcgen05.st) is rare.
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
  float* lse;          // optional [tiles*heads, seq_pad]: log2-domain log-sum-exp of the scaled scores (for backward)
  int num_qblk;        // ceil(seq / 128) query blocks per (tile, head)
  int total_items;     // tiles * heads * num_qblk
  int trim_last;       // ping-pong kernel: compute a last key block with <= 64 valid keys 64 wide (default 1)
};

constexpr int kAttnBQ = 128;       // query rows per work item
constexpr int kAttnBKV = 96;       // keys per inner step (S 96 + O 80 + P 48 TMEM columns <= 256: P does not alias S)
constexpr int kAttnHdPad = 80;     // padded head dim (5 x UMMA_K)
constexpr int kAttnSoftmaxWarps = 8;
constexpr int kAttnSoftmaxThreads = kAttnSoftmaxWarps * 32;
constexpr int kAttnThreads = 64 + kAttnSoftmaxThreads;   // 320
constexpr int kAttnHalf = kAttnBKV / 2;                  // key columns per softmax thread
constexpr int kAttnQBytes = kAttnBQ * kAttnHdPad * 2;    // 20480 : [128 x 128 B] (SW128) + [128 x 32 B] (SW32)
constexpr int kAttnQ2Off = kAttnBQ * 128;                // offset of the SW32 slab (head dims [64,80))
constexpr int kAttnKBytes = kAttnBKV * kAttnHdPad * 2;   // 15360 : [96 x 128 B] (SW128) + [96 x 32 B] (SW32)
constexpr int kAttnK2Off = kAttnBKV * 128;
constexpr int kAttnVBytes = kAttnBKV * kAttnHdPad * 2;   // 15360 : 5 chunks of [96 keys x 32 B] (SW32), MN-major B operand
constexpr int kAttnVChunk = kAttnBKV * 32;               // 3072
constexpr int kAttnStages = 3;                           // K / V ring depth
constexpr int kAttnXchBytes = 2 * 2 * kAttnBQ * 2;       // 1024 : [block parity][column half][row] bf16 block maxima
// No alignment slack: the dynamic shared window of a kernel without static shared memory starts 1024-byte aligned
// (checked at run time).
constexpr int kAttnSmemBytes = kAttnQBytes + kAttnStages * (kAttnKBytes + kAttnVBytes) + kAttnXchBytes + 256;
static_assert(2 * (kAttnSmemBytes + 1024) <= 228 * 1024, "two attention CTAs must fit one SM");
static_assert(kAttnKBytes % 1024 == 0 && kAttnVBytes % 1024 == 0, "operand buffers must stay 1024-byte aligned");
constexpr int kAttnTmemCols = 256;  // S: [0,96)  O: [96,176)  P (bf16 pairs): [176,224)
constexpr int kAttnTmemO = kAttnBKV, kAttnTmemP = kAttnBKV + kAttnHdPad;
static_assert(kAttnTmemP + kAttnBKV / 2 <= kAttnTmemCols, "TMEM budget");
constexpr float kAttnRescaleThreshold = 8.0f;
// Exponentials per 16 computed on the FMA pipe (degree-3 polynomial) instead of MUFU.EX2.
#ifndef RV_ATTN_POLY_PER_16
#define RV_ATTN_POLY_PER_16 2
#endif

// 2^x on the FMA pipe: round-to-nearest split x = n + f (magic-number add), degree-3 minimax of 2^f on
// [-0.5, 0.5] (max relative error 1.1e-4, far below half a bf16 ulp), exponent inserted with an integer add.
__device__ __forceinline__ float exp2_poly3(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23
  const float n = t - 12582912.f;
  const float f = x - n;
  float p = fmaf(f, 0.05459282547f, 0.24221783876f);
  p = fmaf(p, f, 0.69336861372f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Optional timeline for tuning (-DRV_ATTN_TIMELINE): clock64 stamps of the first two work items of a few CTAs go to
// args.lse (reinterpreted as long long[], [cta][64]).  Compiled out by default.
#ifdef RV_ATTN_TIMELINE
#define RV_TL(slot)                                                                                                  \
  do {                                                                                                                 \
    if (args.lse != nullptr && it < 2 && (slot) < 64)                                                                  \
      reinterpret_cast<long long*>(args.lse)[static_cast<size_t>(blockIdx.x) * 64 + (slot)] = clock64();              \
  } while (0)
#else
#define RV_TL(slot) do { } while (0)
#endif

// Work item w = (tile * heads + head) * num_qblk + qblk; CTA c processes w = c, c + gridDim.x, ...  (consecutive
// CTAs share a head's K / V through L2).  `g` counts key blocks over all of a CTA's items: the K / V rings and the
// S / P / O barriers keep running across items, so the producer prefetches the next item's Q / K / V while the
// current item finishes and the per-item cost is only the O read-out.
__global__ void __launch_bounds__(kAttnThreads, 2)
siglip_attention_kernel(const __grid_constant__ CUtensorMap tmap_q,    // Q  columns [0,64)  : SW128 box {64, 128}
                        const __grid_constant__ CUtensorMap tmap_q2,   // Q  columns [64,80) : SW32  box {16, 128}
                        const __grid_constant__ CUtensorMap tmap_k,    // K  columns [0,64)  : SW128 box {64, 96}
                        const __grid_constant__ CUtensorMap tmap_k2,   // K  columns [64,80) : SW32  box {16, 96}
                        const __grid_constant__ CUtensorMap tmap_v,    // V  16-column chunks : SW32  box {16, 96}
                        const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();  // swizzled operand tiles need 1024-byte alignment
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + kAttnQBytes;               // ring: K_g lives in sK + (g % stages) * kAttnKBytes
  const uint32_t sV = sK + kAttnStages * kAttnKBytes; // ring: V_g lives in sV + (g % stages) * kAttnVBytes
  const uint32_t sX = sV + kAttnStages * kAttnVBytes; // [2 parities][2 halves][128 rows] bf16 block maxima
  const uint32_t bar_base = sX + kAttnXchBytes;
  const uint32_t bar_k = bar_base + 0;        // [stages] K_g landed
  const uint32_t bar_v = bar_base + 32;       // [stages] V_g landed
  const uint32_t bar_kfree = bar_base + 64;   // [stages] S_g complete: its K buffer may be refilled
  const uint32_t bar_vfree = bar_base + 96;   // [stages] PV_g complete: its V buffer may be refilled
  const uint32_t bar_s = bar_base + 128;      // S_g complete in TMEM
  const uint32_t bar_sfree = bar_base + 136;  // S_g is in registers: S_{g+1} may overwrite it
  const uint32_t bar_p = bar_base + 144;      // P_g in TMEM, O rescaled
  const uint32_t bar_o = bar_base + 152;      // O += P_g V_g complete (P region free again)
  const uint32_t bar_q = bar_base + 160;      // Q of item `it` landed
  const uint32_t bar_qfree = bar_base + 168;  // last S of the item complete: Q may be overwritten
  const uint32_t bar_ofree = bar_base + 176;  // the item's O is in registers: the next item's PV_0 may overwrite it
  const uint32_t tmem_ptr_smem = bar_base + 184;
  static_assert(kAttnStages <= 4, "barrier layout");

  const int warp = threadIdx.x >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // provably warp-uniform copy (role dispatch)
  const int lane = threadIdx.x & 31;
  const int num_kv = args.seq_pad / kAttnBKV;
  const int num_items = (args.total_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_q2);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_k2);
    tma_prefetch_desc(&tmap_v);
    for (uint32_t i = 0; i < kAttnStages; ++i) {
      mbar_init(bar_k + 8 * i, 1);
      mbar_init(bar_v + 8 * i, 1);
      mbar_init(bar_kfree + 8 * i, 1);
      mbar_init(bar_vfree + 8 * i, 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_sfree, kAttnSoftmaxThreads);
    mbar_init(bar_p, kAttnSoftmaxThreads);
    mbar_init(bar_o, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_qfree, 1);
    mbar_init(bar_ofree, kAttnSoftmaxThreads);
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
  const uint32_t tO = tmem_base + kAttnTmemO;
  const uint32_t tP = tmem_base + kAttnTmemP;

  if (warp_u == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int g = 0;
      uint32_t slot = 0, ring_par = 0;  // slot = g % stages, ring_par = (g / stages) & 1
      for (int it = 0; it < num_items; ++it) {
        const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        const int th = w / args.num_qblk, qblk = w - th * args.num_qblk;
        if (it > 0) mbar_wait(bar_qfree, static_cast<uint32_t>((it - 1) & 1));  // previous item's last S complete
        mbar_arrive_expect_tx(bar_q, kAttnQBytes);
        const int q_row0 = th * args.seq_pad + qblk * kAttnBQ;
        tma_load_2d(sQ, &tmap_q, bar_q, 0, q_row0);
        tma_load_2d(sQ + kAttnQ2Off, &tmap_q2, bar_q, 64, q_row0);
        for (int j = 0; j < num_kv; ++j, ++g) {
          if (g >= kAttnStages) mbar_wait(bar_kfree + 8 * slot, ring_par ^ 1u);  // S_{g-stages} complete
          mbar_arrive_expect_tx(bar_k + 8 * slot, kAttnKBytes);
          const int k_row0 = th * args.seq_pad + j * kAttnBKV;
          tma_load_2d(sK + slot * kAttnKBytes, &tmap_k, bar_k + 8 * slot, 0, k_row0);
          tma_load_2d(sK + slot * kAttnKBytes + kAttnK2Off, &tmap_k2, bar_k + 8 * slot, 64, k_row0);
          if (g >= kAttnStages) mbar_wait(bar_vfree + 8 * slot, ring_par ^ 1u);  // PV_{g-stages} complete
          mbar_arrive_expect_tx(bar_v + 8 * slot, kAttnVBytes);
#pragma unroll
          for (int c = 0; c < 5; ++c)
            tma_load_2d(sV + slot * kAttnVBytes + c * kAttnVChunk, &tmap_v, bar_v + 8 * slot, c * 16, k_row0);
          if (++slot == kAttnStages) { slot = 0; ring_par ^= 1u; }
        }
      }
    }
  } else if (warp_u == 1) {
    // ===================== MMA issuer: the whole warp runs this code converged, elect.sync picks the issuing lane
    // inside each asm block (see umma_bf16_ss_elect) =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(kAttnBQ, kAttnBKV);
    constexpr uint32_t idesc_o = make_idesc_bf16(kAttnBQ, kAttnHdPad) | (1u << 16);  // B (= V) is MN-major
    const uint32_t tS_u = __shfl_sync(0xffffffffu, tS, 0), tO_u = __shfl_sync(0xffffffffu, tO, 0);
    const uint32_t tP_u = __shfl_sync(0xffffffffu, tP, 0);
    const uint64_t qd128 = make_smem_desc(sQ, 1024, kLayoutSw128);             // head dims [0,64): 32 B per K step
    const uint64_t qd32 = make_smem_desc(sQ + kAttnQ2Off, 256, kLayoutSw32);   // head dims [64,80)
    const uint64_t kd128 = make_smem_desc(sK, 1024, kLayoutSw128);
    const uint64_t kd32 = make_smem_desc(sK + kAttnK2Off, 256, kLayoutSw32);
    // V block [96 keys][80] as an MN-major B operand (N = head_dim): 16-column groups kAttnVChunk bytes apart (LBO),
    // 8 keys per 256-byte swizzle atom (SBO), 512 bytes per 16-key K step
    const uint64_t vd = make_smem_desc_lbo(sV, kAttnVChunk, 256, kLayoutSw32);
    const int total_blocks = num_items * num_kv;
    // ring bookkeeping for S issue (runs one block ahead of the PV issue)
    uint32_t s_slot = 0, s_par = 0;
    auto issue_s = [&](int g, int it, int j) {  // S_g = Q_it K_g^T
      if (j == 0) mbar_wait(bar_q, static_cast<uint32_t>(it & 1));  // Q of this item landed
      mbar_wait(bar_k + 8 * s_slot, s_par);                          // K_g landed
      if (g > 0) mbar_wait(bar_sfree, static_cast<uint32_t>((g - 1) & 1));  // S_{g-1} is in registers
      tc_fence_after();
      const uint64_t koff = static_cast<uint64_t>(s_slot * (kAttnKBytes >> 4));
#pragma unroll
      for (int c = 0; c < 4; ++c) umma_bf16_ss_elect(tS_u, qd128 + 2 * c, kd128 + koff + 2 * c, idesc_s, c != 0 ? 1u : 0u);
      umma_bf16_ss_elect(tS_u, qd32, kd32 + koff, idesc_s, 1u);
      umma_commit_elect(bar_s);
      umma_commit_elect(bar_kfree + 8 * s_slot);
      if (lane == 0 && it == 0) RV_TL(48 + 2 * j);
      if (j == num_kv - 1) umma_commit_elect(bar_qfree);
      if (++s_slot == kAttnStages) { s_slot = 0; s_par ^= 1u; }
    };
    if (total_blocks > 0) issue_s(0, 0, 0);
    int g = 0;
    uint32_t slot = 0, ring_par = 0;
    for (int it = 0; it < num_items; ++it) {
      for (int j = 0; j < num_kv; ++j, ++g) {
        // S_{g+1} first: it only needs S_g to have been read out, and runs while the softmax of block g computes
        if (g + 1 < total_blocks) {
          const bool wrap = (j + 1 == num_kv);
          issue_s(g + 1, wrap ? it + 1 : it, wrap ? 0 : j + 1);
        }
        mbar_wait(bar_p, static_cast<uint32_t>(g & 1));  // P_g in TMEM, O rescaled
        mbar_wait(bar_v + 8 * slot, ring_par);            // Vt_g landed
        if (j == 0 && it > 0) mbar_wait(bar_ofree, static_cast<uint32_t>((it - 1) & 1));  // previous O read out
        tc_fence_after();
        const uint64_t voff = static_cast<uint64_t>(slot * (kAttnVBytes >> 4));
#pragma unroll
        for (int s = 0; s < kAttnBKV / 16; ++s)  // A = P_g from TMEM (8 columns = 16 bf16 per K step)
          umma_bf16_ts_elect(tO_u, tP_u + static_cast<uint32_t>(s * 8), vd + voff + 32 * s, idesc_o, (j | s) != 0 ? 1u : 0u);
        umma_commit_elect(bar_o);
        umma_commit_elect(bar_vfree + 8 * slot);
        if (lane == 0 && it == 0) RV_TL(49 + 2 * j);
        if (++slot == kAttnStages) { slot = 0; ring_par ^= 1u; }
      }
    }
  } else {
    // ===================== softmax / correction / output (8 warps, two threads per query row) ==========
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;        // which half of the key columns of a block
    const int r = quad * 32 + lane;          // row within the query block
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tSh = tS + lane_off + static_cast<uint32_t>(half * kAttnHalf);
    const uint32_t tOh = tO + lane_off + static_cast<uint32_t>(half * 40);           // this thread's 40 O columns
    const uint32_t tPh = tP + lane_off + static_cast<uint32_t>(half * (kAttnHalf / 2));  // its packed P columns
    const float sc = args.scale_log2e;
    const uint32_t x_own = sX + static_cast<uint32_t>(half * kAttnBQ + r) * 2u;
    const uint32_t x_other = sX + static_cast<uint32_t>((half ^ 1) * kAttnBQ + r) * 2u;

    int g = 0;
    for (int it = 0; it < num_items; ++it) {
      float m_ref = -INFINITY;  // reference maximum (scaled, log2 domain), identical in both threads of a row
      for (int j = 0; j < num_kv; ++j, ++g) {
        mbar_wait(bar_s, static_cast<uint32_t>(g & 1));
        tc_fence_after();
        if (warp == 2 && lane == 0 && it == 0) RV_TL(6 * j + 0);
        // keys >= nvalid (relative to this thread's first column) are padding: last block only
        const int nvalid = args.seq - j * kAttnBKV - half * kAttnHalf;

        // ---- S_g (this thread's 48 columns) -> registers with one TMEM round trip, then S is released: the MMA
        //      warp computes S_{g+1} while the exponentials below run
        uint32_t s[kAttnHalf];
        tmem_ld_x32(tSh + 0, s + 0);
        tmem_ld_x16(tSh + 32, s + 32);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(bar_sfree);
        if (warp == 2 && lane == 0 && it == 0) RV_TL(6 * j + 1);
        if (nvalid < kAttnHalf) {
#pragma unroll
          for (int i = 0; i < kAttnHalf; ++i)
            if (i >= nvalid) s[i] = 0xFF800000u;  // -inf -> P = 0
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < kAttnHalf; i += 4) {
          mx0 = fmax3(mx0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        }
        // exchange the block maximum with the thread that owns the other half of this row's columns (both use the
        // bf16-rounded values, so both compute the same reference).  The slot alternates with the block parity: a
        // slot is rewritten two blocks later, after another bar.sync that the partner only reaches past its read.
        const uint32_t xpar = static_cast<uint32_t>(g & 1) * (2u * kAttnBQ * 2u);
        const __nv_bfloat16 own = __float2bfloat16_rn(fmaxf(mx0, mx1) * sc);
        const uint16_t own_bits = *reinterpret_cast<const uint16_t*>(&own);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(x_own + xpar), "h"(own_bits) : "memory");
        switch (quad) {  // compile-time barrier ids (a register id would reserve all 16 hardware barriers)
          case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
          case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
          case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
          default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
        }
        uint16_t other_bits;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(other_bits) : "r"(x_other + xpar) : "memory");
        const float mo = __uint_as_float(static_cast<uint32_t>(other_bits) << 16);
        const float mb = fmaxf(__uint_as_float(static_cast<uint32_t>(own_bits) << 16), mo);
        float alpha = 1.f;
        bool need = false;
        if (mb > m_ref + kAttnRescaleThreshold) {
          alpha = exp2f(m_ref - mb);  // 0 on the first block (m_ref = -inf)
          m_ref = mb;
          need = (j > 0);
        }
        // PV_{g-1} must be complete before the P region is overwritten or O is rescaled (it normally is: it was
        // issued a whole softmax block ago)
        if (warp == 2 && lane == 0 && it == 0) RV_TL(6 * j + 2);
        if (g > 0) mbar_wait(bar_o, static_cast<uint32_t>((g - 1) & 1));
        if (warp == 2 && lane == 0 && it == 0) RV_TL(6 * j + 3);
        if (__any_sync(0xffffffffu, need)) {  // rare: the reference moved, rescale this thread's 40 O columns
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            uint32_t o[8];
            tmem_ld_x8(tOh + c * 8, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x8(tOh + c * 8, o);
          }
        }
        const float neg_m = -m_ref;

        // ---- P = 2^(s * scale * log2e - m_ref) -> bf16 pairs -> this thread's 24 packed columns of the P region
        uint32_t pk[kAttnHalf / 2];
#pragma unroll
        for (int i = 0; i < kAttnHalf; i += 2) {
          const float a0 = fmaf(__uint_as_float(s[i]), sc, neg_m);
          const float a1 = fmaf(__uint_as_float(s[i + 1]), sc, neg_m);
          float p0, p1;
          if ((i & 15) >= 16 - RV_ATTN_POLY_PER_16) {
            p0 = exp2_poly3(a0);
            p1 = exp2_poly3(a1);
          } else {
            p0 = ex2_approx(a0);
            p1 = ex2_approx(a1);
          }
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        if (warp == 2 && lane == 0 && it == 0) RV_TL(6 * j + 4);
        tmem_st_x16(tPh, pk);
        tmem_st_x8(tPh + 16, pk + 16);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p);
        if (warp == 2 && lane == 0 && it == 0) RV_TL(6 * j + 5);
      }

      // ---- item done: O / l -> registers (then the next item's PV_0 may overwrite O) -> bf16 ->
      //      out[(tile*seq + t), head*hd + d]; this thread writes columns [40*half, +40) < hd
      mbar_wait(bar_o, static_cast<uint32_t>((g - 1) & 1));  // PV of the item's last key block complete
      tc_fence_after();
      uint32_t o[40];
#pragma unroll
      for (int c = 0; c < 5; ++c) tmem_ld_x8(tOh + c * 8, o + c * 8);
      const uint32_t l_bits = tmem_ld_x1(tO + lane_off + static_cast<uint32_t>(args.hd));  // ones row of V^T
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_ofree);
      const int w = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int th = w / args.num_qblk, qblk = w - th * args.num_qblk;
      const int tile = th / args.heads, head = th - tile * args.heads;
      const int t = qblk * kAttnBQ + r;
      const float inv_l = 1.0f / __uint_as_float(l_bits);
#ifndef RV_ATTN_TIMELINE
      if (args.lse != nullptr && half == 0 && t < args.seq)  // softmax = 2^(s * scale * log2e - lse)
        args.lse[static_cast<size_t>(th) * args.seq_pad + t] = m_ref + log2f(__uint_as_float(l_bits));
#endif
      if (t < args.seq) {
        __nv_bfloat16* dst = args.out +
                             (static_cast<size_t>(tile) * args.seq + t) * (args.heads * args.hd) +
                             head * args.hd + half * 40;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          if (half * 40 + c * 8 < args.hd) {
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}

}  // namespace rv
