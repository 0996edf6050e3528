// Argument block and epilogue selector of the tcgen05 GEMM (shared by device and host code).
#pragma once
#include <cuda_bf16.h>

namespace rv {

enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,       // out_bf16 = acc + bias
  EPI_GELU_TANH_BF16 = 1,  // out_bf16 = gelu_tanh(acc + bias)           (SigLipMLP.fc1, siglip_encoder.py:252-253)
  EPI_GELU_ERF_BF16 = 2,   // out_bf16 = gelu_erf(acc + bias)            (mm_projector.0/1, builder.py:44-46)
  EPI_RESID_F32 = 3,       // out_f32  = acc + bias + resid_f32          (residual adds, siglip_encoder.py:293,298)
  EPI_POS_F32 = 4,         // out_f32  = acc + bias + pos[row % period]  (siglip_encoder.py:170-173)
  EPI_QKV_SPLIT = 5,       // head split scatter of q/k/v                (siglip_encoder.py:207-213)
  EPI_BIAS_F32 = 6,        // out_f32  = acc + bias
  EPI_ATOMIC_F32 = 7,      // out_f32 += acc (red.global.add; split-K partial sums, weight-gradient accumulation)
  EPI_GELU_TANH_DUAL_BF16 = 8,  // out_bf16 = gelu_tanh(u) and out2_bf16 = gelu_tanh'(u), u = acc + bias (training forward
                                // of fc1: the derivative is kept for the backward, one tanh serves both)
  EPI_MUL_BF16 = 10,       // out_bf16 = bf16(acc) * m, m = out2 (bf16 [M, ldo], READ): the fc2 data gradient times the
                           // kept gelu'(u) = the GELU backward fused into the GEMM (replaces a separate HBM pass)
  EPI_BIAS_F16 = 9,        // out_f16 = acc + bias (projector output when the model serves in fp16,
                           // serve/model_worker.py:124-127, model/builder.py:289-294)
  EPI_DELTA_BF16 = 11,     // out_proj when LayerNorm is folded (siglip_encoder.py:293): the fp32 stream is NOT rewritten.
                           //   out_bf16  = d = acc + bias                 (the attention branch, added to the stream by the
                           //                                               fc2 epilogue of the same layer: aux16 there)
                           //   out2_bf16 = bf16(aux16 + d)               (aux16 = bf16 copy of the stream entering the layer:
                           //                                               the LayerNorm-2 input, A operand of fc1)
                           //   ln_part   : row sums of aux16 + d
                           // 8 B per element instead of the 12 of EPI_RESID_F32 + copy: out_proj is HBM-bound
};

struct GemmArgs {
  int M, N, K;
  const float* bias;  // [N] or nullptr
  void* out;          // [M, ldo] bf16 or f32 depending on the epilogue
  void* out2;         // EPI_GELU_TANH_DUAL_BF16: gelu'(u) [M, ldo] bf16 (written); EPI_MUL_BF16: the multiplier (read);
                      // EPI_RESID_F32 / EPI_POS_F32: optional bf16 copy of the fp32 result, [M, ldo] (written) - the A
                      // operand of the next GEMM when LayerNorm is folded into it (ln_stats below)
  int ldo;
  const float* aux;  // EPI_RESID_F32: residual [M, ldo];  EPI_POS_F32: table [aux_period, N]
  const __nv_bfloat16* aux16;  // EPI_DELTA_BF16: bf16 residual [M, ldo] (may alias out: read, then written, by the same
                               // thread);  EPI_RESID_F32: optional second addend [M, ldo] (the bf16 attention branch)
  int aux_period;
  // EPI_QKV_SPLIT
  __nv_bfloat16* q;   // [tiles, heads, seq_pad, hd_pad]
  __nv_bfloat16* k;   // [tiles, heads, seq_pad, hd_pad]
  __nv_bfloat16* vt;  // V: [tiles, heads, seq_pad, hd_pad] (same layout as q / k)
  int seq, seq_pad, heads, hd, hd_pad;
  // Operand layouts (backward GEMMs read activations / weights as they are stored, no transposes):
  //   a_mn = 0: A is [M, K] row-major (K-major operand);  a_mn = 1: A is stored as [K, M] row-major (MN-major)
  //   b_mn = 0: W is [N, K] row-major;                    b_mn = 1: W is stored as [K, N] row-major
  int a_mn, b_mn;
  int k_splits;  // >= 1: the K range is cut into k_splits partial products (EPI_ATOMIC_F32 only)
  // LayerNorm folded into this GEMM (EPI_QKV_SPLIT, EPI_GELU_TANH_BF16, EPI_BIAS_BF16):  with A = bf16(x) (NOT
  // normalised), W = bf16(gamma o W0), ln_s[n] = sum_k W[n, k], bias = b0 + W0 beta and ln_stats[row] = (mean, rstd) of
  // row `row` of x,   out = rstd * (acc - mean * ln_s[n]) + bias[n]  ==  LayerNorm(x) W0^T + b0.   nullptr: plain GEMM.
  const float2* ln_stats;
  const float* ln_s;
  // Producer side of the fold (EPI_RESID_F32 / EPI_POS_F32, scheduled kernel only): partial (sum, sum of squares) of every
  // output row over the columns one epilogue warp drains, one slot per (column tile, half): [M][ln_slots] float2, written
  // without atomics (every slot has exactly one writer); ln_finalize_stats_kernel turns them into (mean, rstd).
  float2* ln_part;
  int ln_slots;
  // Consumer side without a finalize pass: ln_stats == nullptr, ln_s != nullptr and ln_part != nullptr -> every epilogue
  // thread adds up the ln_slots partials of its row itself: mean = sum / ln_dim, rstd = rsqrt(E[x^2] - mean^2 + ln_eps).
  float ln_inv_dim, ln_eps;
  // tuning builds only (-DRV_GEMM_TIMELINE, tools/gemm_timeline.cu): clock64 stamps of CTA pair 0, [tile][8]
  long long* timeline;
};

}  // namespace rv
