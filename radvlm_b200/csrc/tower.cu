// Host orchestration of the encode path: SigLIP tower and mlp2x_gelu projector as a fixed sequence of
// kernel launches on one stream (no allocation, no synchronisation -> CUDA-graph capturable).
//
//   SigLipVisionTower.forward        siglip_encoder.py:576-589
//   SigLipVisionEmbeddings.forward   siglip_encoder.py:169-174
//   SigLipEncoderLayer.forward       siglip_encoder.py:285-305
//   encode_images / mm_projector     llava_arch.py:192-196 ; multimodal_projector/builder.py:41-48
#include <cstdlib>
#include <cstring>

#include "host_util.h"
#include "internal.h"

namespace rv {


struct EncodeLayout {
  int P, T, seq_pad, hd, hd_pad;
  size_t M;
  size_t off_hidden, off_xn, off_xb, off_ao, off_stats, off_part, off_ready, off_q, off_k, off_vt, off_h1, total;
  size_t qkv_bytes;
};

static int make_layout(const radvlm_siglip_weights* tw, const radvlm_projector_weights* pw, int n_tiles,
                       EncodeLayout* L) {
  RV_CHECK_ARG(tw != nullptr && n_tiles > 0, "encode: null weights or n_tiles <= 0");
  RV_CHECK_ARG(tw->patch_size > 0 && tw->heads > 0 && tw->hidden > 0, "encode: bad tower config");
  L->P = tw->image_size / tw->patch_size;
  L->T = L->P * L->P;
  L->seq_pad = (L->T + 383) / 384 * 384;  // multiple of the attention query (128) and key (96) block sizes
  L->hd = tw->hidden / tw->heads;
  L->hd_pad = 80;
  if (L->hd * tw->heads != tw->hidden || L->hd > L->hd_pad || (L->hd % 8) != 0 || (tw->hidden % 8) != 0 ||
      (tw->intermediate % 8) != 0 || (tw->patch_k_pad % 8) != 0 ||
      tw->patch_k_pad < tw->channels * tw->patch_size * tw->patch_size) {
    set_error("encode: unsupported tower geometry hidden=%d heads=%d intermediate=%d patch_k_pad=%d",
              tw->hidden, tw->heads, tw->intermediate, tw->patch_k_pad);
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  L->M = static_cast<size_t>(n_tiles) * L->T;
  const size_t D = tw->hidden;
  size_t xn_cols = D > static_cast<size_t>(tw->patch_k_pad) ? D : tw->patch_k_pad;
  size_t h1_cols = tw->intermediate;
  if (pw != nullptr && static_cast<size_t>(pw->hidden) > h1_cols) h1_cols = pw->hidden;
  L->qkv_bytes = static_cast<size_t>(n_tiles) * tw->heads * L->seq_pad * L->hd_pad * 2;
  size_t off = 0;
  L->off_hidden = off; off = align_up(off + L->M * D * 4, 1024);
  L->off_xn = off;     off = align_up(off + L->M * xn_cols * 2, 1024);
  L->off_xb = off;     off = align_up(off + L->M * D * 2, 1024);   // bf16 copy of the residual stream (LayerNorm fold)
  L->off_ao = off;     off = align_up(off + L->M * D * 2, 1024);   // attention output (LayerNorm fold: xb stays live)
  L->off_stats = off;  off = align_up(off + L->M * 8, 1024);       // (mean, rstd) per row
  L->off_part = off;   off = align_up(off + L->M * 8 * 2 * ((D + 255) / 256), 1024);  // per-row partial sums (epilogues)
  L->off_ready = off;  off = align_up(off + (L->M / 256 + 2) * 4, 1024);  // row-block counters of the chained GEMM kernels
  L->off_q = off;      off = align_up(off + L->qkv_bytes, 1024);
  L->off_k = off;      off = align_up(off + L->qkv_bytes, 1024);
  L->off_vt = off;     off = align_up(off + L->qkv_bytes, 1024);
  L->off_h1 = off;     off = align_up(off + L->M * h1_cols * 2, 1024);
  L->total = off;
  return RADVLM_OK;
}

// Activations kept for the backward pass (training mode): the fp32 residual stream at the input of every layer
// (+ the tower output), the attention output (A operand of out_proj) and the attention log-sum-exp of every layer.
// Everything else is recomputed layer by layer in tower_backward_impl (the reference checkpoints whole encoder
// layers too: siglip_encoder.py:381-387).
struct TowerSaved {
  size_t h_bytes, ao_bytes, lse_bytes, qkv_bytes, u_bytes, total;
  uint8_t* base;
  float* h(int l) const { return reinterpret_cast<float*>(base + static_cast<size_t>(l) * h_bytes); }
  void* ao(int l, int L) const { return base + static_cast<size_t>(L + 1) * h_bytes + static_cast<size_t>(l) * ao_bytes; }
  float* lse(int l, int L) const {
    return reinterpret_cast<float*>(base + static_cast<size_t>(L + 1) * h_bytes + static_cast<size_t>(L) * ao_bytes +
                                    static_cast<size_t>(l) * lse_bytes);
  }
  // mid-layer residual stream (after the attention branch): kept too - 87 MB per tile buys the out_proj recompute
  float* hmid(int l, int L) const {
    return reinterpret_cast<float*>(base + static_cast<size_t>(L + 1) * h_bytes +
                                    static_cast<size_t>(L) * (ao_bytes + lse_bytes) + static_cast<size_t>(l) * h_bytes);
  }
  // q / k / v of every layer (padded head-major): kept too, the backward does not recompute the QKV GEMM
  void* qkv(int l, int which, int L) const {
    return base + static_cast<size_t>(2 * L + 1) * h_bytes + static_cast<size_t>(L) * (ao_bytes + lse_bytes) +
           (static_cast<size_t>(l) * 3 + which) * qkv_bytes;
  }
  // gelu'(fc1 pre-activation) (bf16 [M, I]) of every layer: written by the dual-output GELU epilogue of the forward
  void* u(int l, int L) const {
    return base + static_cast<size_t>(2 * L + 1) * h_bytes +
           static_cast<size_t>(L) * (ao_bytes + lse_bytes + 3 * qkv_bytes) + static_cast<size_t>(l) * u_bytes;
  }
  // LayerNorm outputs (bf16 [M, D]; which = 0: LN1, 1: LN2): the A operands of the qkv / fc1 weight gradients
  void* xn(int l, int which, int L) const {
    return base + static_cast<size_t>(2 * L + 1) * h_bytes +
           static_cast<size_t>(L) * (ao_bytes + lse_bytes + 3 * qkv_bytes + u_bytes) +
           (static_cast<size_t>(l) * 2 + which) * ao_bytes;
  }
  // MLP activation gelu(u) (bf16 [M, I]): the fc1 epilogue writes it here instead of a shared scratch buffer (same
  // traffic), the backward uses it as the A operand of the fc2 weight gradient instead of recomputing it
  void* act(int l, int L) const {
    return base + static_cast<size_t>(2 * L + 1) * h_bytes +
           static_cast<size_t>(L) * (3 * ao_bytes + lse_bytes + 3 * qkv_bytes + u_bytes) + static_cast<size_t>(l) * u_bytes;
  }
};

static TowerSaved make_saved(const radvlm_siglip_weights* tw, const EncodeLayout& L, int n_tiles, void* base) {
  TowerSaved s;
  s.h_bytes = align_up(L.M * tw->hidden * 4, 1024);
  s.ao_bytes = align_up(L.M * tw->hidden * 2, 1024);
  s.lse_bytes = align_up(static_cast<size_t>(n_tiles) * tw->heads * L.seq_pad * 4, 1024);
  s.qkv_bytes = align_up(L.qkv_bytes, 1024);
  s.u_bytes = align_up(L.M * tw->intermediate * 2, 1024);
  s.total = (2 * tw->num_layers + 1) * s.h_bytes +
            tw->num_layers * (3 * s.ao_bytes + s.lse_bytes + 3 * s.qkv_bytes + 2 * s.u_bytes);
  s.base = static_cast<uint8_t*>(base);
  return s;
}

// `hidden`: inference: the in-place residual stream (= output).  Training (save != nullptr): scratch for the
// mid-layer residual; layer l reads save->h(l) and writes save->h(l + 1), the output is save->h(num_layers).
// LayerNorm fold (inference, every layer carries qkv_wf / fc1_wf): no LayerNorm kernel runs.  The epilogues that
// produce the stream (patch embedding, fc2) also write a bf16 copy of it (`xb`) and per-row partial sums; the QKV / fc1
// GEMMs multiply the copy by gamma o W and apply (mean, rstd) in their epilogues.  out_proj does NOT rewrite the fp32
// stream (it is the one HBM-bound GEMM): it reads `xb`, writes the attention branch d = acc + bias over it in bf16 and
// the LayerNorm-2 input bf16(xb + d) to `xn` (EPI_DELTA_BF16, 8 B per element instead of 12); the fc2 epilogue of the same
// layer then forms  stream += d + mlp  in fp32, so the fp32 stream is read and written once per layer, and only the
// branch - not the stream - is rounded to bf16.  *final_bf16 (optional) receives the pointer of the bf16 copy of the
// tower output (the projector's A operand: no cast pass).
static bool tower_ln_folded(const radvlm_siglip_weights* tw) {
  for (int l = 0; l < tw->num_layers; ++l) {
    const radvlm_vit_layer_weights& w = tw->layers[l];
    if (!w.qkv_wf || !w.qkv_sf || !w.qkv_bf || !w.fc1_wf || !w.fc1_sf || !w.fc1_bf) return false;
  }
  return tw->num_layers > 0;
}

static int tower_forward_impl(const radvlm_siglip_weights* tw, const void* pixels, int pixel_dtype,
                              int n_tiles, float* hidden, const EncodeLayout& L, uint8_t* ws,
                              cudaStream_t stream, const TowerSaved* save = nullptr, const void** final_bf16 = nullptr) {
  const int D = tw->hidden, I = tw->intermediate;
  const int M = static_cast<int>(L.M);
  const int NL = tw->num_layers;
  void* xn = ws + L.off_xn;
  void* q = ws + L.off_q;
  void* k = ws + L.off_k;
  void* vt = ws + L.off_vt;
  void* h1 = ws + L.off_h1;
  const bool fold = (save == nullptr) && tower_ln_folded(tw);
  void* xb = ws + L.off_xb;
  float2* stats = reinterpret_cast<float2*>(ws + L.off_stats);
  // row statistics: partial sums from the residual epilogues (scheduled GEMM kernel) + a finalize pass over them, or,
  // when that kernel does not cover the shape, a 2 B/element pass over the bf16 copy
  float2* part = reinterpret_cast<float2*>(ws + L.off_part);
  const int part_slots = fold ? gemm_ln_part_slots(M, D) : 0;
  auto row_stats = [&](const void* xbf) -> int {
    if (part_slots > 0) return RADVLM_OK;   // the consuming GEMM's epilogue adds the partials up itself
    ProfScope ps(PROF_LAYERNORM, stream);
    return ln_row_stats_launch(xbf, stats, M, D, tw->ln_eps, stream);
  };
  // fc1 -> fc2 through the chained kernel (gemm4_sm100.cuh) is opt-in (RADVLM_B200_MLP=chain): measured 55.1-56.2 ms per
  // step against 54.3 ms for the two launches (profiles/r02j_bench_*: the DRAM traffic it saves is hidden behind tensor
  // work anyway, the per-warp publish fences are not); the projector, whose second GEMM is 3x longer, does gain (-12 %)
  static const bool mlp_chain = getenv("RADVLM_B200_MLP") && !strcmp(getenv("RADVLM_B200_MLP"), "chain");
  // A/B switch: RADVLM_B200_OUTPROJ=f32 lets out_proj rewrite the fp32 stream itself (12 B per element) as before
  static const bool outproj_f32 = getenv("RADVLM_B200_OUTPROJ") && !strcmp(getenv("RADVLM_B200_OUTPROJ"), "f32");
  const bool delta = fold && !outproj_f32;
  auto set_consumer_stats = [&](GemmArgs& a, const float* row_sums) {
    a.ln_s = row_sums;
    if (part_slots > 0) {
      a.ln_part = part; a.ln_slots = part_slots;
      a.ln_inv_dim = 1.0f / static_cast<float>(D); a.ln_eps = tw->ln_eps;
    } else {
      a.ln_stats = stats;
    }
  };
  if (final_bf16 != nullptr) *final_bf16 = fold ? xb : nullptr;
  int st;

  // padding of q/k/vt must be zero (never written by the QKV epilogue); training uses per-layer slots instead
  if (save == nullptr) {
    ProfScope ps(PROF_MISC, stream, 0);
    RV_CUDA(cudaMemsetAsync(q, 0, L.qkv_bytes, stream));
    RV_CUDA(cudaMemsetAsync(k, 0, L.qkv_bytes, stream));
    int pst = attention_prepare_vt_launch(vt, n_tiles, tw->heads, L.T, L.seq_pad, L.hd, L.hd_pad, stream);
    if (pst) return pst;
  }

  // --- embeddings: patch GEMM + bias + position embedding (siglip_encoder.py:169-174)
  { ProfScope ps(PROF_MISC, stream);
  st = im2col_launch(pixels, pixel_dtype, xn, n_tiles, tw->channels, tw->image_size, tw->patch_size,
                     tw->patch_k_pad, stream); }
  if (st) return st;
  {
    GemmArgs a{};
    a.M = M; a.N = D; a.K = tw->patch_k_pad;
    a.bias = tw->patch_b;
    a.out = save ? save->h(0) : hidden; a.ldo = D;
    a.out2 = fold ? xb : nullptr;
    if (part_slots > 0) { a.ln_part = part; a.ln_slots = part_slots; }
    a.aux = tw->pos_embed; a.aux_period = L.T;
    { ProfScope ps(PROF_GEMM, stream); st = gemm_dispatch(xn, tw->patch_k_pad, tw->patch_w, tw->patch_k_pad, a, EPI_POS_F32, 0, stream); }
    if (st) return st;
  }

  const float scale = 1.0f / sqrtf(static_cast<float>(L.hd));
  for (int l = 0; l < tw->num_layers; ++l) {
    const radvlm_vit_layer_weights& w = tw->layers[l];
    const float* h_in = save ? save->h(l) : hidden;     // residual stream entering the layer
    float* h_out = save ? save->h(l + 1) : hidden;      // ... leaving it
    void* ao = save ? save->ao(l, NL) : (fold ? (delta ? static_cast<void*>(ws + L.off_ao) : xb) : xn);   // attention output (bf16)
    float* h_mid = save ? save->hmid(l, NL) : hidden;   // residual stream after the attention branch
    if (save != nullptr) {                              // this layer's own q / k / v slots: pads + ones column only
      q = save->qkv(l, 0, NL); k = save->qkv(l, 1, NL); vt = save->qkv(l, 2, NL);
      ProfScope ps(PROF_MISC, stream);
      if ((st = qkv_pad_prepare_launch(q, k, vt, n_tiles, tw->heads, L.T, L.seq_pad, L.hd, L.hd_pad, 1.0f, stream))) return st;
    }
    // x = x + out_proj(attn(LN1(x)))
    void* xn1 = save ? save->xn(l, 0, NL) : xn;
    void* xn2 = save ? save->xn(l, 1, NL) : xn;
    if (fold) st = row_stats(xb);
    else { ProfScope ps(PROF_LAYERNORM, stream); st = layernorm_launch(h_in, w.ln1_gamma, w.ln1_beta, xn1, M, D, tw->ln_eps, stream); }
    if (st) return st;
    {
      GemmArgs a{};
      a.M = M; a.N = 3 * D; a.K = D;
      a.bias = fold ? w.qkv_bf : w.qkv_b;
      if (fold) set_consumer_stats(a, w.qkv_sf);
      a.q = static_cast<__nv_bfloat16*>(q);
      a.k = static_cast<__nv_bfloat16*>(k);
      a.vt = static_cast<__nv_bfloat16*>(vt);
      a.seq = L.T; a.seq_pad = L.seq_pad; a.heads = tw->heads; a.hd = L.hd; a.hd_pad = L.hd_pad;
      { ProfScope ps(PROF_GEMM_QKV, stream); st = gemm_dispatch(fold ? xb : xn1, D, fold ? w.qkv_wf : w.qkv_w, D, a, EPI_QKV_SPLIT, 0, stream); }
      if (st) return st;
    }
    { ProfScope ps(PROF_ATTENTION, stream); st = attention_launch(q, k, vt, ao, save ? save->lse(l, NL) : nullptr, n_tiles, tw->heads, L.T, L.seq_pad, L.hd, L.hd_pad, scale, stream); }
    if (st) return st;
    {
      GemmArgs a{};
      a.M = M; a.N = D; a.K = D;
      a.bias = w.out_b;
      if (delta) {   // branch in bf16 over xb, LayerNorm-2 input to xn, the fp32 stream is left to the fc2 epilogue
        a.out = xb; a.ldo = D; a.aux16 = static_cast<const __nv_bfloat16*>(xb);
        a.out2 = xn;
      } else {
        a.out = h_mid; a.ldo = D; a.aux = h_in;
        a.out2 = fold ? xn : nullptr;   // bf16 copy of the stream after the attention branch
      }
      if (part_slots > 0) { a.ln_part = part; a.ln_slots = part_slots; }
      { ProfScope ps(PROF_GEMM_OUT, stream); st = gemm_dispatch(ao, D, w.out_w, D, a, delta ? EPI_DELTA_BF16 : EPI_RESID_F32, 0, stream); }
      if (st) return st;
    }
    // x = x + fc2(gelu_tanh(fc1(LN2(x))))
    if (fold) st = row_stats(xn);
    else { ProfScope ps(PROF_LAYERNORM, stream); st = layernorm_launch(h_mid, w.ln2_gamma, w.ln2_beta, xn2, M, D, tw->ln_eps, stream); }
    if (st) return st;
    GemmArgs a1{}, a2{};
    a1.M = M; a1.N = I; a1.K = D;
    a1.bias = fold ? w.fc1_bf : w.fc1_b;
    if (fold) set_consumer_stats(a1, w.fc1_sf);
    a1.out = save ? save->act(l, NL) : h1; a1.ldo = I;   // training: the activation is kept too (fc2 weight gradient)
    a1.out2 = save ? save->u(l, NL) : nullptr;          // training: keep gelu'(u) for the GELU backward
    a2.M = M; a2.N = D; a2.K = I;
    a2.bias = w.fc2_b;
    a2.out = h_out; a2.ldo = D; a2.aux = fold ? h_in : h_mid;
    a2.out2 = fold ? xb : nullptr;   // bf16 copy of the stream entering the next layer (or the projector)
    if (delta) a2.aux16 = static_cast<const __nv_bfloat16*>(xb);   // + the attention branch out_proj left there
    if (part_slots > 0 && l + 1 < tw->num_layers) { a2.ln_part = part; a2.ln_slots = part_slots; }
    st = RADVLM_ERR_UNSUPPORTED_SHAPE;
    if (save == nullptr && mlp_chain) {   // fc1 and fc2 as ONE persistent kernel: the activation stays in L2
      ProfScope ps(PROF_GEMM_FC2, stream, 2);
      st = gemm_chain_dispatch(xn2, fold ? w.fc1_wf : w.fc1_w, w.fc2_w, a1, EPI_GELU_TANH_BF16, a2, EPI_RESID_F32,
                               ws + L.off_ready, stream);
    }
    if (st == RADVLM_ERR_UNSUPPORTED_SHAPE) {
      { ProfScope ps(PROF_GEMM_FC1, stream); st = gemm_dispatch(xn2, D, fold ? w.fc1_wf : w.fc1_w, D, a1, save ? EPI_GELU_TANH_DUAL_BF16 : EPI_GELU_TANH_BF16, 0, stream); }
      if (st) return st;
      { ProfScope ps(PROF_GEMM_FC2, stream); st = gemm_dispatch(save ? save->act(l, NL) : h1, I, w.fc2_w, I, a2, EPI_RESID_F32, 0, stream); }
    }
    if (st) return st;
  }
  return RADVLM_OK;
}

// hidden_bf16 != nullptr: the bf16 copy of `hidden` already exists (written by the tower's last epilogue)
// ready != nullptr (rows / 256 unsigned ints of scratch): both GEMMs run as ONE persistent kernel (gemm4_sm100.cuh)
static int projector_forward_impl(const radvlm_projector_weights* pw, const float* hidden, int rows,
                                  void* out, int out_dtype, void* xn, void* h1, cudaStream_t stream,
                                  const void* hidden_bf16 = nullptr, void* ready = nullptr) {
  RV_CHECK_ARG(out_dtype == RADVLM_DT_BF16 || out_dtype == RADVLM_DT_F32 || out_dtype == RADVLM_DT_F16,
               "projector: out_dtype must be bf16, f16 or f32");
  RV_CHECK_ARG((pw->in_dim % 8) == 0 && (pw->hidden % 8) == 0, "projector: dims must be multiples of 8");
  int st;
  if (hidden_bf16 == nullptr) {
    { ProfScope ps(PROF_MISC, stream); st = cast_f32_bf16_launch(hidden, xn, static_cast<size_t>(rows) * pw->in_dim, stream); }
    if (st) return st;
    hidden_bf16 = xn;
  }
  if (ready != nullptr) {
    { ProfScope ps(PROF_GEMM, stream, 2);
      st = projector_chain_dispatch(hidden_bf16, pw->w1, pw->b1, h1, pw->w2, pw->b2, out, out_dtype, rows, pw->in_dim,
                                    pw->hidden, ready, stream); }
    if (st != RADVLM_ERR_UNSUPPORTED_SHAPE) return st;
  }
  {
    GemmArgs a{};
    a.M = rows; a.N = pw->hidden; a.K = pw->in_dim;
    a.bias = pw->b1;
    a.out = h1; a.ldo = pw->hidden;
    { ProfScope ps(PROF_GEMM, stream); st = gemm_dispatch(hidden_bf16, pw->in_dim, pw->w1, pw->in_dim, a, EPI_GELU_ERF_BF16, 0, stream); }
    if (st) return st;
  }
  {
    GemmArgs a{};
    a.M = rows; a.N = pw->hidden; a.K = pw->hidden;
    a.bias = pw->b2;
    a.out = out; a.ldo = pw->hidden;
    { ProfScope ps(PROF_GEMM, stream); st = gemm_dispatch(h1, pw->hidden, pw->w2, pw->hidden, a,
                       out_dtype == RADVLM_DT_BF16 ? EPI_BIAS_BF16 : (out_dtype == RADVLM_DT_F16 ? EPI_BIAS_F16 : EPI_BIAS_F32),
                       0, stream); }
    if (st) return st;
  }
  return RADVLM_OK;
}


// =================================================================================================
// Backward (training mode): gradients of the tower / projector parameters.  Weight gradients are fp32 and are
// ACCUMULATED (+=) with atomics, so micro-batches and tile chunks add up; a null gradient pointer skips that
// parameter (frozen, train.py:1642-1665 mm_tunable_parts).
// =================================================================================================
struct BackwardLayout {
  size_t off_xn1, off_g, off_dx, off_da, off_dqkv, off_attn, off_stats, total;
  size_t attn_bytes;
};

// Nothing of the forward is recomputed (every operand of the backward GEMMs was kept by the forward), so the workspace
// only holds gradients in flight.
static void make_bwd_layout(const radvlm_siglip_weights* tw, const EncodeLayout& L, int n_tiles, BackwardLayout* B) {
  const size_t M = L.M, D = tw->hidden, I = tw->intermediate;
  const size_t xn_cols = D > static_cast<size_t>(tw->patch_k_pad) ? D : tw->patch_k_pad;
  size_t off = 0;
  B->off_xn1 = off;  off = align_up(off + M * xn_cols * 2, 1024);  // im2col rows (patch-embedding weight gradient)
  B->off_g = off;    off = align_up(off + M * D * 2, 1024);        // bf16 copy of the residual gradient
  B->off_dx = off;   off = align_up(off + M * D * 2, 1024);        // dgrad outputs of width D
  B->off_da = off;   off = align_up(off + M * I * 2, 1024);        // dL/du (fc2 data gradient x gelu')
  B->off_dqkv = off; off = align_up(off + M * 3 * D * 2, 1024);    // [dQ | dK | dV]
  B->attn_bytes = attention_bwd_workspace_bytes(n_tiles, tw->heads, L.seq_pad);
  B->off_attn = off; off = align_up(off + B->attn_bytes, 1024);
  B->off_stats = off; off = align_up(off + M * 8 + 2 * D * sizeof(float), 1024);  // LayerNorm row stats + frozen-affine scratch
  B->total = off;
}

// Weight-gradient GEMMs have few output tiles (25..85 for 74 CTA pairs): the K range (= the rows of the batch) is cut
// into `splits` partial products so that tiles * splits items fill whole waves of CTA pairs.  Chosen to maximise
// items / (waves * pairs) with a small charge per extra split (one more atomic epilogue per tile), keeping at least
// 12 K slabs per item.
static int pick_splits(int M, int N, int K) {
  const int pairs = (device_sm_count() > 0 ? device_sm_count() : 148) / 2;
  const int bn = gemm_bmn_block_n(N);
  const int tiles = ((M + 255) / 256) * ((N + bn - 1) / bn);
  const int slabs = (K + 63) / 64;
  int best = 1;
  double best_score = -1.0;
  for (int sp = 1; sp <= 32 && (sp == 1 || slabs / sp >= 12); ++sp) {
    const int items = tiles * sp;
    const int waves = (items + pairs - 1) / pairs;
    const double score = static_cast<double>(items) / (static_cast<double>(waves) * pairs) - 0.004 * sp;
    if (score > best_score) { best_score = score; best = sp; }
  }
  return best;
}

// dW[out, in] += dY^T[out, rows] X[rows, in];  db[out] += colsum(dY)
static int linear_wgrad(const void* dY, int ldy, const void* X, int ldx, int rows, int out_dim, int in_dim, float* dW,
                        int ldw, float* db, cudaStream_t stream) {
  int st = RADVLM_OK;
  if (dW != nullptr) {
    GemmArgs a{};
    a.M = out_dim; a.N = in_dim; a.K = rows;
    a.out = dW; a.ldo = ldw;
    a.a_mn = 1; a.b_mn = 1; a.k_splits = pick_splits(out_dim, in_dim, rows);
    { ProfScope ps(PROF_BWD_WGRAD, stream); st = gemm_dispatch(dY, ldy, X, ldx, a, EPI_ATOMIC_F32, 0, stream); }
    if (st) return st;
  }
  if (db != nullptr) { ProfScope ps(PROF_BWD_WGRAD, stream); st = colsum_bf16_launch(dY, rows, out_dim, ldy, db, stream); }
  return st;
}

// dX[rows, in] = dY[rows, out] W[out, in]   (W read as stored)
// gelu_u != nullptr (bf16 [rows, ldx], the gelu'(u) the forward kept): dX is multiplied by it in the epilogue
// (fc2 data gradient -> dL/du)
static int linear_dgrad(const void* dY, int ldy, const void* W, int ldw, int rows, int out_dim, int in_dim, void* dX,
                        int ldx, bool f32_out, cudaStream_t stream, const void* gelu_u = nullptr) {
  GemmArgs a{};
  a.M = rows; a.N = in_dim; a.K = out_dim;
  a.out = dX; a.ldo = ldx;
  a.out2 = const_cast<void*>(gelu_u);
  a.b_mn = 1;
  ProfScope ps(PROF_BWD_DGRAD, stream);
  return gemm_dispatch(dY, ldy, W, ldw, a, gelu_u ? EPI_MUL_BF16 : (f32_out ? EPI_BIAS_F32 : EPI_BIAS_BF16), 0, stream);
}

// Layers [layer_lo, layer_hi) are processed from the top down; the embeddings follow when layer_lo == 0.  Calling it
// range by range (top ranges first, same d_hidden buffer) lets the host overlap the gradient all-reduce of the
// finished layers with the backward of the next ones.
static int tower_backward_impl(const radvlm_siglip_weights* tw, const radvlm_siglip_grads* gr, const void* pixels,
                               int pixel_dtype, int n_tiles, const TowerSaved& sv, float* dh, const EncodeLayout& L,
                               const BackwardLayout& B, uint8_t* ws, cudaStream_t stream, int layer_lo, int layer_hi) {
  const int D = tw->hidden, I = tw->intermediate, NL = tw->num_layers;
  const int M = static_cast<int>(L.M);
  void* xn_scratch = ws + B.off_xn1;   // im2col rows of the patch-embedding weight gradient
  void* g = ws + B.off_g;
  void* dx = ws + B.off_dx;
  void* da = ws + B.off_da;
  void* dqkv = ws + B.off_dqkv;
  void* attn_ws = ws + B.off_attn;
  void* stats = ws + B.off_stats;
  float* ln_scratch = reinterpret_cast<float*>(ws + B.off_stats + align_up(static_cast<size_t>(M) * 8, 16));
  int st;
  const float scale = 1.0f / sqrtf(static_cast<float>(L.hd));
  const size_t MD = static_cast<size_t>(M) * D;

  bool g_is_dh = false;  // g == bf16(dh)?  (false on entry: dh comes from the caller)
  for (int l = layer_hi - 1; l >= layer_lo; --l) {
    const radvlm_vit_layer_weights& w = tw->layers[l];
    const radvlm_vit_layer_grads* lg = (gr != nullptr && gr->layers != nullptr) ? &gr->layers[l] : nullptr;
    auto G = [&](float* radvlm_vit_layer_grads::*m) -> float* { return lg ? lg->*m : nullptr; };
    const float* h0 = sv.h(l);
    const float* h1 = sv.hmid(l, NL);   // saved by the forward: no out_proj recompute
    void* q = sv.qkv(l, 0, NL);         // ... and no QKV recompute
    void* k = sv.qkv(l, 1, NL);
    void* vt = sv.qkv(l, 2, NL);
    const void* u = sv.u(l, NL);        // ... and no fc1 recompute
    const void* xn1 = sv.xn(l, 0, NL);  // LayerNorm outputs: kept as well, nothing of the forward is recomputed
    const void* xn2 = sv.xn(l, 1, NL);
    const void* ao = sv.ao(l, NL);
    // the backward attention wants plain zeros in the V padding (no ones column): clear it in the saved V
    { ProfScope ps_re(PROF_BWD_RECOMPUTE, stream);
      if ((st = qkv_pad_prepare_launch(nullptr, nullptr, vt, n_tiles, tw->heads, L.T, L.seq_pad, L.hd, L.hd_pad, 0.0f, stream))) return st; }
    // ---- MLP branch: h2 = h1 + fc2(gelu(fc1(LN2(h1))))
    if (!g_is_dh) { ProfScope ps(PROF_BWD_ELEMENTWISE, stream); if ((st = cast_f32_bf16_launch(dh, g, MD, stream))) return st; }
    // dL/du = (g W2) o gelu'(u): one multiply in the epilogue of the data-gradient GEMM; gelu'(u) and a = gelu(u) were kept
    if ((st = linear_dgrad(g, D, w.fc2_w, I, M, D, I, da, I, false, stream, u))) return st;
    const void* act = sv.act(l, NL);
    if ((st = linear_wgrad(g, D, act, I, M, D, I, G(&radvlm_vit_layer_grads::fc2_w), I, G(&radvlm_vit_layer_grads::fc2_b), stream))) return st;
    if ((st = linear_wgrad(da, I, xn2, D, M, I, D, G(&radvlm_vit_layer_grads::fc1_w), D, G(&radvlm_vit_layer_grads::fc1_b), stream))) return st;
    if ((st = linear_dgrad(da, I, w.fc1_w, D, M, I, D, dx, D, false, stream))) return st;           // dL/dLN2
    {
      float* dg = G(&radvlm_vit_layer_grads::ln2_gamma);
      float* db = G(&radvlm_vit_layer_grads::ln2_beta);
      if ((dg == nullptr) != (db == nullptr)) {  // one of the pair frozen: the other's sum goes to scratch
        RV_CUDA(cudaMemsetAsync(ln_scratch, 0, 2 * D * sizeof(float), stream));
        if (!dg) dg = ln_scratch; else db = ln_scratch + D;
      }
      ProfScope ps(PROF_BWD_ELEMENTWISE, stream, 2);
      if ((st = layernorm_bwd_launch(h1, w.ln2_gamma, dx, dh, dg, db, stats, M, D, tw->ln_eps, stream, g))) return st;  // also g = bf16(dh)
    }
    // ---- attention branch: h1 = h0 + out_proj(attn(LN1(h0)))   (g already holds bf16(dh): written by the LayerNorm backward)
    if ((st = linear_wgrad(g, D, ao, D, M, D, D, G(&radvlm_vit_layer_grads::out_w), D, G(&radvlm_vit_layer_grads::out_b), stream))) return st;
    if ((st = linear_dgrad(g, D, w.out_w, D, M, D, D, dx, D, false, stream))) return st;            // dL/d(attn out)
    { ProfScope ps(PROF_BWD_ATTENTION, stream, 4);
    if ((st = attention_bwd_launch(q, k, vt, dx, ao, sv.lse(l, NL), dqkv, attn_ws, B.attn_bytes, n_tiles, tw->heads,
                                   L.T, L.seq_pad, L.hd, L.hd_pad, scale, stream))) return st; }
    if ((st = linear_wgrad(dqkv, 3 * D, xn1, D, M, 3 * D, D, G(&radvlm_vit_layer_grads::qkv_w), D, G(&radvlm_vit_layer_grads::qkv_b), stream))) return st;
    if ((st = linear_dgrad(dqkv, 3 * D, w.qkv_w, D, M, 3 * D, D, dx, D, false, stream))) return st;  // dL/dLN1
    {
      float* dg = G(&radvlm_vit_layer_grads::ln1_gamma);
      float* db = G(&radvlm_vit_layer_grads::ln1_beta);
      if ((dg == nullptr) != (db == nullptr)) {
        RV_CUDA(cudaMemsetAsync(ln_scratch, 0, 2 * D * sizeof(float), stream));
        if (!dg) dg = ln_scratch; else db = ln_scratch + D;
      }
      ProfScope ps(PROF_BWD_ELEMENTWISE, stream, 2);
      if ((st = layernorm_bwd_launch(h0, w.ln1_gamma, dx, dh, dg, db, stats, M, D, tw->ln_eps, stream, g))) return st;
      g_is_dh = true;   // the next (lower) layer and the embeddings start from this bf16 copy
    }
  }
  // ---- embeddings: hidden0 = im2col(pixels) Wp^T + bp + pos  (siglip_encoder.py:169-174)
  if (layer_lo == 0 && gr != nullptr && (gr->patch_w || gr->patch_b || gr->pos_embed)) {
    if (gr->pos_embed && (st = pos_embed_grad_launch(dh, gr->pos_embed, n_tiles, L.T, D, stream))) return st;
    if (gr->patch_w || gr->patch_b) {
      if (!g_is_dh && (st = cast_f32_bf16_launch(dh, g, MD, stream))) return st;
      if (gr->patch_w && (st = im2col_launch(pixels, pixel_dtype, xn_scratch, n_tiles, tw->channels, tw->image_size,
                                             tw->patch_size, tw->patch_k_pad, stream))) return st;
      if ((st = linear_wgrad(g, D, xn_scratch, tw->patch_k_pad, M, D, tw->patch_k_pad, gr->patch_w, tw->patch_k_pad, gr->patch_b, stream))) return st;
    }
  }
  return RADVLM_OK;
}

// projector: y = W2 gelu_erf(W1 x + b1) + b2  (builder.py:41-48); x = bf16(hidden)
static int projector_backward_impl(const radvlm_projector_weights* pw, const radvlm_projector_grads* gr,
                                   const float* hidden, const void* dfeat /* bf16 [rows, P] */, int rows, float* d_hidden,
                                   uint8_t* ws, cudaStream_t stream) {
  const int Din = pw->in_dim, P = pw->hidden;
  const size_t x_bytes = align_up(static_cast<size_t>(rows) * Din * 2, 1024);
  const size_t p_bytes = align_up(static_cast<size_t>(rows) * P * 2, 1024);
  void* x = ws;
  void* u = ws + x_bytes;
  void* da = ws + x_bytes + p_bytes;
  void* act = ws + x_bytes + 2 * p_bytes;
  int st;
  if ((st = cast_f32_bf16_launch(hidden, x, static_cast<size_t>(rows) * Din, stream))) return st;
  {
    GemmArgs a{};
    a.M = rows; a.N = P; a.K = Din;
    a.bias = pw->b1;
    a.out = u; a.ldo = P;
    if ((st = gemm_dispatch(x, Din, pw->w1, Din, a, EPI_BIAS_BF16, 0, stream))) return st;
  }
  if ((st = linear_dgrad(dfeat, P, pw->w2, P, rows, P, P, da, P, false, stream))) return st;
  if ((st = gelu_fwd_bwd_launch(u, da, act, static_cast<size_t>(rows) * P, 1, stream))) return st;
  if ((st = linear_wgrad(dfeat, P, act, P, rows, P, P, gr ? gr->w2 : nullptr, P, gr ? gr->b2 : nullptr, stream))) return st;
  if ((st = linear_wgrad(da, P, x, Din, rows, P, Din, gr ? gr->w1 : nullptr, Din, gr ? gr->b1 : nullptr, stream))) return st;
  if (d_hidden != nullptr && (st = linear_dgrad(da, P, pw->w1, Din, rows, P, Din, d_hidden, Din, true, stream))) return st;
  return RADVLM_OK;
}

}  // namespace rv

using namespace rv;

extern "C" size_t radvlm_encode_workspace_bytes(const radvlm_siglip_weights* tw,
                                                const radvlm_projector_weights* pw, int n_tiles) {
  EncodeLayout L;
  if (make_layout(tw, pw, n_tiles, &L) != RADVLM_OK) return 0;
  return L.total;
}

extern "C" int radvlm_siglip_tower_forward(const radvlm_siglip_weights* tw, const void* pixels,
                                           int pixel_dtype, int n_tiles, float* hidden_out,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  int st = require_sm100();
  if (st) return st;
  EncodeLayout L;
  st = make_layout(tw, nullptr, n_tiles, &L);
  if (st) return st;
  RV_CHECK_ARG(pixels && hidden_out && workspace, "tower: null pointer");
  if (workspace_bytes < L.total) {
    set_error("tower: workspace too small (%zu < %zu)", workspace_bytes, L.total);
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  return tower_forward_impl(tw, pixels, pixel_dtype, n_tiles, hidden_out, L,
                            static_cast<uint8_t*>(workspace), static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_projector_forward(const radvlm_projector_weights* pw, const float* hidden, int rows,
                                        void* features_out, int out_dtype, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  int st = require_sm100();
  if (st) return st;
  RV_CHECK_ARG(pw && hidden && features_out && workspace && rows > 0, "projector: bad arguments");
  const size_t xn_bytes = align_up(static_cast<size_t>(rows) * pw->in_dim * 2, 1024);
  const size_t h1_bytes = align_up(static_cast<size_t>(rows) * pw->hidden * 2, 1024);
  if (workspace_bytes < xn_bytes + h1_bytes) {
    set_error("projector: workspace too small (%zu < %zu)", workspace_bytes, xn_bytes + h1_bytes);
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  // the fused two-GEMM kernel needs rows / 256 counters: taken from the tail of the workspace when the caller left room
  const size_t ready_bytes = align_up((static_cast<size_t>(rows) / 256 + 2) * 4, 1024);
  void* ready = workspace_bytes >= xn_bytes + h1_bytes + ready_bytes ? ws + xn_bytes + h1_bytes : nullptr;
  return projector_forward_impl(pw, hidden, rows, features_out, out_dtype, ws, ws + xn_bytes,
                                static_cast<cudaStream_t>(stream), nullptr, ready);
}

extern "C" int radvlm_encode_images(const radvlm_siglip_weights* tw, const radvlm_projector_weights* pw,
                                    const void* pixels, int pixel_dtype, int n_tiles, void* features_out,
                                    int out_dtype, void* workspace, size_t workspace_bytes, void* stream) {
  int st = require_sm100();
  if (st) return st;
  RV_CHECK_ARG(pw != nullptr, "encode: null projector weights");
  EncodeLayout L;
  st = make_layout(tw, pw, n_tiles, &L);
  if (st) return st;
  RV_CHECK_ARG(pixels && features_out && workspace, "encode: null pointer");
  RV_CHECK_ARG(pw->in_dim == tw->hidden, "encode: projector in_dim %d != tower hidden %d", pw->in_dim,
               tw->hidden);
  if (workspace_bytes < L.total) {
    set_error("encode: workspace too small (%zu < %zu)", workspace_bytes, L.total);
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* hidden = reinterpret_cast<float*>(ws + L.off_hidden);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const void* hidden_bf16 = nullptr;
  st = tower_forward_impl(tw, pixels, pixel_dtype, n_tiles, hidden, L, ws, s, nullptr, &hidden_bf16);
  if (st) return st;
  return projector_forward_impl(pw, hidden, static_cast<int>(L.M), features_out, out_dtype,
                                ws + L.off_xn, ws + L.off_h1, s, hidden_bf16, ws + L.off_q /* q is dead by now */);
}

// ------------------------------------------------------------------------------------------------ training mode
extern "C" size_t radvlm_tower_saved_bytes(const radvlm_siglip_weights* tw, int n_tiles) {
  EncodeLayout L;
  if (make_layout(tw, nullptr, n_tiles, &L) != RADVLM_OK) return 0;
  return make_saved(tw, L, n_tiles, nullptr).total;
}

extern "C" size_t radvlm_tower_saved_hidden_offset(const radvlm_siglip_weights* tw, int n_tiles) {
  EncodeLayout L;
  if (make_layout(tw, nullptr, n_tiles, &L) != RADVLM_OK) return 0;
  return make_saved(tw, L, n_tiles, nullptr).h_bytes * static_cast<size_t>(tw->num_layers);
}

extern "C" int radvlm_siglip_tower_forward_train(const radvlm_siglip_weights* tw, const void* pixels, int pixel_dtype,
                                                 int n_tiles, void* saved, size_t saved_bytes, void* workspace,
                                                 size_t workspace_bytes, void* stream) {
  int st = require_sm100();
  if (st) return st;
  EncodeLayout L;
  st = make_layout(tw, nullptr, n_tiles, &L);
  if (st) return st;
  RV_CHECK_ARG(pixels && saved && workspace, "tower_train: null pointer");
  const TowerSaved sv = make_saved(tw, L, n_tiles, saved);
  if (workspace_bytes < L.total || saved_bytes < sv.total) {
    set_error("tower_train: buffers too small (workspace %zu < %zu or saved %zu < %zu)", workspace_bytes, L.total,
              saved_bytes, sv.total);
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  return tower_forward_impl(tw, pixels, pixel_dtype, n_tiles, reinterpret_cast<float*>(ws + L.off_hidden), L, ws,
                            static_cast<cudaStream_t>(stream), &sv);
}

extern "C" size_t radvlm_tower_backward_workspace_bytes(const radvlm_siglip_weights* tw, int n_tiles) {
  EncodeLayout L;
  if (make_layout(tw, nullptr, n_tiles, &L) != RADVLM_OK) return 0;
  BackwardLayout B;
  make_bwd_layout(tw, L, n_tiles, &B);
  return B.total;
}

extern "C" int radvlm_siglip_tower_backward_range(const radvlm_siglip_weights* tw, const radvlm_siglip_grads* grads,
                                                  const void* pixels, int pixel_dtype, int n_tiles, const void* saved,
                                                  size_t saved_bytes, float* d_hidden, void* workspace,
                                                  size_t workspace_bytes, int layer_lo, int layer_hi, void* stream) {
  int st = require_sm100();
  if (st) return st;
  EncodeLayout L;
  st = make_layout(tw, nullptr, n_tiles, &L);
  if (st) return st;
  RV_CHECK_ARG(pixels && saved && d_hidden && workspace, "tower_backward: null pointer");
  RV_CHECK_ARG(0 <= layer_lo && layer_lo <= layer_hi && layer_hi <= tw->num_layers, "tower_backward: bad layer range [%d, %d)",
               layer_lo, layer_hi);
  const TowerSaved sv = make_saved(tw, L, n_tiles, const_cast<void*>(saved));
  BackwardLayout B;
  make_bwd_layout(tw, L, n_tiles, &B);
  if (workspace_bytes < B.total || saved_bytes < sv.total) {
    set_error("tower_backward: buffers too small (workspace %zu < %zu or saved %zu < %zu)", workspace_bytes, B.total,
              saved_bytes, sv.total);
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  return tower_backward_impl(tw, grads, pixels, pixel_dtype, n_tiles, sv, d_hidden, L, B,
                             static_cast<uint8_t*>(workspace), static_cast<cudaStream_t>(stream), layer_lo, layer_hi);
}

extern "C" int radvlm_siglip_tower_backward(const radvlm_siglip_weights* tw, const radvlm_siglip_grads* grads,
                                            const void* pixels, int pixel_dtype, int n_tiles, const void* saved,
                                            size_t saved_bytes, float* d_hidden, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  RV_CHECK_ARG(tw != nullptr, "tower_backward: null weights");
  return radvlm_siglip_tower_backward_range(tw, grads, pixels, pixel_dtype, n_tiles, saved, saved_bytes, d_hidden,
                                            workspace, workspace_bytes, 0, tw->num_layers, stream);
}

extern "C" size_t radvlm_projector_backward_workspace_bytes(const radvlm_projector_weights* pw, int rows) {
  if (pw == nullptr || rows <= 0) return 0;
  return align_up(static_cast<size_t>(rows) * pw->in_dim * 2, 1024) + 3 * align_up(static_cast<size_t>(rows) * pw->hidden * 2, 1024);
}

extern "C" int radvlm_projector_backward(const radvlm_projector_weights* pw, const radvlm_projector_grads* grads,
                                         const float* hidden, const void* d_features, int rows, float* d_hidden,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  int st = require_sm100();
  if (st) return st;
  RV_CHECK_ARG(pw && hidden && d_features && workspace && rows > 0, "projector_backward: bad arguments");
  RV_CHECK_ARG((pw->in_dim % 8) == 0 && (pw->hidden % 8) == 0, "projector_backward: dims must be multiples of 8");
  if (workspace_bytes < radvlm_projector_backward_workspace_bytes(pw, rows)) {
    set_error("projector_backward: workspace too small");
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  return projector_backward_impl(pw, grads, hidden, d_features, rows, d_hidden, static_cast<uint8_t*>(workspace),
                                 static_cast<cudaStream_t>(stream));
}
