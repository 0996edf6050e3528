// Host orchestration of the encode path: SigLIP tower and mlp2x_gelu projector as a fixed sequence of
// kernel launches on one stream (no allocation, no synchronisation -> CUDA-graph capturable).
//
//   SigLipVisionTower.forward        siglip_encoder.py:576-589
//   SigLipVisionEmbeddings.forward   siglip_encoder.py:169-174
//   SigLipEncoderLayer.forward       siglip_encoder.py:285-305
//   encode_images / mm_projector     llava_arch.py:192-196 ; multimodal_projector/builder.py:41-48
#include "host_util.h"
#include "internal.h"

namespace rv {


struct EncodeLayout {
  int P, T, seq_pad, hd, hd_pad;
  size_t M;
  size_t off_hidden, off_xn, off_q, off_k, off_vt, off_h1, total;
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
  L->off_q = off;      off = align_up(off + L->qkv_bytes, 1024);
  L->off_k = off;      off = align_up(off + L->qkv_bytes, 1024);
  L->off_vt = off;     off = align_up(off + L->qkv_bytes, 1024);
  L->off_h1 = off;     off = align_up(off + L->M * h1_cols * 2, 1024);
  L->total = off;
  return RADVLM_OK;
}

static int tower_forward_impl(const radvlm_siglip_weights* tw, const void* pixels, int pixel_dtype,
                              int n_tiles, float* hidden, const EncodeLayout& L, uint8_t* ws,
                              cudaStream_t stream) {
  const int D = tw->hidden, I = tw->intermediate;
  const int M = static_cast<int>(L.M);
  void* xn = ws + L.off_xn;
  void* q = ws + L.off_q;
  void* k = ws + L.off_k;
  void* vt = ws + L.off_vt;
  void* h1 = ws + L.off_h1;
  int st;

  // padding of q/k/vt must be zero (never written by the QKV epilogue)
  {
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
    a.out = hidden; a.ldo = D;
    a.aux = tw->pos_embed; a.aux_period = L.T;
    { ProfScope ps(PROF_GEMM, stream); st = gemm_dispatch(xn, tw->patch_k_pad, tw->patch_w, tw->patch_k_pad, a, EPI_POS_F32, 0, stream); }
    if (st) return st;
  }

  const float scale = 1.0f / sqrtf(static_cast<float>(L.hd));
  for (int l = 0; l < tw->num_layers; ++l) {
    const radvlm_vit_layer_weights& w = tw->layers[l];
    // x = x + out_proj(attn(LN1(x)))
    { ProfScope ps(PROF_LAYERNORM, stream); st = layernorm_launch(hidden, w.ln1_gamma, w.ln1_beta, xn, M, D, tw->ln_eps, stream); }
    if (st) return st;
    {
      GemmArgs a{};
      a.M = M; a.N = 3 * D; a.K = D;
      a.bias = w.qkv_b;
      a.q = static_cast<__nv_bfloat16*>(q);
      a.k = static_cast<__nv_bfloat16*>(k);
      a.vt = static_cast<__nv_bfloat16*>(vt);
      a.seq = L.T; a.seq_pad = L.seq_pad; a.heads = tw->heads; a.hd = L.hd; a.hd_pad = L.hd_pad;
      { ProfScope ps(PROF_GEMM_QKV, stream); st = gemm_dispatch(xn, D, w.qkv_w, D, a, EPI_QKV_SPLIT, 0, stream); }
      if (st) return st;
    }
    { ProfScope ps(PROF_ATTENTION, stream); st = attention_launch(q, k, vt, xn, nullptr, n_tiles, tw->heads, L.T, L.seq_pad, L.hd, L.hd_pad, scale, stream); }
    if (st) return st;
    {
      GemmArgs a{};
      a.M = M; a.N = D; a.K = D;
      a.bias = w.out_b;
      a.out = hidden; a.ldo = D; a.aux = hidden;
      { ProfScope ps(PROF_GEMM_OUT, stream); st = gemm_dispatch(xn, D, w.out_w, D, a, EPI_RESID_F32, 0, stream); }
      if (st) return st;
    }
    // x = x + fc2(gelu_tanh(fc1(LN2(x))))
    { ProfScope ps(PROF_LAYERNORM, stream); st = layernorm_launch(hidden, w.ln2_gamma, w.ln2_beta, xn, M, D, tw->ln_eps, stream); }
    if (st) return st;
    {
      GemmArgs a{};
      a.M = M; a.N = I; a.K = D;
      a.bias = w.fc1_b;
      a.out = h1; a.ldo = I;
      { ProfScope ps(PROF_GEMM_FC1, stream); st = gemm_dispatch(xn, D, w.fc1_w, D, a, EPI_GELU_TANH_BF16, 0, stream); }
      if (st) return st;
    }
    {
      GemmArgs a{};
      a.M = M; a.N = D; a.K = I;
      a.bias = w.fc2_b;
      a.out = hidden; a.ldo = D; a.aux = hidden;
      { ProfScope ps(PROF_GEMM_FC2, stream); st = gemm_dispatch(h1, I, w.fc2_w, I, a, EPI_RESID_F32, 0, stream); }
      if (st) return st;
    }
  }
  return RADVLM_OK;
}

static int projector_forward_impl(const radvlm_projector_weights* pw, const float* hidden, int rows,
                                  void* out, int out_dtype, void* xn, void* h1, cudaStream_t stream) {
  RV_CHECK_ARG(out_dtype == RADVLM_DT_BF16 || out_dtype == RADVLM_DT_F32,
               "projector: out_dtype must be bf16 or f32");
  RV_CHECK_ARG((pw->in_dim % 8) == 0 && (pw->hidden % 8) == 0, "projector: dims must be multiples of 8");
  int st;
  { ProfScope ps(PROF_MISC, stream); st = cast_f32_bf16_launch(hidden, xn, static_cast<size_t>(rows) * pw->in_dim, stream); }
  if (st) return st;
  {
    GemmArgs a{};
    a.M = rows; a.N = pw->hidden; a.K = pw->in_dim;
    a.bias = pw->b1;
    a.out = h1; a.ldo = pw->hidden;
    { ProfScope ps(PROF_GEMM, stream); st = gemm_dispatch(xn, pw->in_dim, pw->w1, pw->in_dim, a, EPI_GELU_ERF_BF16, 0, stream); }
    if (st) return st;
  }
  {
    GemmArgs a{};
    a.M = rows; a.N = pw->hidden; a.K = pw->hidden;
    a.bias = pw->b2;
    a.out = out; a.ldo = pw->hidden;
    { ProfScope ps(PROF_GEMM, stream); st = gemm_dispatch(h1, pw->hidden, pw->w2, pw->hidden, a,
                       out_dtype == RADVLM_DT_BF16 ? EPI_BIAS_BF16 : EPI_BIAS_F32, 0, stream); }
    if (st) return st;
  }
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
  return projector_forward_impl(pw, hidden, rows, features_out, out_dtype, ws, ws + xn_bytes,
                                static_cast<cudaStream_t>(stream));
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
  st = tower_forward_impl(tw, pixels, pixel_dtype, n_tiles, hidden, L, ws, s);
  if (st) return st;
  return projector_forward_impl(pw, hidden, static_cast<int>(L.M), features_out, out_dtype,
                                ws + L.off_xn, ws + L.off_h1, s);
}
