/*
 * radvlm_b200.h — C ABI of the B200-native RadVLM multimodal encode path.
 *
 * The reference (rfahrn/RadVLM, finetuning/llava) is pure Python and has no FFI; this header is the
 * boundary a maintainer binds with ctypes.  Each entry point cites the
 * reference code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - plain pointers and sizes only; device pointers are raw CUDA device addresses owned by the caller
 *   - every GPU entry point takes a cudaStream_t (passed as void*), never allocates device memory,
 *     never synchronises, and returns an int status (0 = OK)
 *   - scratch memory is caller-provided; each entry point that needs it has a *_workspace_bytes query
 *   - there is NO CPU fallback: a device that is not sm_100 yields RADVLM_ERR_UNSUPPORTED_DEVICE
 *   - radvlm_last_error() returns a thread-local, human readable message for the last failure
 */
#ifndef RADVLM_B200_H_
#define RADVLM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADVLM_OK 0
#define RADVLM_ERR_BAD_ARGUMENT 1
#define RADVLM_ERR_UNSUPPORTED_SHAPE 2
#define RADVLM_ERR_CUDA 3
#define RADVLM_ERR_WORKSPACE_TOO_SMALL 4
#define RADVLM_ERR_UNSUPPORTED_DEVICE 5

/* dtype codes for pixel / feature buffers */
#define RADVLM_DT_F32 0
#define RADVLM_DT_BF16 1
#define RADVLM_DT_F16 2

const char* radvlm_last_error(void);
int radvlm_abi_version(void);

/* Optional device timing per kernel class (CUDA events on the launch stream; off by default).
 * classes: 0 gemm (patch embed + projector), 1 attention, 2 layernorm, 3 misc (im2col/cast/memset),
 * 4 preprocess, 5 merge_splice, 6 gemm qkv, 7 gemm out_proj, 8 gemm fc1, 9 gemm fc2.
 * radvlm_profile_read waits for the recorded events, returns summed milliseconds and kernel-launch
 * counts per class, and clears the record list. */
#define RADVLM_PROF_NUM_CLASSES 10
int radvlm_profile_enable(int on);
int radvlm_profile_read(float* ms_per_class, int64_t* launches_per_class, int n_classes);

/* TMA descriptors (CUtensorMap) are cached per calling thread, keyed on the full argument tuple of
 * cuTensorMapEncodeTiled (base address, dims, pitches, box, swizzle): a step re-uses the descriptors of its
 * workspace / weight buffers instead of re-encoding ~600 of them.  Returns the calling thread's hit / miss counts
 * (host only, no device needed).  RADVLM_B200_TMAP_CACHE=0 disables the cache. */
int radvlm_tmap_cache_stats(uint64_t* hits, uint64_t* misses);

/* ------------------------------------------------------------------------------------------------
 * GEMM building block:  out[M,N] = A[M,K](bf16) * W[N,K]^T(bf16)  (+ fused epilogue), fp32 accumulate
 * in TMEM (tcgen05.mma, TMA-fed).  Replaces nn.Linear.forward at
 *   siglip_encoder.py:207-209 (q/k/v_proj), :237 (out_proj), :252-254 (fc1/act/fc2)
 *   multimodal_projector/builder.py:44-48 (mlp2x_gelu)
 * epilogue: one of RADVLM_EPI_*.   lda/ldw/ldo are row pitches in ELEMENTS.
 * ---------------------------------------------------------------------------------------------- */
#define RADVLM_EPI_BIAS_BF16 0      /* out bf16 = acc + bias */
#define RADVLM_EPI_GELU_TANH_BF16 1 /* out bf16 = gelu_tanh(acc + bias)      siglip_encoder.py:83,247,253 */
#define RADVLM_EPI_GELU_ERF_BF16 2  /* out bf16 = gelu_erf(acc + bias)       builder.py:46 */
#define RADVLM_EPI_RESID_F32 3      /* out f32  = acc + bias + aux[M,ldo]    siglip_encoder.py:293,298 */
#define RADVLM_EPI_POS_F32 4        /* out f32  = acc + bias + aux[row % aux_period, N]  siglip_encoder.py:173 */
#define RADVLM_EPI_BIAS_F32 6       /* out f32  = acc + bias */
#define RADVLM_EPI_ATOMIC_F32 7     /* out f32 += acc (atomic; split-K partial sums, gradient accumulation) */
#define RADVLM_EPI_BIAS_F16 9       /* out f16  = acc + bias                 (fp16 serving: model_worker.py:124-127) */

int radvlm_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     const float* bias, int epilogue, void* out, int64_t ldo, const float* aux,
                     int aux_period, int block_n /* 0 = auto, else 128|192|256 */, void* stream);

/* Backward-pass form of the same GEMM: the operands are read as autograd leaves them, without transposes.
 *   a_layout 0: A is [M, K] row-major            1: A is stored as [K, M] row-major (e.g. dY for a weight gradient)
 *   b_layout 0: W is [N, K] row-major            1: W is stored as [K, N] row-major (e.g. the Linear weight [out, in]
 *                                                   when the contraction runs over `out`: dX = dY W)
 *   k_splits > 1 cuts the K range into partial products that are summed with atomics (RADVLM_EPI_ATOMIC_F32 only:
 *   out must hold the running sum, e.g. zeros or the gradient accumulated so far).
 * nn.Linear backward (torch/nn/functional.linear autograd, used by every Linear of siglip_encoder.py / builder.py):
 *   dX[M, in]    = dY[M, out] W[out, in]       -> a_layout 0, b_layout 1, K = out
 *   dW[out, in] += dY^T[out, M] X[M, in]       -> a_layout 1, b_layout 1, K = M, RADVLM_EPI_ATOMIC_F32 */
int radvlm_gemm_bf16_ex(const void* A, int64_t lda, int a_layout, const void* W, int64_t ldw, int b_layout, int M, int N,
                        int K, const float* bias, int epilogue, void* out, int64_t ldo, const float* aux,
                        int aux_period, int k_splits, void* stream);

/* The same GEMM with a LayerNorm folded in (siglip_encoder.py:264,266,287,296 followed by a Linear):
 *   A = bf16(x) (not normalised), W = bf16(gamma o W0), ln_s[n] = sum_k W[n, k], bias = b0 + W0 beta,
 *   ln_stats[row] = (mean, rstd) of row `row` of x (radvlm_ln_row_stats_bf16):
 *   out = epilogue( rstd * (acc - mean * ln_s[n]) + bias[n] )  ==  epilogue( LayerNorm(x) W0^T + b0 ).
 * epilogue: RADVLM_EPI_BIAS_BF16 | RADVLM_EPI_GELU_TANH_BF16 | RADVLM_EPI_GELU_ERF_BF16; N % 8 == 0. */
int radvlm_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, const float* bias,
                        const void* ln_stats, const float* ln_s, int epilogue, void* out, int64_t ldo, void* stream);

/* Tile-shape policy of the GEMM: 0 = auto (CTA pairs, tcgen05 cta_group::2, 256 x BN tiles when M >= 512),
 * 1 = force single-CTA 128 x BN tiles, 2 = force CTA-pair tiles.  Process-wide; meant for tests / tuning. */
int radvlm_gemm_set_mode(int mode);

/* Host-only (no CUDA): the tile schedule of the scheduled GEMM kernel for an [M, N] output on `pairs` CTA pairs (74 on
 * B200).  A row block of 256 rows is cut into 256-wide column tiles plus ONE 128-wide tile when its remainder is <= 128
 * columns; tiles are dealt to the pairs by list scheduling (cost 100 / 82).  tiles_per_pair[pairs] (or NULL), *n_tiles,
 * *max_load, *min_load describe the result; RADVLM_ERR_UNSUPPORTED_SHAPE when the kernel does not cover the shape. */
int radvlm_gemm_schedule_stats(int M, int N, int pairs, int* tiles_per_pair, int* n_tiles, int* max_load, int* min_load);

/* QKV projection with the head-split scatter fused into the epilogue
 * (siglip_encoder.py:207-213: three Linear + view/transpose).  W is the row-concatenation
 * [q_proj; k_proj; v_proj] = [3*heads*hd, K]; bias likewise.  Outputs (bf16):
 *   q, k, v : bf16 [tiles, heads, seq_pad, hd_pad]  (the `vt` argument is V in the SAME layout as q / k: the
 *   attention kernels read it as an MN-major tensor-core operand, no transposed copy exists)
 * Padding regions are never written: the caller zero-fills q / k once and prepares vt with
 * radvlm_attention_prepare_vt (zero padding + the ones column the attention kernel sums P with). */
int radvlm_gemm_qkv_split(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K,
                          const float* bias, void* q, void* k, void* vt, int seq, int seq_pad,
                          int heads, int hd, int hd_pad, int block_n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused non-causal attention for one ViT block (siglip_encoder.py:216-235: q k^T * scale ->
 * softmax(fp32) -> p v -> transpose/reshape).  Inputs in the layout radvlm_gemm_qkv_split writes;
 * out: bf16 [tiles*seq, heads*hd] token-major (A operand of out_proj).  vt must have been prepared with
 * radvlm_attention_prepare_vt before the QKV epilogue filled it.
 * Supported geometry: hd_pad == 80, hd < hd_pad, hd % 8 == 0, seq_pad % 384 == 0 (query groups of 128 rows, two per
 * work item; key blocks of 96), 1 <= seq <= seq_pad; out 16-byte aligned (it is written with TMA tile stores).
 * ---------------------------------------------------------------------------------------------- */
/* Zero the padded V buffer (bf16 [tiles, heads, seq_pad, hd_pad]) and write a one into column `hd` of every valid
 * key: the PV tensor-core product then also accumulates the softmax row sum (O[:, hd]). */
int radvlm_attention_prepare_vt(void* vt, int tiles, int heads, int seq, int seq_pad, int hd, int hd_pad,
                                void* stream);
int radvlm_attention_fwd(const void* q, const void* k, const void* vt, void* out, int tiles, int heads,
                         int seq, int seq_pad, int hd, int hd_pad, float scale, void* stream);
/* Same, also writing lse fp32 [tiles*heads, seq_pad]: the base-2 log-sum-exp of the scaled scores of every valid
 * query row (softmax = 2^(s * scale * log2(e) - lse)), which the backward kernel recomputes P from. */
int radvlm_attention_fwd_lse(const void* q, const void* k, const void* vt, void* out, float* lse, int tiles, int heads,
                             int seq, int seq_pad, int hd, int hd_pad, float scale, void* stream);

/* Backward of the attention (autograd of siglip_encoder.py:216-235).  q, k: the padded head-major buffers of the
 * forward; vt: V ([tiles, heads, seq_pad, hd_pad]) with plain ZEROS in its padding (no ones column); dout / out: gradient and value of the
 * attention output, bf16 token-major [tiles*seq, heads*hd]; lse from radvlm_attention_fwd_lse.
 * dqkv: bf16 [tiles*seq, 3*heads*hd] = [dQ | dK | dV] in the column order of the concatenated QKV projection, i.e.
 * the dY of that Linear.  workspace: radvlm_attention_bwd_workspace_bytes (row dots + fp32 dQ accumulator). */
size_t radvlm_attention_bwd_workspace_bytes(int tiles, int heads, int seq_pad);
int radvlm_attention_bwd(const void* q, const void* k, const void* vt, const void* dout, const void* out,
                         const float* lse, void* dqkv, void* workspace, size_t workspace_bytes, int tiles, int heads,
                         int seq, int seq_pad, int hd, int hd_pad, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound tower helpers.
 *   radvlm_layernorm_f32_bf16 : nn.LayerNorm(eps) over the fp32 residual stream, bf16 result
 *                               (siglip_encoder.py:264,266,287,296)
 *   radvlm_patch_im2col       : pixel tiles [n,C,S,S] (RADVLM_DT_*) -> bf16 [n*P*P, k_pad] patch rows,
 *                               column order == patch_embedding.weight.flatten(1) (siglip_encoder.py:156-171)
 *   radvlm_cast_f32_bf16      : tower output -> projector operand
 * ---------------------------------------------------------------------------------------------- */
int radvlm_layernorm_f32_bf16(const float* x, const float* gamma, const float* beta, void* y, int rows,
                              int D, float eps, void* stream);
int radvlm_patch_im2col(const void* pixels, int dtype, void* out, int n_tiles, int channels,
                        int image_size, int patch_size, int k_pad, void* stream);
int radvlm_cast_f32_bf16(const float* x, void* y, size_t n, void* stream);
/* (mean, rstd) of every row of a bf16 [rows, D] matrix -> stats: fp32 [rows][2].  The LayerNorm statistics when the
 * normalisation itself is folded into the consuming GEMM (radvlm_vit_layer_weights.qkv_wf). */
int radvlm_ln_row_stats_bf16(const void* x, void* stats, int rows, int D, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-path entry points: SigLIP tower (SigLipVisionTower.forward, siglip_encoder.py:576-589,
 * i.e. embeddings + the executed encoder layers, hidden_states[-1], no post_layernorm) and the
 * mlp2x_gelu projector (llava_arch.py:192-196 encode_images; builder.py:41-48).
 *
 * Weight blocks hold DEVICE pointers to bf16 matrices in nn.Linear layout [out, in] and fp32 vectors;
 * the structs themselves live in HOST memory.  The host side (radvlm_b200/encoder.py) packs them
 * from the reference's Parameters (state-dict names in SURVEY.md section 5) and caches by version.
 * ---------------------------------------------------------------------------------------------- */
typedef struct radvlm_vit_layer_weights {
  const float* ln1_gamma;
  const float* ln1_beta;
  const void* qkv_w;   /* bf16 [3*hidden, hidden] = rows of q_proj; k_proj; v_proj */
  const float* qkv_b;  /* [3*hidden] */
  const void* out_w;   /* bf16 [hidden, hidden] */
  const float* out_b;
  const float* ln2_gamma;
  const float* ln2_beta;
  const void* fc1_w;   /* bf16 [intermediate, hidden] */
  const float* fc1_b;
  const void* fc2_w;   /* bf16 [hidden, intermediate] */
  const float* fc2_b;
  /* LayerNorm folded into the QKV / fc1 GEMMs (inference; all six NULL = stand-alone LayerNorm kernels).
   *   LN(x) W^T + b  =  rstd * (x (gamma o W)^T - mean * s) + b',   s[n] = sum_k (gamma o W)[n, k],  b' = b + W beta
   * qkv_wf / fc1_wf: bf16 [out, hidden] = bf16(gamma o W);  *_sf: fp32 [out] row sums of the bf16 matrix;
   * *_bf: fp32 [out].  The residual GEMM epilogues then also emit a bf16 copy of the stream (the A operand) and a
   * 2 B/element pass (radvlm_ln_row_stats_bf16) supplies (mean, rstd); siglip_encoder.py:264,266,287,296. */
  const void* qkv_wf;
  const float* qkv_sf;
  const float* qkv_bf;
  const void* fc1_wf;
  const float* fc1_sf;
  const float* fc1_bf;
} radvlm_vit_layer_weights;

typedef struct radvlm_siglip_weights {
  int hidden;       /* 1152 */
  int intermediate; /* 4304 */
  int heads;        /* 16 */
  int num_layers;   /* executed layers: 26 (siglip_encoder.py:570 drops the 27th) */
  int image_size;   /* 384 */
  int patch_size;   /* 14 */
  int channels;     /* 3 */
  int patch_k_pad;  /* 640: 3*14*14 = 588 zero padded to a multiple of 64 */
  float ln_eps;     /* 1e-6 */
  const void* patch_w;    /* bf16 [hidden, patch_k_pad] = patch_embedding.weight.flatten(1), zero padded */
  const float* patch_b;   /* [hidden] */
  const float* pos_embed; /* fp32 [P*P, hidden] */
  const radvlm_vit_layer_weights* layers; /* host array [num_layers] */
} radvlm_siglip_weights;

typedef struct radvlm_projector_weights {
  int in_dim; /* 1152 */
  int hidden; /* 3584 */
  const void* w1;  /* bf16 [hidden, in_dim] */
  const float* b1;
  const void* w2;  /* bf16 [hidden, hidden] */
  const float* b2;
} radvlm_projector_weights;

/* scratch needed by radvlm_encode_images for n_tiles tiles (0 on bad arguments) */
size_t radvlm_encode_workspace_bytes(const radvlm_siglip_weights* tw, const radvlm_projector_weights* pw,
                                     int n_tiles);

/* pixels: [n_tiles, C, S, S] of pixel_dtype -> hidden_out: fp32 [n_tiles*P*P, hidden] */
int radvlm_siglip_tower_forward(const radvlm_siglip_weights* tw, const void* pixels, int pixel_dtype,
                                int n_tiles, float* hidden_out, void* workspace, size_t workspace_bytes,
                                void* stream);

/* hidden: fp32 [rows, in_dim] -> features_out: [rows, hidden] of out_dtype (RADVLM_DT_BF16 | RADVLM_DT_F16 | RADVLM_DT_F32).
 * workspace: rows*in_dim*2 + rows*hidden*2 bytes (each rounded up to 1 KB).  With (rows/256 + 2)*4 more bytes (rounded
 * up to 1 KB) both GEMMs run as ONE persistent kernel (the second waits per row block for the first; the intermediate
 * stays in L2); without them, or for shapes that kernel does not cover, as two launches - same result bit for bit. */
int radvlm_projector_forward(const radvlm_projector_weights* pw, const float* hidden, int rows,
                             void* features_out, int out_dtype, void* workspace, size_t workspace_bytes,
                             void* stream);

/* encode_images: tower + projector.  features_out: [n_tiles*P*P, projector hidden] of out_dtype */
int radvlm_encode_images(const radvlm_siglip_weights* tw, const radvlm_projector_weights* pw,
                         const void* pixels, int pixel_dtype, int n_tiles, void* features_out,
                         int out_dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training mode (BASELINE config 5; mm_tunable_parts = mm_vision_tower, mm_mlp_adapter: train.py:1642-1665).
 * The forward keeps, per layer, the fp32 residual stream entering it and after its attention branch, q / k / v, the
 * attention output, the attention log-sum-exp, the fc1 pre-activation and both LayerNorm outputs (`saved`,
 * radvlm_tower_saved_bytes: 624 MB per tile at the SigLIP-so400m widths); the backward recomputes nothing (the reference checkpoints whole encoder layers:
 * siglip_encoder.py:381-387).  Gradient pointers are fp32, same shapes as the
 * packed weights, ACCUMULATED (+=); a null pointer freezes that parameter.
 * ---------------------------------------------------------------------------------------------- */
typedef struct radvlm_vit_layer_grads {
  float* ln1_gamma; float* ln1_beta;
  float* qkv_w;  /* [3*hidden, hidden] */
  float* qkv_b;
  float* out_w;  float* out_b;
  float* ln2_gamma; float* ln2_beta;
  float* fc1_w;  float* fc1_b;
  float* fc2_w;  float* fc2_b;
} radvlm_vit_layer_grads;

typedef struct radvlm_siglip_grads {
  float* patch_w;   /* [hidden, patch_k_pad] */
  float* patch_b;
  float* pos_embed; /* [P*P, hidden] */
  const radvlm_vit_layer_grads* layers; /* host array [num_layers] or NULL */
} radvlm_siglip_grads;

typedef struct radvlm_projector_grads {
  float* w1; float* b1; float* w2; float* b2;
} radvlm_projector_grads;

size_t radvlm_tower_saved_bytes(const radvlm_siglip_weights* tw, int n_tiles);
/* byte offset, inside `saved`, of the tower output: fp32 [n_tiles*P*P, hidden] */
size_t radvlm_tower_saved_hidden_offset(const radvlm_siglip_weights* tw, int n_tiles);
/* workspace: radvlm_encode_workspace_bytes(tw, NULL, n_tiles) */
int radvlm_siglip_tower_forward_train(const radvlm_siglip_weights* tw, const void* pixels, int pixel_dtype,
                                      int n_tiles, void* saved, size_t saved_bytes, void* workspace,
                                      size_t workspace_bytes, void* stream);
size_t radvlm_tower_backward_workspace_bytes(const radvlm_siglip_weights* tw, int n_tiles);
/* d_hidden: fp32 [n_tiles*P*P, hidden], dL/d(tower output) on entry; used as the running residual gradient. */
int radvlm_siglip_tower_backward(const radvlm_siglip_weights* tw, const radvlm_siglip_grads* grads, const void* pixels,
                                 int pixel_dtype, int n_tiles, const void* saved, size_t saved_bytes, float* d_hidden,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* Same, for encoder layers [layer_lo, layer_hi) only (top down; the embeddings follow when layer_lo == 0).  Call it
 * range by range from the top with the same d_hidden / workspace to overlap the gradient all-reduce of finished
 * layers (radvlm_b200.dist.allreduce_gradients) with the backward of the next range. */
int radvlm_siglip_tower_backward_range(const radvlm_siglip_weights* tw, const radvlm_siglip_grads* grads,
                                       const void* pixels, int pixel_dtype, int n_tiles, const void* saved,
                                       size_t saved_bytes, float* d_hidden, void* workspace, size_t workspace_bytes,
                                       int layer_lo, int layer_hi, void* stream);
size_t radvlm_projector_backward_workspace_bytes(const radvlm_projector_weights* pw, int rows);
/* hidden: fp32 [rows, in_dim] (the projector input); d_features: bf16 [rows, hidden];
 * d_hidden: fp32 [rows, in_dim] written (may be NULL when the tower is frozen). */
int radvlm_projector_backward(const radvlm_projector_weights* pw, const radvlm_projector_grads* grads,
                              const float* hidden, const void* d_features, int rows, float* d_hidden, void* workspace,
                              size_t workspace_bytes, void* stream);
/* helpers exposed for tests */
int radvlm_colsum_bf16(const void* x, int rows, int cols, int ld, float* out, void* stream);
int radvlm_gelu_fwd_bwd_bf16(const void* u, void* da_du, void* a, int64_t n, int erf_form, void* stream);
/* dres += dL/dx; dgamma / dbeta += (both NULL: frozen affine); row_stats_scratch: rows * 8 bytes */
int radvlm_layernorm_bwd(const float* x, const float* gamma, const void* dy, float* dres, float* dgamma, float* dbeta,
                         void* row_stats_scratch, int rows, int D, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Planner (CPU only, no CUDA): every integer decision of the path with the reference's Python
 * float64/int semantics, bit-exact.
 *   radvlm_plan_select_best_resolution : mm_utils.py:119-149 (candidates are (width,height) pairs)
 *   radvlm_plan_image                  : mm_utils.py:152-188,213-240 ; llava_arch.py:127-159,386-390
 *   radvlm_plan_splice                 : llava_arch.py:428-531
 * ---------------------------------------------------------------------------------------------- */
typedef struct radvlm_image_plan {
  int32_t width, height;          /* original size, PIL order (W,H)  (llava_arch.py:138) */
  int32_t best_w, best_h;         /* select_best_resolution */
  int32_t grid_w, grid_h;         /* (num_patch_width, num_patch_height) */
  int32_t resized_w, resized_h;   /* aspect-preserving resize inside the canvas */
  int32_t paste_x, paste_y;       /* centred paste offset */
  int32_t n_tiles;                /* 1 + grid_w*grid_h (base tile first) */
  int32_t crop_r0, crop_c0, crop_h, crop_w; /* unpad window in the (S*grid_h, S*grid_w) feature map */
  int32_t pool;                   /* 1 when times > 1.1 (bilinear resize to out_h x out_w) */
  int32_t out_h, out_w;           /* feature rows / cols before the newline column */
  int32_t n_tokens;               /* S*S + out_h*(out_w+1) */
} radvlm_image_plan;

int radvlm_plan_select_best_resolution(int W, int H, const int32_t* pinpoints, int n, int* best_w,
                                       int* best_h);
int radvlm_plan_image(int W, int H, const int32_t* pinpoints, int n_pinpoints, int tile_size,
                      int patches_per_side, int max_num_patches /* <=0: never pool */,
                      radvlm_image_plan* out);

#define RADVLM_SEG_PAD 0
#define RADVLM_SEG_TEXT 1
#define RADVLM_SEG_IMAGE 2
typedef struct radvlm_splice_segment {
  int64_t dst_row;  /* first output row, b*max_len + position */
  int32_t length;   /* rows */
  int32_t kind;     /* RADVLM_SEG_* */
  int32_t src_off;  /* TEXT: offset into text_src[];  IMAGE: first visual token (0) */
  int32_t image;    /* IMAGE: index into the merge-image table */
  int32_t pos0;     /* position id of the first row */
  int32_t reserved;
} radvlm_splice_segment;

/* input_ids: host int64 [B,L]; attention_mask: host uint8 [B,L] or NULL (all ones);
 * image_tokens[n_images]: visual tokens each image contributes (in consumption order).
 * Outputs: segments (sorted by dst_row, covering all B*max_len rows), text_src (flat b*L+i source
 * positions of kept text tokens), lengths[B], max_len.  A text-only sample consumes one image
 * (llava_arch.py:452-459); when the list runs out inside a sample the previous image is reused
 * (llava_arch.py:478-481); otherwise RADVLM_ERR_BAD_ARGUMENT ("IndexError"). */
int radvlm_plan_splice(const int64_t* input_ids, const uint8_t* attention_mask, int B, int L,
                       int image_token_index, const int32_t* image_tokens, int n_images,
                       int64_t max_length, int left_pad, radvlm_splice_segment* segments,
                       int segment_capacity, int* n_segments, int32_t* text_src, int text_capacity,
                       int* n_text, int32_t* lengths, int* max_len_out);

/* ------------------------------------------------------------------------------------------------
 * Merge + splice gather kernel: spatial unpad / anyres_max bilinear pool / image_newline / base-tile
 * prepend (llava_arch.py:350-412), embed_tokens gather + interleave (llava_arch.py:449-493) and
 * truncate / pad / stack with labels, attention_mask, position_ids (llava_arch.py:495-531) in ONE pass:
 * every output row [b, p, :] is produced once, directly in the padded [B, max_len, H] tensor.
 * ---------------------------------------------------------------------------------------------- */
#define RADVLM_MERGE_ANYRES 0  /* base tile + unpadded (optionally pooled) grid with newline column; `reserved` holds
                                * RADVLM_ANYRES_* flags for the other spatial merge types (llava_arch.py:376-404) */
#define RADVLM_ANYRES_NO_NEWLINE 1 /* no image_newline column: "spatial" without 'unpad' (:398-400), 'maxpool2x2' (:376-380) */
#define RADVLM_ANYRES_NO_BASE 2    /* 'nobase' in mm_patch_merge_type: the base tile is not prepended (:401-404) */
/* pool == RADVLM_POOL_MAX in an ANYRES entry: nn.functional.max_pool2d(., 2) of the whole (un-cropped) grid
 * ('maxpool2x2'): out_h x out_w = floor(S*grid_h / 2) x floor(S*grid_w / 2). */
#define RADVLM_MERGE_SINGLE 1  /* one tile + one newline token (llava_arch.py:407-412) */
#define RADVLM_MERGE_FLAT 2    /* tiles*T tokens, no newline ("flat", or a 4-D image batch) */
/* Video sample (llava_arch.py:171-190 get_2dPool, :222-249 add_token_per_grid / _frame, :310-349): the entry's
 * `grid_w` tiles are FRAMES; each S x S frame is pooled with stride 2 to out_h x out_w (`pool` = RADVLM_POOL_*:
 * bilinear -> ceil(S/2) with ATen's align_corners=False taps, average / max -> floor(S/2) 2x2 windows) and the frames
 * are laid out frame-major with image_newline rows placed as `reserved` = RADVLM_NEWLINE_* says.  The backward
 * covers bilinear and average pooling (max pooling needs the forward argmax; the host mirror raises for it). */
#define RADVLM_MERGE_VIDEO 3
#define RADVLM_POOL_NONE 0
#define RADVLM_POOL_BILINEAR 1 /* also the anyres_max pooling flag of RADVLM_MERGE_ANYRES */
#define RADVLM_POOL_AVERAGE 2
#define RADVLM_POOL_MAX 3
#define RADVLM_NEWLINE_NONE 0  /* mm_newline_position "no_token" (and "one_token" without 'unpad', and "flat") */
#define RADVLM_NEWLINE_ONE 1   /* "one_token": a single newline after the last frame */
#define RADVLM_NEWLINE_FRAME 2 /* "frame": one newline after every frame */
#define RADVLM_NEWLINE_GRID 3  /* "grid": one newline after every pooled row of every frame */
typedef struct radvlm_merge_image {
  int32_t tile_base; /* index of the image's first (base) tile in the feature buffer */
  int32_t mode;      /* RADVLM_MERGE_* */
  int32_t grid_w;
  int32_t crop_r0, crop_c0, crop_h, crop_w;
  int32_t pool, out_h, out_w;
  int32_t n_tokens;
  int32_t reserved;  /* RADVLM_MERGE_VIDEO: RADVLM_NEWLINE_*;  RADVLM_MERGE_ANYRES: RADVLM_ANYRES_* flags */
} radvlm_merge_image;

/* All pointers are DEVICE pointers.  dtype: element type of features / newline / embed table / out.
 *   features [tiles, T, H]; newline [H]; embed_table [vocab, H]; input_ids int64 [B*L];
 *   labels_in int64 [B*L] or NULL (=> all IGNORE_INDEX); segments / images / text_src: tables from the planner.
 *   out_embeds [B*max_len, H]; out_labels int64; out_mask uint8; out_pos int64 (each [B*max_len]; may be NULL) */
int radvlm_merge_splice(const void* features, const void* newline, const void* embed_table, int dtype,
                        int hidden, int tokens_per_tile, int patches_per_side, const int64_t* input_ids,
                        const int64_t* labels_in, const int32_t* text_src,
                        const radvlm_splice_segment* segments, int n_segments,
                        const radvlm_merge_image* images, int n_images, int64_t total_rows,
                        void* out_embeds, int64_t* out_labels, uint8_t* out_mask, int64_t* out_pos,
                        int64_t ignore_index, void* stream);

/* Fused merge + all-gather (one process per GPU, images sharded by rank): the same gather, but every
 * embedding row is written to `n_peers` destinations, destination d = this rank's [total_rows, H] slot inside rank
 * d's gathered buffer, through peer-mapped pointers (plain stores over NVLink / NVSwitch).  The merged tokens are read
 * once and no separate collective moves them; labels / mask / position ids stay local.  max_ctas > 0 caps the grid so
 * the kernel can run on a side stream beside the next step's tower.  The reference has no counterpart (its data
 * parallelism lives in DeepSpeed); the NCCL all-gather of radvlm_b200/dist.py is the equivalent two-step form. */
#define RADVLM_MAX_PEERS 8
int radvlm_merge_splice_scatter(const void* features, const void* newline, const void* embed_table, int dtype,
                                int hidden, int tokens_per_tile, int patches_per_side, const int64_t* input_ids,
                                const int64_t* labels_in, const int32_t* text_src,
                                const radvlm_splice_segment* segments, int n_segments,
                                const radvlm_merge_image* images, int n_images, int64_t total_rows,
                                void* const* out_peers, int n_peers, int max_ctas, int64_t* out_labels,
                                uint8_t* out_mask, int64_t* out_pos, int64_t ignore_index, void* stream);
/* Peer memory for the call above: radvlm_peer_alloc = cudaMalloc (zero-filled) + a 64-byte cudaIpc handle the host
 * exchanges (e.g. torch.distributed.all_gather); radvlm_peer_open maps a peer's allocation into this process.
 * radvlm_peer_signal_wait (on `stream`): publish `value` into slot [rank] of every rank's flag array (uint64[n],
 * peer-mapped device pointers in the DEVICE array flags_peers_dev), then wait until all n slots of flags_local have
 * reached it - the step barrier that orders the scattered rows of all ranks before any rank reads its buffer (and,
 * on a second flag array, the consumer-release barrier that orders every rank's reads of a slot before the slot is
 * overwritten).  The wait is bounded by wall time: timeout_s <= 0 waits for ever; on a timeout the kernel stores
 * 1 + (first missing rank) into *status (pinned host or device int, 0 on entry, may be NULL) and returns - it never
 * traps, the CUDA context survives and the host decides.
 * radvlm_peer_copy: cudaMemcpyAsync(cudaMemcpyDefault) of a finished slice into a peer's buffer: the copy-engine form
 * of the exchange (DMA over NVLink, no SM). */
int radvlm_peer_alloc(size_t bytes, void** ptr, uint8_t* handle64);
int radvlm_peer_open(const uint8_t* handle64, void** ptr);
int radvlm_peer_close(void* ptr);
int radvlm_peer_free(void* ptr);
int radvlm_peer_signal_wait(void* const* flags_peers_dev, void* flags_local, int n, int rank,
                            unsigned long long value, double timeout_s, int* status, void* stream);
int radvlm_peer_copy(void* dst, const void* src, size_t bytes, void* stream);

/* Backward of radvlm_merge_splice (autograd of llava_arch.py:350-531).  d_out_embeds: [B*max_len, H] of `dtype`.
 *   d_features fp32 [tiles*T, H] and d_newline fp32 [H]: accumulated with atomics (zero or running sums on entry);
 *   d_text [n_text, H] of `dtype` (or NULL): gradient rows of the text tokens in text_src order;
 *   features: the forward's feature tensor ([tiles*T, H] of `dtype`), needed only when an entry pools with
 *   RADVLM_POOL_MAX (video 'max' pooling, 'maxpool2x2'): the gradient of a window goes to its arg-max, first maximum
 *   in row-major window order like ATen's max_pool2d backward.  NULL: max-pooled rows get no gradient. */
int radvlm_merge_splice_backward(const void* d_out_embeds, int dtype, int hidden, int tokens_per_tile,
                                 int patches_per_side, const radvlm_splice_segment* segments, int n_segments,
                                 const radvlm_merge_image* images, int n_images, int64_t total_rows,
                                 float* d_features, float* d_newline, void* d_text, const void* features, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused anyres preprocessing (uint8 -> resize -> pad -> tile -> normalise), bit-exact with
 * process_anyres_image + SigLipImageProcessor.preprocess (mm_utils.py:152-210,243-293;
 * siglip_encoder.py:47-67; Pillow 8bpc BICUBIC ImagingResample restated on the device).
 *
 * src: device buffer holding the uint8 images (HWC, 3 channels, or HW single channel that is
 * replicated to RGB).  Per image a descriptor (same content on host and device): geometry from
 * radvlm_plan_image, the first output tile index, and byte offsets into src / scratch.
 * tiles_out: [total_tiles, 3, tile_size, tile_size] of out_dtype; tile order per image = base tile
 * first, then the crops row-major (mm_utils.py:204-208,291).
 * ---------------------------------------------------------------------------------------------- */
typedef struct radvlm_preprocess_image {
  int64_t src_offset;     /* byte offset of this image in src */
  int64_t scratch_offset; /* byte offset of this image's scratch region (16-byte aligned) */
  int32_t width, height, channels; /* channels: 3 (HWC) or 1 (grayscale, replicated) */
  int32_t grid_w, grid_h;
  int32_t resized_w, resized_h;
  int32_t paste_x, paste_y;
  int32_t tile_base;
} radvlm_preprocess_image;

size_t radvlm_preprocess_scratch_bytes(int width, int height, int channels, int resized_w, int resized_h,
                                       int tile_size);
int radvlm_preprocess_anyres(const uint8_t* src, const radvlm_preprocess_image* images_dev,
                             const radvlm_preprocess_image* images_host, int n_images, int tile_size,
                             void* tiles_out, int out_dtype, void* scratch, size_t scratch_bytes,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADVLM_B200_H_ */
