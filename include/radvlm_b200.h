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

int radvlm_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     const float* bias, int epilogue, void* out, int64_t ldo, const float* aux,
                     int aux_period, int block_n /* 0 = auto, else 128|192|256 */, void* stream);

/* QKV projection with the head-split scatter fused into the epilogue
 * (siglip_encoder.py:207-213: three Linear + view/transpose).  W is the row-concatenation
 * [q_proj; k_proj; v_proj] = [3*heads*hd, K]; bias likewise.  Outputs (bf16):
 *   q, k : [tiles, heads, seq_pad, hd_pad]       vt : [tiles, heads, hd_pad, seq_pad]  (V transposed)
 * Padding regions are never written: the caller zero-fills the buffers once. */
int radvlm_gemm_qkv_split(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K,
                          const float* bias, void* q, void* k, void* vt, int seq, int seq_pad,
                          int heads, int hd, int hd_pad, int block_n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused non-causal attention for one ViT block (siglip_encoder.py:216-235: q k^T * scale ->
 * softmax(fp32) -> p v -> transpose/reshape).  Inputs in the layout radvlm_gemm_qkv_split writes;
 * out: bf16 [tiles*seq, heads*hd] token-major (A operand of out_proj).
 * Supported geometry: hd_pad == 80, hd % 8 == 0, seq_pad % 128 == 0, seq_pad - 128 < seq <= seq_pad.
 * ---------------------------------------------------------------------------------------------- */
int radvlm_attention_fwd(const void* q, const void* k, const void* vt, void* out, int tiles, int heads,
                         int seq, int seq_pad, int hd, int hd_pad, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADVLM_B200_H_ */
