// Internal launch functions shared between translation units of the library.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "gemm_args.h"

namespace rv {

int gemm_dispatch(const void* A, int64_t lda, const void* W, int64_t ldw, const GemmArgs& args,
                  int epilogue, int block_n, cudaStream_t stream);
// mlp2x_gelu projector as one persistent two-GEMM kernel (gemm4_sm100.cuh); RADVLM_ERR_UNSUPPORTED_SHAPE = not covered
int projector_chain_dispatch(const void* X, const void* W1, const float* b1, void* H, const void* W2, const float* b2,
                             void* out, int out_dtype, int rows, int in_dim, int hidden, void* ready, cudaStream_t stream);
// two dependent GEMMs as one persistent kernel (gemm4_sm100.cuh): g0 writes the bf16 intermediate, g1 consumes it
int gemm_chain_dispatch(const void* X, const void* W1, const void* W2, const GemmArgs& g0, int epi0, const GemmArgs& g1,
                        int epi1, void* ready, cudaStream_t stream);
int gemm_pick_block_n(int M, int N, int cta_group);
int gemm_bmn_block_n(int N);  // tile width of the GEMMs whose B operand is MN-major (data / weight gradients)
int attention_launch(const void* q, const void* k, const void* vt, void* out, float* lse, int tiles, int heads,
                     int seq, int seq_pad, int hd, int hd_pad, float scale, cudaStream_t stream);
size_t attention_bwd_workspace_bytes(int tiles, int heads, int seq_pad);
int attention_bwd_launch(const void* q, const void* k, const void* vt, const void* dout, const void* out,
                         const float* lse, void* dqkv, void* workspace, size_t workspace_bytes, int tiles, int heads,
                         int seq, int seq_pad, int hd, int hd_pad, float scale, cudaStream_t stream);
int qkv_pad_prepare_launch(void* q, void* k, void* vt, int tiles, int heads, int seq, int seq_pad, int hd, int hd_pad,
                           float v_one, cudaStream_t stream);
int attention_prepare_vt_launch(void* vt, int tiles, int heads, int seq, int seq_pad, int hd, int hd_pad,
                                cudaStream_t stream);
int colsum_bf16_launch(const void* x, int rows, int cols, int ld, float* out, cudaStream_t stream);
int gelu_fwd_bwd_launch(const void* u, void* da_du, void* a, size_t n, int erf_form, cudaStream_t stream);
int layernorm_bwd_launch(const float* x, const float* gamma, const void* dy, float* dres, float* dgamma, float* dbeta,
                         void* stats, int rows, int D, float eps, cudaStream_t stream, void* dres_bf16 = nullptr);
int pos_embed_grad_launch(const float* dh, float* dpos, int tiles, int T, int D, cudaStream_t stream);
int layernorm_launch(const float* x, const float* gamma, const float* beta, void* y, int rows, int D,
                     float eps, cudaStream_t stream);
int ln_finalize_stats_launch(const void* part, void* stats, int rows, int slots, int D, float eps, cudaStream_t stream);
// slots of GemmArgs::ln_part the scheduled kernel fills for an [M, N] output; 0 when that kernel does not cover the shape
int gemm_ln_part_slots(int M, int N);
int ln_row_stats_launch(const void* x, void* stats, int rows, int D, float eps, cudaStream_t stream);
int cast_f32_bf16_launch(const float* x, void* y, size_t n, cudaStream_t stream);
int im2col_launch(const void* pixels, int dtype, void* out, int n_tiles, int C, int S, int ps,
                  int Kpad, cudaStream_t stream);

}  // namespace rv
