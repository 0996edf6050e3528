// C-ABI launcher for the attention backward (attention_bwd_sm100.cuh).
#include <algorithm>

#include "attention_bwd_sm100.cuh"
#include "host_util.h"
#include "internal.h"

namespace rv {

#ifdef RV_ABWD_TIMELINE
static long long* g_abwd_timeline = nullptr;
#endif

static size_t attn_bwd_delta_bytes(int tiles, int heads, int seq_pad) {
  return align_up(static_cast<size_t>(tiles) * heads * seq_pad * sizeof(float), 1024);
}

size_t attention_bwd_workspace_bytes(int tiles, int heads, int seq_pad) {
  return attn_bwd_delta_bytes(tiles, heads, seq_pad) + static_cast<size_t>(tiles) * heads * seq_pad * kAbDqPitch * sizeof(float);
}

int attention_bwd_launch(const void* q, const void* k, const void* vt, const void* dout, const void* out,
                         const float* lse, void* dqkv, void* workspace, size_t workspace_bytes, int tiles, int heads,
                         int seq, int seq_pad, int hd, int hd_pad, float scale, cudaStream_t stream) {
  int st = require_sm100();
  if (st != RADVLM_OK) return st;
  RV_CHECK_ARG(q && k && vt && dout && out && lse && dqkv && workspace, "attention_bwd: null pointer");
  RV_CHECK_ARG(tiles > 0 && heads > 0, "attention_bwd: bad batch (tiles=%d heads=%d)", tiles, heads);
  if (hd_pad != 80 || hd >= hd_pad || (hd % 8) != 0 || (seq_pad % 128) != 0 || seq > seq_pad || seq < 1 ||
      ((heads * hd) % 8) != 0) {
    set_error("attention_bwd: unsupported geometry seq=%d seq_pad=%d hd=%d hd_pad=%d", seq, seq_pad, hd, hd_pad);
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  if (workspace_bytes < attention_bwd_workspace_bytes(tiles, heads, seq_pad)) {
    set_error("attention_bwd: workspace too small (%zu < %zu)", workspace_bytes,
              attention_bwd_workspace_bytes(tiles, heads, seq_pad));
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  static thread_local bool configured = false;
  if (!configured) {
    RV_CUDA(cudaFuncSetAttribute(siglip_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes));
    configured = true;
  }
  float* delta = static_cast<float*>(workspace);
  float* dq_acc = reinterpret_cast<float*>(static_cast<char*>(workspace) + attn_bwd_delta_bytes(tiles, heads, seq_pad));
  const uint64_t th = static_cast<uint64_t>(tiles) * heads;
  const int tokens = tiles * seq;
  const int D = heads * hd;
  RV_CUDA(cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(th) * seq_pad * kAbDqPitch * sizeof(float), stream));
  {
    RV_CHECK_ARG(heads * (hd / 8) <= 256, "attention_bwd: heads * hd / 8 must be <= 256");
    const int blocks = (tokens + 7) / 8;
    attn_delta_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dout),
                                                  static_cast<const __nv_bfloat16*>(out), delta, tokens, seq, seq_pad,
                                                  heads, hd);
    RV_CUDA(cudaGetLastError());
  }
  CUtensorMap tq, tk, tv, tdo;
  const uint64_t pitch = static_cast<uint64_t>(hd_pad) * 2;
  st = make_tmap_bf16_2d(&tq, q, hd_pad, th * seq_pad, pitch, 16, 128, CU_TENSOR_MAP_SWIZZLE_32B);
  if (st != RADVLM_OK) return st;
  st = make_tmap_bf16_2d(&tk, k, hd_pad, th * seq_pad, pitch, 16, 128, CU_TENSOR_MAP_SWIZZLE_32B);
  if (st != RADVLM_OK) return st;
  st = make_tmap_bf16_2d(&tv, vt, hd_pad, th * seq_pad, pitch, 16, 128, CU_TENSOR_MAP_SWIZZLE_32B);
  if (st != RADVLM_OK) return st;
  st = make_tmap_bf16_2d(&tdo, dout, D, tokens, static_cast<uint64_t>(D) * 2, 16, 128, CU_TENSOR_MAP_SWIZZLE_32B);
  if (st != RADVLM_OK) return st;
  AttnBwdArgs a;
  a.lse = lse;
  a.delta = delta;
  a.dq_acc = dq_acc;
  a.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  a.seq = seq; a.seq_pad = seq_pad; a.heads = heads; a.hd = hd;
  a.scale = scale;
  a.scale_log2e = scale * 1.4426950408889634f;
#ifdef RV_ABWD_TIMELINE
  a.timeline = g_abwd_timeline;
#endif
  dim3 grid((seq + 127) / 128, heads, tiles);
  siglip_attention_bwd_kernel<<<grid, kAbThreads, kAbSmemBytes, stream>>>(tq, tk, tv, tdo, a);
  RV_CUDA(cudaGetLastError());
  {
    const size_t total = static_cast<size_t>(tokens) * heads * (hd / 8);
    attn_dq_store_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(dq_acc, a.dqkv, tokens, seq, seq_pad,
                                                                                   heads, hd);
    RV_CUDA(cudaGetLastError());
  }
  return RADVLM_OK;
}

}  // namespace rv

extern "C" size_t radvlm_attention_bwd_workspace_bytes(int tiles, int heads, int seq_pad) {
  return rv::attention_bwd_workspace_bytes(tiles, heads, seq_pad);
}

extern "C" int radvlm_attention_bwd(const void* q, const void* k, const void* vt, const void* dout, const void* out,
                                    const float* lse, void* dqkv, void* workspace, size_t workspace_bytes, int tiles,
                                    int heads, int seq, int seq_pad, int hd, int hd_pad, float scale, void* stream) {
  return rv::attention_bwd_launch(q, k, vt, dout, out, lse, dqkv, workspace, workspace_bytes, tiles, heads, seq,
                                  seq_pad, hd, hd_pad, scale, static_cast<cudaStream_t>(stream));
}

#ifdef RV_ABWD_TIMELINE
extern "C" void radvlm_debug_abwd_timeline(void* buffer) { rv::g_abwd_timeline = static_cast<long long*>(buffer); }
#endif
