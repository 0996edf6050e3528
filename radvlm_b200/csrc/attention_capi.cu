// C-ABI launcher for the fused SigLIP attention kernels (attention_pp_sm100.cuh: the product path;
// attention_sm100.cuh: the earlier two-CTAs-per-SM organisation, kept as the A/B baseline of tools/).
#include <algorithm>

#include <cstdlib>

#include "attention_pp_sm100.cuh"
#include "host_util.h"

namespace rv {

int attention_launch(const void* q, const void* k, const void* vt, void* out, float* lse, int tiles, int heads,
                     int seq, int seq_pad, int hd, int hd_pad, float scale, cudaStream_t stream) {
  int st = require_sm100();
  if (st != RADVLM_OK) return st;
  RV_CHECK_ARG(q && k && vt && out, "attention: null pointer");
  RV_CHECK_ARG(tiles > 0 && heads > 0, "attention: bad batch (tiles=%d heads=%d)", tiles, heads);
  if (hd_pad != kAttnHdPad || hd >= hd_pad || (hd % 8) != 0 || (seq_pad % kAttnBKV) != 0 ||
      (seq_pad % kAttnBQ) != 0 || seq > seq_pad || seq < 1) {
    set_error("attention: unsupported geometry seq=%d seq_pad=%d hd=%d hd_pad=%d (need hd_pad=80, hd<80, "
              "hd%%8==0, seq_pad%%384==0, 1 <= seq <= seq_pad)", seq, seq_pad, hd, hd_pad);
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  static thread_local bool configured = false;
  if (!configured) {
    RV_CUDA(cudaFuncSetAttribute(siglip_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kAttnSmemBytes));
    // two CTAs per SM need the full 228 KB shared-memory carveout
    RV_CUDA(cudaFuncSetAttribute(siglip_attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    RV_CUDA(cudaFuncSetAttribute(siglip_attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kPpSmemBytes));
    configured = true;
  }
  const uint64_t th = static_cast<uint64_t>(tiles) * heads;
  CUtensorMap tq, tq2, tk, tk2, tv;
  const uint64_t pitch = static_cast<uint64_t>(hd_pad) * 2;
  st = make_tmap_bf16_2d(&tq, q, hd_pad, th * seq_pad, pitch, 64, kAttnBQ, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != RADVLM_OK) return st;
  st = make_tmap_bf16_2d(&tq2, q, hd_pad, th * seq_pad, pitch, 16, kAttnBQ, CU_TENSOR_MAP_SWIZZLE_32B);
  if (st != RADVLM_OK) return st;
  st = make_tmap_bf16_2d(&tk, k, hd_pad, th * seq_pad, pitch, 64, kAttnBKV, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != RADVLM_OK) return st;
  st = make_tmap_bf16_2d(&tk2, k, hd_pad, th * seq_pad, pitch, 16, kAttnBKV, CU_TENSOR_MAP_SWIZZLE_32B);
  if (st != RADVLM_OK) return st;
  st = make_tmap_bf16_2d(&tv, vt, hd_pad, th * seq_pad, pitch, 16, kAttnBKV, CU_TENSOR_MAP_SWIZZLE_32B);
  if (st != RADVLM_OK) return st;
  AttnArgs a;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.seq = seq;
  a.seq_pad = seq_pad;
  a.heads = heads;
  a.hd = hd;
  a.scale_log2e = scale * 1.4426950408889634f;
  a.lse = lse;
  static const bool no_trim = [] { const char* e = std::getenv("RADVLM_B200_ATTN_TRIM"); return e && e[0] == '0'; }();
  a.trim_last = no_trim ? 0 : 1;   // A/B switch: RADVLM_B200_ATTN_TRIM=0 computes the padded last key block 96 wide
  // Tuning switch (bring-up A/B only): RADVLM_B200_ATTN=2cta selects the two-CTAs-per-SM kernel.
  static const bool use_2cta = [] { const char* e = std::getenv("RADVLM_B200_ATTN"); return e && e[0] == '2'; }();
  if (use_2cta) {
    a.num_qblk = (seq + kAttnBQ - 1) / kAttnBQ;
    a.total_items = tiles * heads * a.num_qblk;
    const int grid = std::min(a.total_items, 2 * device_sm_count());  // persistent: two resident CTAs per SM
    siglip_attention_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tq, tq2, tk, tk2, tv, a);
  } else {
    a.num_qblk = (seq + kPpItemRows - 1) / kPpItemRows;
    a.total_items = tiles * heads * a.num_qblk;
    const int grid = std::min(a.total_items, device_sm_count());  // persistent: one CTA (two query groups) per SM
    CUtensorMap tout;  // out viewed as [tiles][seq][heads*hd]: the store of a 128-row tile is clipped at seq
    const uint64_t row_bytes = static_cast<uint64_t>(heads) * hd * 2;
    st = make_tmap_bf16_3d(&tout, out, static_cast<uint64_t>(heads) * hd, seq, tiles, row_bytes, row_bytes * seq, hd,
                           kAttnBQ, 1);
    if (st != RADVLM_OK) return st;
    RV_CUDA(launch_kernel_pdl(siglip_attention_pp_kernel, grid, kPpThreads, kPpSmemBytes, stream, pdl_enabled(), tq, tq2, tk,
                              tk2, tv, tout, a));
  }
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

}  // namespace rv

extern "C" int radvlm_attention_fwd(const void* q, const void* k, const void* vt, void* out, int tiles,
                                    int heads, int seq, int seq_pad, int hd, int hd_pad, float scale,
                                    void* stream) {
  return rv::attention_launch(q, k, vt, out, nullptr, tiles, heads, seq, seq_pad, hd, hd_pad, scale,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_attention_fwd_lse(const void* q, const void* k, const void* vt, void* out, float* lse, int tiles,
                                        int heads, int seq, int seq_pad, int hd, int hd_pad, float scale,
                                        void* stream) {
  return rv::attention_launch(q, k, vt, out, lse, tiles, heads, seq, seq_pad, hd, hd_pad, scale,
                              static_cast<cudaStream_t>(stream));
}
