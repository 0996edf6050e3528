// Host-side helpers shared by the C-ABI translation units: status codes, thread-local error string,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/radvlm_b200.h"

namespace rv {

void set_error(const char* fmt, ...);
const char* last_error();

#define RV_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      rv::set_error(__VA_ARGS__);      \
      return RADVLM_ERR_BAD_ARGUMENT;  \
    }                                  \
  } while (0)

#define RV_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      rv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                    __LINE__);                                                         \
      return RADVLM_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

// Returns 0 when the current device is sm_100 (B200); otherwise sets the error and returns a status.
int require_sm100();
int device_sm_count();
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows, row pitch in bytes.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                      CUtensorMapSwizzle swizzle);  // (any 16-bit element type: the copy is bit-exact)

// 3-D bf16 tensor map (d0 contiguous; byte pitches of d1 and d2), no swizzle.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes,
                      uint64_t pitch2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);

// RADVLM_B200_PDL (default on; 0 = ordinary launches): programmatic dependent launch of the persistent tensor-core
// kernels (common.cuh griddep_wait).
bool pdl_enabled();

// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute when `pdl` is set.
template <typename... Params, typename... Args>
static inline cudaError_t launch_kernel_pdl(void (*kernel)(Params...), unsigned grid, unsigned block, size_t smem,
                                            cudaStream_t stream, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

// Hit / miss counts of the calling thread's descriptor cache (host_util.cu).
void tmap_cache_stats(uint64_t* hits, uint64_t* misses);

// Optional per-kernel-class device timing (CUDA events recorded on the launch stream).  Off by default;
// bench.py turns it on to report the live duration / launch count of each kernel class.
enum ProfClass : int {
  PROF_GEMM = 0,        // patch embed + projector GEMMs
  PROF_ATTENTION = 1,
  PROF_LAYERNORM = 2,
  PROF_MISC = 3,        // im2col, cast, padding memsets
  PROF_PREPROCESS = 4,
  PROF_MERGE_SPLICE = 5,
  PROF_GEMM_QKV = 6,
  PROF_GEMM_OUT = 7,
  PROF_GEMM_FC1 = 8,
  PROF_GEMM_FC2 = 9,
  PROF_BWD_RECOMPUTE = 10,  // training mode: forward pieces recomputed inside the backward (LN, QKV, out_proj, fc1)
  PROF_BWD_DGRAD = 11,      // dX = dY W
  PROF_BWD_WGRAD = 12,      // dW += dY^T X (+ bias column sums)
  PROF_BWD_ATTENTION = 13,
  PROF_BWD_ELEMENTWISE = 14, // LayerNorm / GELU backward, casts, position-table sum
  PROF_NUM_CLASSES = 15
};
struct ProfScope {
  ProfScope(int cls, cudaStream_t stream, int launches = 1);
  ~ProfScope();
  int idx_;
  cudaStream_t stream_;
};

}  // namespace rv
