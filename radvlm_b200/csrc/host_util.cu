#include "host_util.h"

#include <string.h>

namespace rv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

int require_sm100() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s (no CUDA device; this library has no CPU fallback)",
              cudaGetErrorString(e));
    return RADVLM_ERR_UNSUPPORTED_DEVICE;
  }
  static thread_local int cached_dev = -1;
  static thread_local int cached_ok = 0;
  if (cached_dev != dev) {
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cached_dev = dev;
    cached_ok = (major == 10);
    if (!cached_ok)
      set_error("device %d is sm_%d%d; radvlm_b200 kernels are built for sm_100a only", dev, major,
                minor);
  }
  return cached_ok ? RADVLM_OK : RADVLM_ERR_UNSUPPORTED_DEVICE;
}

int device_sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int sms = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                      CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return RADVLM_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (pitch_bytes & 15u) != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (base=%p pitch=%llu)",
              base, static_cast<unsigned long long>(pitch_bytes));
    return RADVLM_ERR_BAD_ARGUMENT;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu box=%ux%u)",
              static_cast<int>(r), static_cast<unsigned long long>(inner),
              static_cast<unsigned long long>(outer), static_cast<unsigned long long>(pitch_bytes),
              box_inner, box_outer);
    return RADVLM_ERR_CUDA;
  }
  return RADVLM_OK;
}

}  // namespace rv

extern "C" const char* radvlm_last_error(void) { return rv::last_error(); }
extern "C" int radvlm_abi_version(void) { return 1; }
