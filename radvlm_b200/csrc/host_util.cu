#include "host_util.h"

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

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

bool pdl_enabled() {
  static const bool on = !(getenv("RADVLM_B200_PDL") && atoi(getenv("RADVLM_B200_PDL")) == 0);
  return on;
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

// ---- descriptor cache --------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a pure function of its arguments (it touches neither the device nor the memory behind
// `base`), and a step re-creates the same few hundred descriptors every time (workspace and weight addresses are
// stable: PackedWeights / the per-device workspace are allocated once).  A per-thread direct-mapped table keyed on the
// FULL argument tuple returns the encoded 128 bytes instead of calling the driver again: a hit can never hand back a
// descriptor that differs from what the driver would encode.  RADVLM_B200_TMAP_CACHE=0 disables it (A/B switch).
struct TmapKey {
  uint64_t base, dims[3], strides[2];
  uint32_t box[3], rank, swizzle, valid;
};
struct TmapSlot {
  TmapKey key;
  alignas(64) CUtensorMap map;
};
static const int kTmapSlots = 2048;   // power of two; ~420 KB per calling thread, allocated on first use
static thread_local TmapSlot* g_tmap_cache = nullptr;
static thread_local uint64_t g_tmap_hits = 0, g_tmap_misses = 0;

static bool tmap_cache_enabled() {
  static const bool on = !(getenv("RADVLM_B200_TMAP_CACHE") && atoi(getenv("RADVLM_B200_TMAP_CACHE")) == 0);
  return on;
}

static TmapSlot* tmap_slot(const TmapKey& k) {
  if (g_tmap_cache == nullptr) g_tmap_cache = new TmapSlot[kTmapSlots]();
  uint64_t h = 0x9E3779B97F4A7C15ull;
  const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
  for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) {
    h ^= w[i];
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 29;
  }
  return &g_tmap_cache[h & (kTmapSlots - 1)];
}

static int encode_tmap(CUtensorMap* out, const void* base, uint32_t rank, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box, CUtensorMapSwizzle swizzle, CUresult* res) {
  EncodeTiledFn fn = encode_fn();
  *res = CUDA_SUCCESS;
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return RADVLM_ERR_CUDA;
  }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = reinterpret_cast<uint64_t>(base);
  for (uint32_t i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (uint32_t i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
  key.rank = rank;
  key.swizzle = static_cast<uint32_t>(swizzle);
  key.valid = 1;
  TmapSlot* slot = tmap_cache_enabled() ? tmap_slot(key) : nullptr;
  if (slot != nullptr && !memcmp(&slot->key, &key, sizeof(key))) {
    memcpy(out, &slot->map, sizeof(CUtensorMap));
    ++g_tmap_hits;
    return RADVLM_OK;
  }
  cuuint32_t estr[3] = {1, 1, 1};
  *res = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (*res != CUDA_SUCCESS) return RADVLM_ERR_CUDA;
  ++g_tmap_misses;
  if (slot != nullptr) {
    slot->key = key;
    memcpy(&slot->map, out, sizeof(CUtensorMap));
  }
  return RADVLM_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                      CUtensorMapSwizzle swizzle) {
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (pitch_bytes & 15u) != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (base=%p pitch=%llu)",
              base, static_cast<unsigned long long>(pitch_bytes));
    return RADVLM_ERR_BAD_ARGUMENT;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  CUresult r;
  const int st = encode_tmap(out, base, 2, dims, strides, box, swizzle, &r);
  if (st != RADVLM_OK && r != CUDA_SUCCESS)
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu box=%ux%u)",
              static_cast<int>(r), static_cast<unsigned long long>(inner),
              static_cast<unsigned long long>(outer), static_cast<unsigned long long>(pitch_bytes),
              box_inner, box_outer);
  return st;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes,
                      uint64_t pitch2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (pitch1_bytes & 15u) != 0 || (pitch2_bytes & 15u) != 0 ||
      ((box0 * 2u) & 15u) != 0) {
    set_error("TMA operand must be 16-byte aligned with 16-byte-multiple pitches and box rows (base=%p pitches=%llu,%llu "
              "box0=%u)", base, static_cast<unsigned long long>(pitch1_bytes),
              static_cast<unsigned long long>(pitch2_bytes), box0);
    return RADVLM_ERR_BAD_ARGUMENT;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  CUresult r;
  const int st = encode_tmap(out, base, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, &r);
  if (st != RADVLM_OK && r != CUDA_SUCCESS)
    set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (dims=%llu,%llu,%llu box=%u,%u,%u)", static_cast<int>(r),
              static_cast<unsigned long long>(d0), static_cast<unsigned long long>(d1),
              static_cast<unsigned long long>(d2), box0, box1, box2);
  return st;
}

void tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  *hits = g_tmap_hits;
  *misses = g_tmap_misses;
}

// ---------------------------------------------------------------------------------------------
// profiling
// ---------------------------------------------------------------------------------------------
struct ProfRec {
  cudaEvent_t beg, end;
  int cls;
  int launches;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;      // recorded scopes since the last reset
static std::vector<ProfRec> g_prof_pool; // reusable event pairs
static std::mutex g_prof_mu;

ProfScope::ProfScope(int cls, cudaStream_t stream, int launches) : idx_(-1), stream_(stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  if (!g_prof_pool.empty()) {
    r = g_prof_pool.back();
    g_prof_pool.pop_back();
  } else {
    if (cudaEventCreate(&r.beg) != cudaSuccess || cudaEventCreate(&r.end) != cudaSuccess) return;
  }
  r.cls = cls;
  r.launches = launches;
  cudaEventRecord(r.beg, stream);
  g_prof.push_back(r);
  idx_ = static_cast<int>(g_prof.size()) - 1;
}

ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx_ < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[idx_].end, stream_);
}

}  // namespace rv

extern "C" int radvlm_profile_enable(int on) {
  rv::g_prof_on = (on != 0);
  return RADVLM_OK;
}

// Sums the recorded scopes per class (waits for their completion), then clears the record list.
extern "C" int radvlm_profile_read(float* ms_per_class, int64_t* launches_per_class, int n_classes) {
  std::lock_guard<std::mutex> lk(rv::g_prof_mu);
  for (int i = 0; i < n_classes; ++i) {
    if (ms_per_class) ms_per_class[i] = 0.f;
    if (launches_per_class) launches_per_class[i] = 0;
  }
  for (auto& r : rv::g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.end) == cudaSuccess) cudaEventElapsedTime(&ms, r.beg, r.end);
    if (r.cls >= 0 && r.cls < n_classes) {
      if (ms_per_class) ms_per_class[r.cls] += ms;
      if (launches_per_class) launches_per_class[r.cls] += r.launches;
    }
    rv::g_prof_pool.push_back(r);
  }
  rv::g_prof.clear();
  return RADVLM_OK;
}

extern "C" int radvlm_tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  uint64_t h = 0, m = 0;
  rv::tmap_cache_stats(&h, &m);
  if (hits) *hits = h;
  if (misses) *misses = m;
  return RADVLM_OK;
}

extern "C" const char* radvlm_last_error(void) { return rv::last_error(); }
extern "C" int radvlm_abi_version(void) { return 3; }
