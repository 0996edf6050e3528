// Micro-benchmark (bring-up tool, not product): issue-to-completion cost of back-to-back tcgen05.mma of the shapes
// the attention kernel uses, with operands already resident in shared / tensor memory.  One thread issues
// `reps` MMAs, commits, waits; cycles / MMA is printed per variant.  Run with 1 and 2 CTAs per SM.
#include <cstdio>
#include <cstdint>
#include "../radvlm_b200/csrc/common.cuh"

using namespace rv;

struct Variant {
  int n;        // MMA N
  int ts;       // 1: A from TMEM, 0: A from smem
  int layout;   // kLayoutSw128 / kLayoutSw32
  int sbo;      // stride between 8-row groups
  int accumulate_same_d;  // 1: all MMAs accumulate into the same D columns; 0: alternate two D regions
};

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
  return pred;
}

// STYLE 0: `if (threadIdx.x == 0)` single-thread issue loop.  STYLE 1: the whole warp 0 runs the loop convergently and
// an elected lane issues (operands stay warp-uniform).
template <int STYLE>
__global__ void __launch_bounds__(128) umma_bench(Variant v, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768;
  const uint32_t bar = base + 65536, tptr = base + 65536 + 16;
  for (uint32_t i = threadIdx.x; i < 65536 / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0u));
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(tptr));
  const uint32_t warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (STYLE == 0 ? (threadIdx.x == 0) : (warp_u == 0)) {
    if (STYLE == 1) tm = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t idesc = make_idesc_bf16(128, v.n);
    const uint64_t ad = make_smem_desc(sA, v.sbo, v.layout);
    const uint64_t bd = make_smem_desc(sB, v.sbo, v.layout);
    // warm-up
    for (int i = 0; i < 8; ++i)
      if (STYLE == 0 || elect_one()) umma_bf16_ss(tm, ad, bd, idesc, 1);
    if (STYLE == 0 || elect_one()) umma_commit(bar);
    mbar_wait(bar, 0);
    tc_fence_after();
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const uint32_t d = tm + (v.accumulate_same_d ? 0u : static_cast<uint32_t>((i & 1) * 128));
      if (STYLE == 0 || elect_one()) {
        if (v.ts) umma_bf16_ts(d, tm + 128 + (i & 7) * 8, bd + (i & 3) * 2, idesc, 1);
        else umma_bf16_ss(d, ad + (i & 3) * 2, bd + (i & 3) * 2, idesc, 1);
      }
    }
    const long long t1 = clock64();
    if (STYLE == 0 || elect_one()) umma_commit(bar);
    mbar_wait(bar, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

// Two warps of ONE CTA issue independent MMA streams (warp 0: SS N=128 into D0, warp 1: TS N=80 into D1), the
// situation of a CTA whose S = Q K^T and O += P V products are issued by different warps.
__global__ void __launch_bounds__(128) umma_dual(int reps, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768;
  const uint32_t bar = base + 65536, tptr = base + 65536 + 32;
  for (uint32_t i = threadIdx.x; i < 65536 / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0u));
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(tptr));
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < 2 && (mode == 2 || w == mode)) {
    const uint64_t ad = make_smem_desc(sA, 1024, kLayoutSw128);
    const uint64_t bd = make_smem_desc(sB, 1024, kLayoutSw128);
    const long long t0 = clock64();
    if (w == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 128);
      for (int i = 0; i < reps; ++i) umma_bf16_ss(tm, ad + (i & 3) * 2, bd + (i & 3) * 2, idesc, 1);
    } else {
      const uint32_t idesc = make_idesc_bf16(128, 80);
      for (int i = 0; i < reps; ++i) umma_bf16_ts(tm + 128, tm + 256 + (i & 7) * 8, bd + (i & 3) * 2, idesc, 1);
    }
    umma_commit(bar + 8 * w);
    mbar_wait(bar + 8 * w, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) out[w] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}


// Whole-warp (convergent) issue: elect.sync picks the issuing lane inside the asm block (common.cuh *_elect).
__global__ void __launch_bounds__(128) umma_dual_elect(int reps, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768;
  const uint32_t bar = base + 65536, tptr = base + 65536 + 32;
  for (uint32_t i = threadIdx.x; i < 65536 / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0u));
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(tptr));
  tm = __shfl_sync(0xffffffffu, tm, 0);
  const int w = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (w < 2 && (mode == 2 || w == mode)) {
    const uint64_t ad = make_smem_desc(sA, 1024, kLayoutSw128);
    const uint64_t bd = make_smem_desc(sB, 1024, kLayoutSw128);
    const long long t0 = clock64();
    if (w == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 128);
      for (int i = 0; i < reps; ++i) umma_bf16_ss_elect(tm, ad + (i & 3) * 2, bd + (i & 3) * 2, idesc, 1);
    } else {
      const uint32_t idesc = make_idesc_bf16(128, 80);
      for (int i = 0; i < reps; ++i) umma_bf16_ts_elect(tm + 128, tm + 256 + (i & 7) * 8, bd + (i & 3) * 2, idesc, 1);
    }
    umma_commit_elect(bar + 8 * w);
    mbar_wait(bar + 8 * w, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[w] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 65536 + 1024 + 64;
  cudaFuncSetAttribute(umma_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(umma_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2048;
  struct { const char* name; Variant v; } vs[] = {
      {"SS M128 N128 K16 SW128 (S = Q K^T step)", {128, 0, (int)kLayoutSw128, 1024, 1}},
      {"SS M128 N128 K16 SW128, alternating D", {128, 0, (int)kLayoutSw128, 1024, 0}},
      {"SS M128 N128 K16 SW32", {128, 0, (int)kLayoutSw32, 256, 1}},
      {"SS M128 N80  K16 SW128 (P from smem)", {80, 0, (int)kLayoutSw128, 1024, 1}},
      {"TS M128 N80  K16 (P from TMEM)", {80, 1, (int)kLayoutSw128, 1024, 1}},
      {"TS M128 N128 K16", {128, 1, (int)kLayoutSw128, 1024, 1}},
      {"SS M128 N256 K16 SW128 (GEMM step)", {256, 0, (int)kLayoutSw128, 1024, 1}},
      {"SS M128 N64  K16 SW128", {64, 0, (int)kLayoutSw128, 1024, 1}},
  };
  for (int style = 0; style < 2; ++style)
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
    printf("--- issue style %d (%s), %d CTA(s) per SM issuing concurrently\n", style,
           style ? "convergent warp + elect.sync" : "if (threadIdx.x == 0)", ctas_per_sm);
    for (auto& e : vs) {
      if (style) umma_bench<1><<<148 * ctas_per_sm, 128, smem>>>(e.v, reps, d);
      else umma_bench<0><<<148 * ctas_per_sm, 128, smem>>>(e.v, reps, d);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("%s: %s\n", e.name, cudaGetErrorString(err)); return 1; }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const double flop = 2.0 * 128 * e.v.n * 16;
      printf("%-44s issue %7.1f cyc/MMA   complete %7.1f cyc/MMA   (%.0f FLOP/clk/CTA)\n", e.name,
             (double)h[0] / reps, (double)h[1] / reps, flop * reps / h[1]);
    }
  }
  cudaFuncSetAttribute(umma_dual, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* mn[3] = {"warp0 only (SS N128)", "warp1 only (TS N80)", "both warps concurrently"};
  cudaFuncSetAttribute(umma_dual_elect, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int pass = 0; pass < 2; ++pass)
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 0) printf("--- %s\n", pass ? "whole warp + elect.sync inside the asm block" : "single thread (lane 0) issue");
    cudaMemset(d, 0, 16);
    if (pass) umma_dual_elect<<<148, 128, smem>>>(reps, mode, d);
    else umma_dual<<<148, 128, smem>>>(reps, mode, d);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("dual: %s\n", cudaGetErrorString(err)); return 1; }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("one CTA, %-26s SS stream %7.1f cyc/MMA   TS stream %7.1f cyc/MMA\n", mn[mode], (double)h[0] / reps,
           (double)h[1] / reps);
  }
  return 0;
}
