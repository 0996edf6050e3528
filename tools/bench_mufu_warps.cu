// Micro-benchmark (bring-up tool, not product): how many warps per scheduler must issue MUFU.EX2 to keep the XU pipe
// busy, with and without an FFMA-only "filler" warp next to them (the situation of the ping-pong attention kernel:
// one group's warps burst exponentials while the other group's warps do MUFU-free work on the same schedulers).
// Prints cycles per MUFU warp-instruction per scheduler (the pipe floor is 8: 16 ex2 / clk / SM).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int kRounds = 512;
constexpr int kUnroll = 16;

// mufu_warps / filler_warps: per scheduler (SMSP).  mode 0: bare ex2 chains; 1: ffma -> ex2; 2: ffma -> ex2 -> pack
// of the pair right behind it (the order ptxas produced in the kernel).
// Own-warp interleave: every ex2 is followed by NF independent FFMAs of the SAME warp (8 separate chains).
template <int NF>
__global__ void __launch_bounds__(1024) k_own(long long* cycles, float* out, float seed) {
  float x[kUnroll], y[8];
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) x[i] = seed * (threadIdx.x + i + 1) * 1e-4f;
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = seed * (threadIdx.x + i + 3) * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kRounds; ++it) {
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int f = 0; f < NF; ++f) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0fBA83126F;" : "+f"(y[(i * NF + f) & 7]));
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) acc += x[i] + y[i & 7];
  if (acc == 123.456f) out[0] = acc;
}

// Same with the packed half2 exponential: one MUFU instruction produces two results (two XU passes).
template <int NF>
__global__ void __launch_bounds__(1024) k_own_h2(long long* cycles, float* out, float seed) {
  uint32_t x[kUnroll];
  float y[8];
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) x[i] = 0x3c003c00u + threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = seed * (threadIdx.x + i + 3) * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kRounds; ++it) {
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
#pragma unroll
      for (int f = 0; f < NF; ++f) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0fBA83126F;" : "+f"(y[(i * NF + f) & 7]));
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) acc += __uint_as_float(x[i]) + y[i & 7];
  if (acc == 123.456f) out[0] = acc;
}

template <int NF>
void run_own_h2(int warps, int sms, long long* dc, float* d) {
  k_own_h2<NF><<<sms, 128 * warps>>>(dc, d, 1.0f);
  cudaDeviceSynchronize();
  k_own_h2<NF><<<sms, 128 * warps>>>(dc, d, 1.0f);
  long long c = 0;
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("own-warp interleave, ex2.f16x2: %d FFMA per instr, %d warps/SMSP: %6.2f cycles per MUFU warp-instruction (2 results) per SMSP\n",
         NF, warps, double(c) / (double(kRounds) * kUnroll * warps));
}

template <int NF>
void run_own(int warps, int sms, long long* dc, float* d) {
  k_own<NF><<<sms, 128 * warps>>>(dc, d, 1.0f);
  cudaDeviceSynchronize();
  k_own<NF><<<sms, 128 * warps>>>(dc, d, 1.0f);
  long long c = 0;
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("own-warp interleave: %d FFMA per ex2, %d warps/SMSP: %6.2f cycles per MUFU warp-instruction per SMSP\n", NF, warps,
         double(c) / (double(kRounds) * kUnroll * warps));
}

template <int MODE>
__global__ void __launch_bounds__(1024) k(long long* cycles, float* out, int mufu_warps, float seed) {
  const int warp = threadIdx.x >> 5;
  const bool is_mufu = warp < 4 * mufu_warps;
  float x[kUnroll];
  uint32_t h[kUnroll / 2];
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) x[i] = seed * (threadIdx.x + i + 1) * 1e-4f;
#pragma unroll
  for (int i = 0; i < kUnroll / 2; ++i) h[i] = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (is_mufu) {
    for (int it = 0; it < kRounds; ++it) {
#pragma unroll
      for (int i = 0; i < kUnroll; ++i) {
        if (MODE == 0) {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        } else {
          const float a = fmaf(x[i], 0.999f, -0.001f);
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(a));
          if (MODE == 2 && (i & 1)) {
            uint32_t p;
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x[i]), "f"(x[i - 1]));
            h[i >> 1] |= p;
          }
        }
      }
    }
  } else {
    for (int it = 0; it < kRounds * 8; ++it) {
#pragma unroll
      for (int i = 0; i < kUnroll; ++i) x[i] = fmaf(x[i], 0.999f, -0.001f);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) acc += x[i] + __uint_as_float(h[i >> 1]);
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int mufu_warps, int filler_warps, int sms, long long* dc, float* d) {
  const int threads = 128 * (mufu_warps + filler_warps);
  k<MODE><<<sms, threads>>>(dc, d, mufu_warps, 1.0f);
  cudaDeviceSynchronize();
  k<MODE><<<sms, threads>>>(dc, d, mufu_warps, 1.0f);
  long long c = 0;
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s mufu warps/SMSP %d, filler warps/SMSP %d: %6.2f cycles per MUFU warp-instruction per SMSP\n", name,
         mufu_warps, filler_warps, double(c) / (double(kRounds) * kUnroll * mufu_warps));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* dc;
  float* d;
  cudaMalloc(&dc, 8 * sms);
  cudaMalloc(&d, 4);
  for (int mw = 1; mw <= 4; mw *= 2)
    for (int fw = 0; fw <= 2; ++fw) {
      run<0>("ex2 chains", mw, fw, sms, dc, d);
      run<1>("ffma -> ex2", mw, fw, sms, dc, d);
      run<2>("ffma -> ex2 -> pack", mw, fw, sms, dc, d);
    }
  for (int w = 1; w <= 4; w *= 2) {
    run_own<2>(w, sms, dc, d);
    run_own<4>(w, sms, dc, d);
    run_own<6>(w, sms, dc, d);
    run_own<8>(w, sms, dc, d);
    run_own<12>(w, sms, dc, d);
  }
  for (int w = 1; w <= 2; ++w) {
    run_own_h2<0>(w, sms, dc, d);
    run_own_h2<4>(w, sms, dc, d);
    run_own_h2<8>(w, sms, dc, d);
    run_own_h2<12>(w, sms, dc, d);
    run_own_h2<16>(w, sms, dc, d);
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}
