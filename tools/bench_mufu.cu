// Micro-benchmark (bring-up tool, not product): per-SM throughput of the instructions the attention softmax is
// made of, to decide how exp2 should be computed on sm_100a.  Prints elements / clock / SM for each variant.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int kIters = 4096;
constexpr int kUnroll = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
  float x[kUnroll];
  uint32_t h[kUnroll];
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) {
    x[i] = seed * (threadIdx.x + i + 1) * 1e-4f;
    h[i] = 0x3c003c00u + threadIdx.x + i;
  }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      if (MODE == 0) {  // f32 ex2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 1) {  // f16x2 ex2 (2 elements per instruction)
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 2) {  // cvt pack (2 elements per instruction)
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[(i + 1) % kUnroll]));
        x[i] = __uint_as_float(h[i]);
      } else if (MODE == 3) {  // softmax element, f32 path: ffma + ex2 (+ cvt per pair)
        float a = fmaf(x[i], 0.999f, -0.001f);
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(a));
        if (i & 1) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[i - 1]));
      } else if (MODE == 4) {  // softmax pair, f16x2 path: 2 ffma + cvt + ex2.f16x2 (2 elements per iteration of i)
        float a0 = fmaf(x[i], 0.999f, -0.001f);
        float a1 = fmaf(x[i], 0.998f, -0.002f);
        uint32_t p;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(a1), "f"(a0));
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(h[i]) : "r"(p));
        x[i] = __uint_as_float(h[i] & 0x3fffffffu);
      } else if (MODE == 5) {  // tanh.approx.f32
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 6) {  // ex2.approx.ftz.bf16x2
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 7) {  // fmax3-like: two fmax
        x[i] = fmaxf(fmaxf(x[i], x[(i + 1) % kUnroll]), seed);
      } else if (MODE == 8) {  // pure FFMA
        x[i] = fmaf(x[i], 0.999f, -0.001f);
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kUnroll; ++i) acc += x[i] + __uint_as_float(h[i]);
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int elems_per_instr, int sms, float* d) {
  const int blocks = sms * 4, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(d, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double elems = double(blocks) * threads * kIters * kUnroll * elems_per_instr;
  const double clk = ms * 1e-3 * khz * 1e3;
  printf("%-44s %8.3f ms  %7.2f elements/clk/SM (at the %d MHz attribute clock)\n", name, ms, elems / clk / sms,
         khz / 1000);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d;
  cudaMalloc(&d, 4);
  run<0>("ex2.approx.ftz.f32", 1, sms, d);
  run<1>("ex2.approx.ftz.f16x2 (2 el/instr)", 2, sms, d);
  run<6>("ex2.approx.ftz.bf16x2 (2 el/instr)", 2, sms, d);
  run<2>("cvt.rn.f16x2.f32 (2 el/instr)", 2, sms, d);
  run<5>("tanh.approx.f32", 1, sms, d);
  run<7>("2 x fmax", 1, sms, d);
  run<8>("ffma", 1, sms, d);
  run<3>("softmax elem f32: ffma+ex2+0.5cvt", 1, sms, d);
  run<4>("softmax pair f16x2: 2ffma+cvt+ex2.f16x2", 2, sms, d);
  cudaError_t err = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}
