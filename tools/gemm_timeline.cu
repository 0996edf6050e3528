// Tuning tool: per-tile timeline of the scheduled CTA-pair GEMM (MMA issuer and one epilogue warp of CTA pair 0).
// Needs a library built with -DRV_GEMM_TIMELINE=0 (tools/build_variant.sh gtl -DRV_GEMM_TIMELINE=0), run with
// LD_LIBRARY_PATH=build/var_gtl.  Stamps per tile: MMA warp {loop top, accumulator free, first operands landed, last
// MMA issued}, epilogue warp {stats fetched, accumulator full, drained}.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../include/radvlm_b200.h"
extern "C" void radvlm_gemm_set_timeline(long long* p);

static void run(int M, int N, int K, int epi, bool ln, bool zeros = false) {
  __nv_bfloat16 *A, *W;
  void* out;
  float *b, *s, *stats, *aux;
  long long* tl;
  cudaMalloc(&A, (size_t)M * K * 2); cudaMalloc(&W, (size_t)N * K * 2); cudaMalloc(&out, (size_t)M * N * 4);
  cudaMalloc(&b, N * 4); cudaMalloc(&s, N * 4); cudaMalloc(&stats, (size_t)M * 8); cudaMalloc(&aux, (size_t)M * N * 4);
  cudaMalloc(&tl, 24 * 8 * 8);
  std::vector<__nv_bfloat16> h((size_t)M * K);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16(zeros ? 0.f : (float)((i * 2654435761u) >> 20 & 255) / 256.f - 0.5f);
  cudaMemcpy(A, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  for (size_t off = 0; off < (size_t)N * K; off += h.size())
    cudaMemcpy(W + off, h.data(), (((size_t)N * K - off) < h.size() ? ((size_t)N * K - off) : h.size()) * 2, cudaMemcpyHostToDevice);
  cudaMemset(b, 0, N * 4); cudaMemset(s, 0, N * 4); cudaMemset(stats, 0, (size_t)M * 8); cudaMemset(aux, 0, (size_t)M * N * 4);
  cudaMemset(tl, 0, 24 * 64);
  radvlm_gemm_set_timeline(tl);
  for (int rep = 0; rep < 3; ++rep) {
    int st = ln ? radvlm_gemm_bf16_ln(A, K, W, K, M, N, K, b, stats, s, epi, out, N, nullptr)
                : radvlm_gemm_bf16(A, K, W, K, M, N, K, b, epi, out, N, aux, 0, 0, nullptr);
    if (st) { printf("error %s\n", radvlm_last_error()); exit(1); }
  }
  cudaDeviceSynchronize();
  long long t[24 * 8];
  cudaMemcpy(t, tl, sizeof(t), cudaMemcpyDeviceToHost);
  printf("%sM=%d N=%d K=%d epi=%d ln=%d : cycles from the first stamp (CTA pair 0)\n", zeros ? "ZERO DATA " : "", M, N, K, epi, (int)ln);
  const long long t0 = t[0];
  for (int e = 0; e < 12; ++e) {
    const long long* x = t + e * 8;
    if (x[3] == 0) break;
    printf("  tile %2d  mma: top %7lld acc_free %7lld first_ops %7lld issued %7lld (issue span %6lld) | epi: ready %7lld full %7lld drained %7lld (drain %6lld)\n",
           e, x[0] - t0, x[1] - t0, x[2] - t0, x[3] - t0, x[3] - x[2], x[4] - t0, x[5] - t0, x[6] - t0, x[6] - x[5]);
  }
  cudaFree(A); cudaFree(W); cudaFree(out); cudaFree(b); cudaFree(s); cudaFree(stats); cudaFree(aux); cudaFree(tl);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'o') {   // out_proj shape: residual epilogue vs the same stores without the residual read
    run(58320, 1152, 1152, RADVLM_EPI_RESID_F32, false);
    run(58320, 1152, 1152, RADVLM_EPI_BIAS_F32, false);
    run(58320, 1152, 1152, RADVLM_EPI_BIAS_BF16, false);
    return 0;
  }
  if (argc > 1) { run(58320, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, false); return 0; }
  run(58320, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, false, true);
  run(58320, 4304, 1152, RADVLM_EPI_BIAS_BF16, false, true);
  run(58320, 4304, 1152, RADVLM_EPI_BIAS_BF16, false);
  run(58320, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, false);
  run(58320, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, true);
  run(58320, 3456, 1152, RADVLM_EPI_BIAS_BF16, false);
  run(58320, 1152, 1152, RADVLM_EPI_RESID_F32, false);
  run(58320, 1152, 4304, RADVLM_EPI_RESID_F32, false);
  return 0;
}
