// Standalone bring-up test for the tcgen05 GEMM through the C ABI (no torch needed on the GPU box).
//   build: see Makefile (`make tools`)     run: ./build/test_gemm [quick]
// Compares against a naive CUDA-core fp32 GEMM on the same bf16 inputs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/radvlm_b200.h"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, float* C,
                         int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc += __bfloat162float(A[(size_t)m * K + k]) * __bfloat162float(W[(size_t)n * K + k]);
  C[(size_t)m * N + n] = acc + (bias ? bias[n] : 0.f);
}

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

static float gelu_tanh_h(float x) {
  return 0.5f * x * (1.f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
}
static float gelu_erf_h(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }

struct Case {
  int M, N, K, epi, bn;
};

static int run_case(const Case& c) {
  const int M = c.M, N = c.N, K = c.K;
  std::vector<__nv_bfloat16> hA((size_t)M * K), hW((size_t)N * K);
  std::vector<float> hb(N), haux;
  for (auto& v : hA) v = __float2bfloat16(frand());
  for (auto& v : hW) v = __float2bfloat16(frand());
  for (auto& v : hb) v = frand();
  const int period = 7;
  if (c.epi == RADVLM_EPI_RESID_F32) haux.resize((size_t)M * N);
  if (c.epi == RADVLM_EPI_POS_F32) haux.resize((size_t)period * N);
  for (auto& v : haux) v = frand();

  __nv_bfloat16 *dA, *dW;
  float *db, *dref, *daux = nullptr;
  void* dout;
  const bool out_f32 = (c.epi == RADVLM_EPI_RESID_F32 || c.epi == RADVLM_EPI_POS_F32 ||
                        c.epi == RADVLM_EPI_BIAS_F32);
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&db, N * 4));
  CK(cudaMalloc(&dref, (size_t)M * N * 4));
  CK(cudaMalloc(&dout, (size_t)M * N * 4));
  CK(cudaMemset(dout, 0xFF, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), N * 4, cudaMemcpyHostToDevice));
  if (!haux.empty()) {
    CK(cudaMalloc(&daux, haux.size() * 4));
    CK(cudaMemcpy(daux, haux.data(), haux.size() * 4, cudaMemcpyHostToDevice));
  }
  ref_gemm<<<dim3((N + 127) / 128, M), 128>>>(dA, dW, db, dref, M, N, K);
  CK(cudaGetLastError());

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int st = 0;
  const int reps = (M >= 4096) ? 5 : 1;
  // for the residual epilogue run out-of-place (aux != out) so repetitions stay idempotent
  for (int i = 0; i < reps + 1; ++i) {
    if (i == 1) cudaEventRecord(e0);
    st = radvlm_gemm_bf16(dA, K, dW, K, M, N, K, db, c.epi, dout, N, daux, period, c.bn, nullptr);
    if (st != 0) break;
  }
  cudaEventRecord(e1);
  if (st != 0) {
    printf("case M=%d N=%d K=%d epi=%d bn=%d: API error %d: %s\n", M, N, K, c.epi, c.bn, st,
           radvlm_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("case M=%d N=%d K=%d epi=%d bn=%d: kernel failed: %s\n", M, N, K, c.epi, c.bn,
           cudaGetErrorString(e));
    exit(3);
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= (reps > 0 ? reps : 1);

  std::vector<float> href((size_t)M * N), hout((size_t)M * N);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  if (out_f32) {
    CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> tmp((size_t)M * N);
    CK(cudaMemcpy(tmp.data(), dout, tmp.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) hout[i] = __bfloat162float(tmp[i]);
  }
  double max_err = 0, max_ref = 0;
  size_t bad = 0, worst = 0;
  for (size_t i = 0; i < href.size(); ++i) {
    float r = href[i];
    const int row = (int)(i / N), col = (int)(i % N);
    if (c.epi == RADVLM_EPI_GELU_TANH_BF16) r = gelu_tanh_h(r);
    if (c.epi == RADVLM_EPI_GELU_ERF_BF16) r = gelu_erf_h(r);
    if (c.epi == RADVLM_EPI_RESID_F32) r += haux[i];
    if (c.epi == RADVLM_EPI_POS_F32) r += haux[(size_t)(row % period) * N + col];
    const double tol = out_f32 ? 2e-3 : (2e-3 + fabs(r) * 8e-3);
    const double err = fabs((double)hout[i] - (double)r);
    if (!(err <= tol)) ++bad;
    if (err > max_err || isnan(hout[i])) {
      max_err = err;
      worst = i;
    }
    if (fabs(r) > max_ref) max_ref = fabs(r);
  }
  const double tflops = 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12;
  printf("case M=%5d N=%5d K=%5d epi=%d bn=%3d : max_err=%.3e max_ref=%.3e bad=%zu/%zu  %.3f ms %.1f TFLOP/s %s\n",
         M, N, K, c.epi, c.bn, max_err, max_ref, bad, href.size(), ms, tflops, bad ? "FAIL" : "ok");
  if (bad) {
    printf("  worst at (%zu,%zu): got %f want %f\n", worst / N, worst % N, hout[worst], href[worst]);
    printf("  got[0,0..7]:");
    for (int j = 0; j < 8 && j < N; ++j) printf(" %8.4f", hout[j]);
    printf("\n  ref[0,0..7]:");
    for (int j = 0; j < 8 && j < N; ++j) printf(" %8.4f", href[j]);
    printf("\n");
  }
  cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dref); cudaFree(dout);
  if (daux) cudaFree(daux);
  return bad ? 1 : 0;
}

// Backward-form GEMM: operands stored transposed (MN-major tcgen05 operands), optional split-K with the atomic epilogue.
static int run_case_ex(int M, int N, int K, int a_layout, int b_layout, int k_splits, int epi) {
  std::vector<__nv_bfloat16> hA((size_t)M * K), hW((size_t)N * K);   // logical row-major [M,K], [N,K]
  for (auto& v : hA) v = __float2bfloat16(frand());
  for (auto& v : hW) v = __float2bfloat16(frand());
  const int lda = a_layout ? (M + 7) / 8 * 8 : K, ldw = b_layout ? (N + 7) / 8 * 8 : K;
  std::vector<__nv_bfloat16> sA(a_layout ? (size_t)K * lda : (size_t)M * K, __float2bfloat16(0.f));
  std::vector<__nv_bfloat16> sW(b_layout ? (size_t)K * ldw : (size_t)N * K, __float2bfloat16(0.f));
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) sA[a_layout ? (size_t)k * lda + m : (size_t)m * K + k] = hA[(size_t)m * K + k];
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) sW[b_layout ? (size_t)k * ldw + n : (size_t)n * K + k] = hW[(size_t)n * K + k];
  __nv_bfloat16 *dA, *dW, *dAs, *dWs;
  float *dref, *dout;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dAs, sA.size() * 2)); CK(cudaMalloc(&dWs, sW.size() * 2));
  CK(cudaMalloc(&dref, (size_t)M * N * 4)); CK(cudaMalloc(&dout, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAs, sA.data(), sA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dWs, sW.data(), sW.size() * 2, cudaMemcpyHostToDevice));
  ref_gemm<<<dim3((N + 127) / 128, M), 128>>>(dA, dW, nullptr, dref, M, N, K);
  CK(cudaGetLastError());
  const bool atomic = (epi == RADVLM_EPI_ATOMIC_F32);
  const float init = atomic ? 1.5f : -7.f;
  std::vector<float> hinit((size_t)M * N, init);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 3;
  int st = 0;
  for (int i = 0; i < reps; ++i) {
    CK(cudaMemcpy(dout, hinit.data(), hinit.size() * 4, cudaMemcpyHostToDevice));
    if (i == reps - 1) cudaEventRecord(e0);
    st = radvlm_gemm_bf16_ex(dAs, lda, a_layout, dWs, ldw, b_layout, M, N, K, nullptr, epi, dout, N, nullptr, 0, k_splits, nullptr);
    if (st) break;
  }
  cudaEventRecord(e1);
  if (st) { printf("ex case: API error %d: %s\n", st, radvlm_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ex case kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<float> href((size_t)M * N), hout((size_t)M * N);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0; size_t bad = 0;
  for (size_t i = 0; i < href.size(); ++i) {
    const double want = href[i] + (atomic ? init : 0.f);
    const double err = fabs(hout[i] - want);
    if (!(err <= 2e-3 + 2e-4 * fabs(want))) ++bad;
    if (err > max_err) max_err = err;
    if (fabs(want) > max_ref) max_ref = fabs(want);
  }
  printf("ex   M=%5d N=%5d K=%6d a_layout=%d b_layout=%d splits=%d epi=%d : max_err=%.3e max_ref=%.3e bad=%zu/%zu  %.3f ms %.1f TFLOP/s %s\n",
         M, N, K, a_layout, b_layout, k_splits, epi, max_err, max_ref, bad, href.size(), ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12,
         bad ? "FAIL" : "ok");
  cudaFree(dA); cudaFree(dW); cudaFree(dAs); cudaFree(dWs); cudaFree(dref); cudaFree(dout);
  return bad ? 1 : 0;
}

// LayerNorm-fold epilogue cost: the same GEMM with and without (mean, rstd, row sums) applied in the epilogue
static void run_ln_cost(int M, int N, int K, int epi) {
  __nv_bfloat16 *dA, *dW, *dout;
  float *db, *ds, *dstats;
  CK(cudaMalloc(&dA, (size_t)M * K * 2)); CK(cudaMalloc(&dW, (size_t)N * K * 2)); CK(cudaMalloc(&dout, (size_t)M * N * 2));
  CK(cudaMalloc(&db, N * 4)); CK(cudaMalloc(&ds, N * 4)); CK(cudaMalloc(&dstats, (size_t)M * 8));
  CK(cudaMemset(dA, 0, (size_t)M * K * 2)); CK(cudaMemset(dW, 0, (size_t)N * K * 2));
  CK(cudaMemset(db, 0, N * 4)); CK(cudaMemset(ds, 0, N * 4)); CK(cudaMemset(dstats, 0, (size_t)M * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      float ms = 0;
      for (int i = 0; i < 11; ++i) {
        if (i == 1) cudaEventRecord(e0);
        int st = mode ? radvlm_gemm_bf16_ln(dA, K, dW, K, M, N, K, db, dstats, ds, epi, dout, N, nullptr)
                      : radvlm_gemm_bf16(dA, K, dW, K, M, N, K, db, epi, dout, N, nullptr, 0, 0, nullptr);
        if (st) { printf("API error %s\n", radvlm_last_error()); exit(2); }
      }
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms, e0, e1);
      printf("ln-cost M=%d N=%d K=%d epi=%d %s : %.4f ms per launch\n", M, N, K, epi, mode ? "folded" : "plain ", ms / 10);
    }
  }
  cudaFree(dA); cudaFree(dW); cudaFree(dout); cudaFree(db); cudaFree(ds); cudaFree(dstats);
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "lncost")) {
    run_ln_cost(58320, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16);
    run_ln_cost(58320, 3456, 1152, RADVLM_EPI_BIAS_BF16);
    return 0;
  }
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  std::vector<Case> cases = {
      {128, 128, 64, RADVLM_EPI_BIAS_F32, 128},    // one tile, one K slab
      {128, 128, 256, RADVLM_EPI_BIAS_F32, 128},   // K loop within one stage ring
      {128, 256, 1152, RADVLM_EPI_BIAS_F32, 256},  // ring wrap-around
      {128, 192, 1152, RADVLM_EPI_BIAS_F32, 192},
      {300, 200, 136, RADVLM_EPI_BIAS_F32, 128},   // ragged M/N/K tails
      {729, 1152, 588 + 4, RADVLM_EPI_POS_F32, 0},
      {1458, 1152, 1152, RADVLM_EPI_RESID_F32, 0},
      {1458, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, 0},
      {1458, 1152, 4304, RADVLM_EPI_BIAS_BF16, 0},
      {1458, 3584, 1152, RADVLM_EPI_GELU_ERF_BF16, 0},
  };
  if (!quick) {
    cases.push_back({7290, 3456, 1152, RADVLM_EPI_BIAS_BF16, 256});
    cases.push_back({7290, 3456, 1152, RADVLM_EPI_BIAS_BF16, 192});
    cases.push_back({7290, 3456, 1152, RADVLM_EPI_BIAS_BF16, 128});
    cases.push_back({7290, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, 0});
    cases.push_back({7290, 1152, 4304, RADVLM_EPI_RESID_F32, 0});
    cases.push_back({7290, 3584, 3584, RADVLM_EPI_BIAS_BF16, 0});
    cases.push_back({29160, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, 256});
    cases.push_back({29160, 1152, 4304, RADVLM_EPI_RESID_F32, 0});
  }
  if (argc > 1 && !strcmp(argv[1], "bn")) {  // ./test_gemm bn : tile width sweep on the N = 1152 GEMMs of a 80-tile call
    radvlm_gemm_set_mode(2);
    int f = 0;
    for (int rep = 0; rep < 2; ++rep)
      for (int bn : {256, 192, 128}) {
        f += run_case({58320, 1152, 4304, RADVLM_EPI_RESID_F32, bn});
        f += run_case({58320, 1152, 1152, RADVLM_EPI_RESID_F32, bn});
      }
    return f;
  }
  if (argc > 1 && !strcmp(argv[1], "schedncu")) {  // ncu target: round-robin vs scheduled kernel on the fc1 / fc2 shapes
    int f = 0;
    for (int bn : {256, -1}) f += run_case({58320, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, bn});
    for (int bn : {192, -1}) f += run_case({58320, 1152, 4304, RADVLM_EPI_RESID_F32, bn});
    return f;
  }
  if (argc > 1 && !strcmp(argv[1], "sched")) {  // ./test_gemm sched : scheduled variable-width tiles (block_n = -1)
    int f = 0;                                   // correctness on ragged shapes, then the N = 1152 / 3456 shapes timed
    for (int epi : {RADVLM_EPI_BIAS_F32, RADVLM_EPI_RESID_F32, RADVLM_EPI_BIAS_BF16, RADVLM_EPI_GELU_TANH_BF16}) {
      f += run_case({300, 200, 136, epi, -1});
      f += run_case({1458, 1152, 1152, epi, -1});
      f += run_case({729, 384, 4304, epi, -1});
      f += run_case({2000, 648, 200, epi, -1});
    }
    f += run_case({729, 1152, 588 + 4, RADVLM_EPI_POS_F32, -1});
    f += run_case({7290, 3456, 1152, RADVLM_EPI_BIAS_BF16, -1});
    f += run_case({7290, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, -1});
    f += run_case({7290, 1152, 4304, RADVLM_EPI_RESID_F32, -1});
    // cost of a 128-wide tile against a full one: N = 128 (strips only) vs N = 256 (full tiles only), same M and K
    for (int K : {4304, 1152})
      for (int bn : {-1, 128, 256}) {
        f += run_case({58320, 128, K, RADVLM_EPI_RESID_F32, bn});
        f += run_case({58320, 256, K, RADVLM_EPI_RESID_F32, bn});
        f += run_case({58320, 1280, K, RADVLM_EPI_RESID_F32, bn});
      }
    for (int rep = 0; rep < 2; ++rep)
      for (int bn : {-1, 192, 256}) {
        f += run_case({58320, 1152, 4304, RADVLM_EPI_RESID_F32, bn});
        f += run_case({58320, 1152, 1152, RADVLM_EPI_RESID_F32, bn});
        f += run_case({58320, 3456, 1152, RADVLM_EPI_BIAS_BF16, bn});
        f += run_case({58320, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, bn});
        f += run_case({7290, 1152, 4304, RADVLM_EPI_RESID_F32, bn});
        f += run_case({7290, 1152, 1152, RADVLM_EPI_RESID_F32, bn});
      }
    printf("%s (%d failing cases)\n", f ? "SCHED GEMM TEST FAILED" : "SCHED GEMM TEST PASSED", f);
    return f;
  }
  if (argc > 2 && !strcmp(argv[1], "perf")) {  // ./test_gemm perf <mode> : the three big shapes only (ncu target)
    radvlm_gemm_set_mode(atoi(argv[2]));
    int f = 0;
    f += run_case({29160, 4304, 1152, RADVLM_EPI_GELU_TANH_BF16, 256});
    f += run_case({29160, 1152, 4304, RADVLM_EPI_RESID_F32, 0});
    f += run_case({29160, 3456, 1152, RADVLM_EPI_BIAS_BF16, 256});
    return f;
  }
  int fails = 0;
  for (int mode = 1; mode <= 2; ++mode) {
    radvlm_gemm_set_mode(mode);
    printf("---- GEMM mode %d (%s) ----\n", mode, mode == 1 ? "single-CTA 128xBN tiles" : "CTA-pair 256xBN tiles");
    for (const auto& c : cases) fails += run_case(c);
  }
  radvlm_gemm_set_mode(0);
  // backward forms: dgrad (W read as stored, contraction over its rows) and wgrad (both operands transposed, split-K)
  fails += run_case_ex(512, 256, 128, 0, 1, 1, RADVLM_EPI_BIAS_F32);
  fails += run_case_ex(512, 256, 128, 1, 0, 1, RADVLM_EPI_BIAS_F32);
  fails += run_case_ex(512, 256, 256, 1, 1, 1, RADVLM_EPI_BIAS_F32);
  fails += run_case_ex(1458, 1152, 4304, 0, 1, 1, RADVLM_EPI_BIAS_F32);     // dX = dY W   (fc1: out 4304 -> in 1152)
  fails += run_case_ex(1152, 4304, 1458, 1, 1, 1, RADVLM_EPI_ATOMIC_F32);   // dW2 += dY^T A
  fails += run_case_ex(4304, 1152, 1458, 1, 1, 3, RADVLM_EPI_ATOMIC_F32);   // dW1 += dU^T X, split-K 3
  fails += run_case_ex(300, 200, 1000, 1, 1, 4, RADVLM_EPI_ATOMIC_F32);     // ragged everything
  if (!quick) {
    fails += run_case_ex(29160, 1152, 4304, 0, 1, 1, RADVLM_EPI_BIAS_F32);
    fails += run_case_ex(4304, 1152, 29160, 1, 1, 4, RADVLM_EPI_ATOMIC_F32);
    fails += run_case_ex(1152, 4304, 29160, 1, 1, 4, RADVLM_EPI_ATOMIC_F32);
  }
  printf("%s (%d failing cases)\n", fails ? "GEMM TEST FAILED" : "GEMM TEST PASSED", fails);
  return fails ? 1 : 0;
}
