// Stand-alone bring-up test: QKV head-split GEMM epilogue + fused attention, through the C ABI.
// Reference = naive fp32 CUDA-core kernels on the same bf16 inputs.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

static uint32_t rng_state = 777u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

// one thread per (th, query row): naive softmax(q k^T * scale) v
__global__ void ref_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                              float* out, int th_count, int heads, int seq, int seq_pad, int hd,
                              int hd_pad, float scale) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int th = blockIdx.y;
  if (t >= seq) return;
  const __nv_bfloat16* qr = q + ((size_t)th * seq_pad + t) * hd_pad;
  float mx = -1e30f;
  for (int j = 0; j < seq; ++j) {
    const __nv_bfloat16* kr = k + ((size_t)th * seq_pad + j) * hd_pad;
    float s = 0;
    for (int d = 0; d < hd; ++d) s += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
    mx = fmaxf(mx, s * scale);
  }
  float l = 0;
  float acc[80];
  for (int d = 0; d < hd; ++d) acc[d] = 0;
  for (int j = 0; j < seq; ++j) {
    const __nv_bfloat16* kr = k + ((size_t)th * seq_pad + j) * hd_pad;
    float s = 0;
    for (int d = 0; d < hd; ++d) s += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
    float p = expf(s * scale - mx);
    l += p;
    for (int d = 0; d < hd; ++d)
      acc[d] += p * __bfloat162float(vt[((size_t)th * seq_pad + j) * hd_pad + d]);
  }
  int tile = th / heads, head = th % heads;
  for (int d = 0; d < hd; ++d)
    out[((size_t)tile * seq + t) * (heads * hd) + head * hd + d] = acc[d] / l;
}

__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, float* C,
                         int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int kk = 0; kk < K; ++kk)
    acc += __bfloat162float(A[(size_t)m * K + kk]) * __bfloat162float(W[(size_t)n * K + kk]);
  C[(size_t)m * N + n] = acc + bias[n];
}

static int test_attention(int tiles, int heads, float amp) {
  const int seq = 729, seq_pad = 768, hd = 72, hd_pad = 80;
  const int th = tiles * heads;
  const size_t nq = (size_t)th * seq_pad * hd_pad;
  std::vector<__nv_bfloat16> hq(nq, __float2bfloat16(0.f)), hk(nq, __float2bfloat16(0.f));
  std::vector<__nv_bfloat16> hv(nq, __float2bfloat16(0.f));
  for (int a = 0; a < th; ++a)
    for (int t = 0; t < seq; ++t)
      for (int d = 0; d < hd; ++d) {
        hq[((size_t)a * seq_pad + t) * hd_pad + d] = __float2bfloat16(frand() * amp);
        hk[((size_t)a * seq_pad + t) * hd_pad + d] = __float2bfloat16(frand() * amp);
        hv[((size_t)a * seq_pad + t) * hd_pad + d] = __float2bfloat16(frand() * 2.f);
      }
  __nv_bfloat16 *dq, *dk, *dout;
  __nv_bfloat16* dv;
  float* dref;
  const size_t nout = (size_t)tiles * seq * heads * hd;
  CK(cudaMalloc(&dq, nq * 2)); CK(cudaMalloc(&dk, nq * 2)); CK(cudaMalloc(&dv, nq * 2));
  CK(cudaMalloc(&dout, nout * 2)); CK(cudaMalloc(&dref, nout * 4));
  CK(cudaMemcpy(dq, hq.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), nq * 2, cudaMemcpyHostToDevice));
  for (int a = 0; a < th; ++a)
    for (int t = 0; t < seq; ++t) hv[((size_t)a * seq_pad + t) * hd_pad + hd] = __float2bfloat16(1.0f);  // ones row
  CK(cudaMemcpy(dv, hv.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, nout * 2));
  const float scale = 1.0f / sqrtf((float)hd);
  ref_attention<<<dim3((seq + 63) / 64, th), 64>>>(dq, dk, dv, dref, th, heads, seq, seq_pad, hd, hd_pad, scale);
  CK(cudaGetLastError());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int st = 0;
  const int reps = 5;
  for (int i = 0; i < reps + 1; ++i) {
    if (i == 1) cudaEventRecord(e0);
    st = radvlm_attention_fwd(dq, dk, dv, dout, tiles, heads, seq, seq_pad, hd, hd_pad, scale, nullptr);
    if (st) break;
  }
  cudaEventRecord(e1);
  if (st) { printf("attention API error %d: %s\n", st, radvlm_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("attention kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  std::vector<__nv_bfloat16> ho(nout);
  std::vector<float> hr(nout);
  CK(cudaMemcpy(ho.data(), dout, nout * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hr.data(), dref, nout * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0; size_t bad = 0, worst = 0;
  for (size_t i = 0; i < nout; ++i) {
    double g = __bfloat162float(ho[i]), r = hr[i];
    double err = fabs(g - r);
    if (!(err <= 4e-3 + 1e-2 * fabs(r))) ++bad;
    if (err > max_err || isnan(g)) { max_err = err; worst = i; }
    if (fabs(r) > max_ref) max_ref = fabs(r);
  }
  const double flops = 4.0 * th * (double)seq * seq * hd;
  printf("attention tiles=%d heads=%d amp=%.1f: max_err=%.3e max_ref=%.3e bad=%zu/%zu  %.3f ms  %.1f TFLOP/s(alg) %s\n",
         tiles, heads, amp, max_err, max_ref, bad, nout, ms, flops / (ms * 1e-3) / 1e12, bad ? "FAIL" : "ok");
  if (bad) {
    printf("  worst idx %zu (row %zu col %zu): got %f want %f\n", worst, worst / (heads * hd),
           worst % (heads * hd), __bfloat162float(ho[worst]), hr[worst]);
    printf("  got[0,0..7]:"); for (int j = 0; j < 8; ++j) printf(" %8.4f", __bfloat162float(ho[j]));
    printf("\n  ref[0,0..7]:"); for (int j = 0; j < 8; ++j) printf(" %8.4f", hr[j]);
    printf("\n");
  }
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dout); cudaFree(dref);
  return bad ? 1 : 0;
}

static int test_qkv_split(int tiles) {
  const int seq = 729, seq_pad = 768, heads = 16, hd = 72, hd_pad = 80, K = 1152, D = heads * hd;
  const int M = tiles * seq, N = 3 * D;
  std::vector<__nv_bfloat16> hA((size_t)M * K), hW((size_t)N * K);
  std::vector<float> hb(N);
  for (auto& v : hA) v = __float2bfloat16(frand());
  for (auto& v : hW) v = __float2bfloat16(frand() * 0.2f);
  for (auto& v : hb) v = frand();
  __nv_bfloat16 *dA, *dW, *dq, *dk; __nv_bfloat16* dv; float *db, *dref;
  const size_t nq = (size_t)tiles * heads * seq_pad * hd_pad;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2)); CK(cudaMalloc(&db, N * 4));
  CK(cudaMalloc(&dref, (size_t)M * N * 4));
  CK(cudaMalloc(&dq, nq * 2)); CK(cudaMalloc(&dk, nq * 2)); CK(cudaMalloc(&dv, nq * 2));
  CK(cudaMemset(dq, 0, nq * 2)); CK(cudaMemset(dk, 0, nq * 2));
  if (radvlm_attention_prepare_vt(dv, tiles, heads, seq, seq_pad, hd, hd_pad, nullptr)) { printf("prepare_vt: %s\n", radvlm_last_error()); return 1; }
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), N * 4, cudaMemcpyHostToDevice));
  ref_gemm<<<dim3((N + 127) / 128, M), 128>>>(dA, dW, db, dref, M, N, K);
  CK(cudaGetLastError());
  int st = radvlm_gemm_qkv_split(dA, K, dW, K, M, K, db, dq, dk, dv, seq, seq_pad, heads, hd, hd_pad, 0, nullptr);
  if (st) { printf("qkv API error %d: %s\n", st, radvlm_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("qkv kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<float> hr((size_t)M * N);
  std::vector<__nv_bfloat16> hq(nq), hk(nq);
  std::vector<__nv_bfloat16> hv(nq);
  CK(cudaMemcpy(hr.data(), dref, hr.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hq.data(), dq, nq * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hk.data(), dk, nq * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hv.data(), dv, nq * 2, cudaMemcpyDeviceToHost));
  size_t bad = 0, nonzero_pad = 0; double max_err = 0;
  for (int tile = 0; tile < tiles; ++tile)
    for (int h = 0; h < heads; ++h)
      for (int t = 0; t < seq_pad; ++t)
        for (int d = 0; d < hd_pad; ++d) {
          const size_t th = (size_t)tile * heads + h;
          float gq = __bfloat162float(hq[(th * seq_pad + t) * hd_pad + d]);
          float gk = __bfloat162float(hk[(th * seq_pad + t) * hd_pad + d]);
          float gv = __bfloat162float(hv[(th * seq_pad + t) * hd_pad + d]);
          if (t >= seq || d >= hd) {
            const float want_v = (d == hd && t < seq) ? 1.f : 0.f;  // ones row
            if (gq != 0.f || gk != 0.f || gv != want_v) ++nonzero_pad;
            continue;
          }
          const size_t row = (size_t)tile * seq + t;
          float rq = hr[row * N + 0 * D + h * hd + d];
          float rk = hr[row * N + 1 * D + h * hd + d];
          float rv = hr[row * N + 2 * D + h * hd + d];
          double eq = fabs(gq - rq), ek = fabs(gk - rk), ev = fabs(gv - rv);
          double em = fmax(eq, fmax(ek, ev));
          if (em > max_err) max_err = em;
          if (eq > 2e-3 + 8e-3 * fabs(rq) || ek > 2e-3 + 8e-3 * fabs(rk) || ev > 2e-3 + 8e-3 * fabs(rv)) ++bad;
        }
  printf("qkv_split tiles=%d: max_err=%.3e bad=%zu nonzero_pad=%zu %s\n", tiles, max_err, bad, nonzero_pad,
         (bad || nonzero_pad) ? "FAIL" : "ok");
  cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dref); cudaFree(dq); cudaFree(dk); cudaFree(dv);
  return (bad || nonzero_pad) ? 1 : 0;
}

// ---- attention backward: reference = one thread per (th, query row), fp32, atomics for dK / dV
__global__ void ref_attention_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                                  const __nv_bfloat16* dout, float* dq, float* dk, float* dv, int th_count, int heads,
                                  int seq, int seq_pad, int hd, int hd_pad, float scale) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int th = blockIdx.y;
  if (t >= seq || th >= th_count) return;
  const int tile = th / heads, head = th % heads;
  const __nv_bfloat16* qr = q + ((size_t)th * seq_pad + t) * hd_pad;
  const __nv_bfloat16* dor = dout + ((size_t)tile * seq + t) * (heads * hd) + head * hd;
  float mx = -1e30f;
  for (int j = 0; j < seq; ++j) {
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s += __bfloat162float(qr[d]) * __bfloat162float(k[((size_t)th * seq_pad + j) * hd_pad + d]);
    mx = fmaxf(mx, s * scale);
  }
  float l = 0.f, o[72], dqa[72];
  for (int d = 0; d < hd; ++d) { o[d] = 0.f; dqa[d] = 0.f; }
  for (int j = 0; j < seq; ++j) {
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s += __bfloat162float(qr[d]) * __bfloat162float(k[((size_t)th * seq_pad + j) * hd_pad + d]);
    float p = expf(s * scale - mx);
    l += p;
    for (int d = 0; d < hd; ++d) o[d] += p * __bfloat162float(vt[((size_t)th * seq_pad + j) * hd_pad + d]);
  }
  float delta = 0.f;
  for (int d = 0; d < hd; ++d) delta += (o[d] / l) * __bfloat162float(dor[d]);
  for (int j = 0; j < seq; ++j) {
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < hd; ++d) {
      s += __bfloat162float(qr[d]) * __bfloat162float(k[((size_t)th * seq_pad + j) * hd_pad + d]);
      dp += __bfloat162float(dor[d]) * __bfloat162float(vt[((size_t)th * seq_pad + j) * hd_pad + d]);
    }
    const float p = expf(s * scale - mx) / l;
    const float ds = p * (dp - delta) * scale;
    for (int d = 0; d < hd; ++d) {
      dqa[d] += ds * __bfloat162float(k[((size_t)th * seq_pad + j) * hd_pad + d]);
      atomicAdd(&dk[((size_t)th * seq + j) * hd + d], ds * __bfloat162float(qr[d]));
      atomicAdd(&dv[((size_t)th * seq + j) * hd + d], p * __bfloat162float(dor[d]));
    }
  }
  for (int d = 0; d < hd; ++d) dq[((size_t)th * seq + t) * hd + d] = dqa[d];
}

static int test_attention_bwd(int tiles, int heads, float amp) {
  const int seq = 729, seq_pad = 768, hd = 72, hd_pad = 80;
  const int th = tiles * heads, D = heads * hd;
  const size_t nq = (size_t)th * seq_pad * hd_pad, ntok = (size_t)tiles * seq;
  std::vector<__nv_bfloat16> hq(nq, __float2bfloat16(0.f)), hk(nq, __float2bfloat16(0.f)), hv(nq, __float2bfloat16(0.f));
  std::vector<__nv_bfloat16> hdo(ntok * D);
  for (int a = 0; a < th; ++a)
    for (int t = 0; t < seq; ++t)
      for (int d = 0; d < hd; ++d) {
        hq[((size_t)a * seq_pad + t) * hd_pad + d] = __float2bfloat16(frand() * amp);
        hk[((size_t)a * seq_pad + t) * hd_pad + d] = __float2bfloat16(frand() * amp);
        hv[((size_t)a * seq_pad + t) * hd_pad + d] = __float2bfloat16(frand() * 2.f);
      }
  for (auto& v : hdo) v = __float2bfloat16(frand());
  std::vector<__nv_bfloat16> hv1 = hv;  // forward copy with the ones row
  for (int a = 0; a < th; ++a)
    for (int t = 0; t < seq; ++t) hv1[((size_t)a * seq_pad + t) * hd_pad + hd] = __float2bfloat16(1.0f);
  __nv_bfloat16 *dq, *dk, *dv, *dv1, *ddo, *dout, *dqkv;
  float *dlse, *rdq, *rdk, *rdv;
  void* ws;
  const size_t wsb = radvlm_attention_bwd_workspace_bytes(tiles, heads, seq_pad);
  CK(cudaMalloc(&dq, nq * 2)); CK(cudaMalloc(&dk, nq * 2)); CK(cudaMalloc(&dv, nq * 2)); CK(cudaMalloc(&dv1, nq * 2));
  CK(cudaMalloc(&ddo, ntok * D * 2)); CK(cudaMalloc(&dout, ntok * D * 2)); CK(cudaMalloc(&dqkv, ntok * 3 * D * 2));
  CK(cudaMalloc(&dlse, (size_t)th * seq_pad * 4)); CK(cudaMalloc(&ws, wsb));
  CK(cudaMalloc(&rdq, (size_t)th * seq * hd * 4)); CK(cudaMalloc(&rdk, (size_t)th * seq * hd * 4));
  CK(cudaMalloc(&rdv, (size_t)th * seq * hd * 4));
  CK(cudaMemset(rdq, 0, (size_t)th * seq * hd * 4)); CK(cudaMemset(rdk, 0, (size_t)th * seq * hd * 4));
  CK(cudaMemset(rdv, 0, (size_t)th * seq * hd * 4));
  CK(cudaMemcpy(dq, hq.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv1, hv1.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ddo, hdo.data(), ntok * D * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dqkv, 0xFF, ntok * 3 * D * 2));
  const float scale = 1.0f / sqrtf((float)hd);
  int st = radvlm_attention_fwd_lse(dq, dk, dv1, dout, dlse, tiles, heads, seq, seq_pad, hd, hd_pad, scale, nullptr);
  if (st) { printf("fwd_lse API error %d: %s\n", st, radvlm_last_error()); return 1; }
  ref_attention_bwd<<<dim3((seq + 63) / 64, th), 64>>>(dq, dk, dv, ddo, rdq, rdk, rdv, th, heads, seq, seq_pad, hd, hd_pad, scale);
  CK(cudaGetLastError());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 3;
  for (int i = 0; i < reps + 1; ++i) {
    if (i == 1) cudaEventRecord(e0);
    st = radvlm_attention_bwd(dq, dk, dv, ddo, dout, dlse, dqkv, ws, wsb, tiles, heads, seq, seq_pad, hd, hd_pad, scale, nullptr);
    if (st) break;
  }
  cudaEventRecord(e1);
  if (st) { printf("attention_bwd API error %d: %s\n", st, radvlm_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("attention_bwd kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  std::vector<__nv_bfloat16> hg(ntok * 3 * D);
  std::vector<float> hr[3];
  CK(cudaMemcpy(hg.data(), dqkv, hg.size() * 2, cudaMemcpyDeviceToHost));
  float* rp[3] = {rdq, rdk, rdv};
  int fails = 0;
  const char* names[3] = {"dQ", "dK", "dV"};
  for (int w = 0; w < 3; ++w) {
    hr[w].resize((size_t)th * seq * hd);
    CK(cudaMemcpy(hr[w].data(), rp[w], hr[w].size() * 4, cudaMemcpyDeviceToHost));
    double max_ref = 0, max_err = 0, dot = 0, n1 = 0, n2 = 0;
    for (int a = 0; a < th; ++a)
      for (int t = 0; t < seq; ++t)
        for (int d = 0; d < hd; ++d) {
          const int tile = a / heads, head = a % heads;
          const double g = __bfloat162float(hg[((size_t)tile * seq + t) * 3 * D + w * D + head * hd + d]);
          const double r = hr[w][((size_t)a * seq + t) * hd + d];
          max_ref = fmax(max_ref, fabs(r));
          if (!(fabs(g - r) <= max_err)) max_err = fabs(g - r);
          dot += g * r; n1 += g * g; n2 += r * r;
        }
    const double cosv = dot / (sqrt(n1) * sqrt(n2) + 1e-30);
    const bool ok = cosv > 0.999 && max_err <= 3e-2 * max_ref;
    printf("attention_bwd tiles=%d heads=%d %s: cos=%.6f max_err=%.3e max_ref=%.3e %s\n", tiles, heads, names[w], cosv,
           max_err, max_ref, ok ? "ok" : "FAIL");
    fails += ok ? 0 : 1;
  }
  const double flops = 2.5 * 4.0 * th * (double)seq * seq * hd;
  printf("attention_bwd tiles=%d heads=%d: %.3f ms  %.1f TFLOP/s(alg)\n", tiles, heads, ms, flops / (ms * 1e-3) / 1e12);
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dv1); cudaFree(ddo); cudaFree(dout); cudaFree(dqkv); cudaFree(dlse);
  cudaFree(ws); cudaFree(rdq); cudaFree(rdk); cudaFree(rdv);
  return fails;
}

int main(int argc, char** argv) {
  int fails = 0;
  fails += test_qkv_split(2);
  fails += test_attention(1, 2, 2.0f);
  fails += test_attention(2, 16, 6.0f);   // larger logits: exercises the lazy rescale path
  fails += test_attention(10, 16, 2.0f);
  fails += test_attention(37, 16, 2.0f);  // 12 full waves of CTAs: steady-state throughput  // one 1024^2 image worth of tiles
  fails += test_attention_bwd(1, 2, 2.0f);
  fails += test_attention_bwd(2, 16, 4.0f);
  fails += test_attention_bwd(10, 16, 2.0f);
  printf("%s (%d failing)\n", fails ? "ATTENTION TEST FAILED" : "ATTENTION TEST PASSED", fails);
  return fails ? 1 : 0;
}
