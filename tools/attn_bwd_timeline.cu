// Tuning tool: prints the per-step timeline of the attention backward kernel (MMA issuer, softmax-gradient warps, dQ
// warps) for the six key-block CTAs of one (tile, head).  Needs a library built with -DRV_ABWD_TIMELINE=<tile>
// (tools/build_variant.sh abtl -DRV_ABWD_TIMELINE=20), run with LD_LIBRARY_PATH=build/var_abtl.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../include/radvlm_b200.h"

extern "C" void radvlm_debug_abwd_timeline(void* buffer) __attribute__((weak));  // only in -DRV_ABWD_TIMELINE builds

int main() {
  const int tiles = 40, heads = 16, seq = 729, seq_pad = 768, hd = 72, hd_pad = 80;
  const size_t nq = (size_t)tiles * heads * seq_pad * hd_pad;
  const size_t nd = (size_t)tiles * seq * heads * hd;
  __nv_bfloat16 *q, *k, *v, *dout, *out, *dqkv;
  float* lse;
  void* ws;
  long long* tl;
  cudaMalloc(&q, nq * 2); cudaMalloc(&k, nq * 2); cudaMalloc(&v, nq * 2);
  cudaMemset(q, 0, nq * 2); cudaMemset(k, 0, nq * 2); cudaMemset(v, 0, nq * 2);
  cudaMalloc(&dout, nd * 2); cudaMalloc(&out, nd * 2); cudaMalloc(&dqkv, nd * 3 * 2);
  cudaMemset(dout, 0, nd * 2); cudaMemset(out, 0, nd * 2);
  cudaMalloc(&lse, (size_t)tiles * heads * seq_pad * 4);
  cudaMemset(lse, 0, (size_t)tiles * heads * seq_pad * 4);
  const size_t wsb = radvlm_attention_bwd_workspace_bytes(tiles, heads, seq_pad);
  cudaMalloc(&ws, wsb);
  cudaMalloc(&tl, 6 * 128 * 8);
  cudaMemset(tl, 0, 6 * 128 * 8);
  if (!radvlm_debug_abwd_timeline) {
    printf("this libradvlm_b200.so was not built with -DRV_ABWD_TIMELINE (see the header of this file)\n");
    return 2;
  }
  radvlm_debug_abwd_timeline(tl);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (radvlm_attention_bwd(q, k, v, dout, out, lse, dqkv, ws, wsb, tiles, heads, seq, seq_pad, hd, hd_pad, 0.1f, nullptr)) {
      printf("error: %s\n", radvlm_last_error());
      return 1;
    }
    cudaEventRecord(e1);
  }
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("attention backward, %d tiles: %.3f ms per launch (memset + delta + kernel + dq store)\n", tiles, ms);
  std::vector<long long> h(6 * 128);
  cudaMemcpy(h.data(), tl, h.size() * 8, cudaMemcpyDeviceToHost);
  for (int j = 0; j < 6; ++j) {
    const long long* e = &h[j * 128];
    const long long t0 = e[96];
    printf("key-block CTA %d: set up 0, K/V landed %lld, dQ warps done %lld, dK/dV stored %lld\n", j, e[97] - t0, e[98] - t0,
           e[99] - t0);
    for (int t = 0; t < 6; ++t) {
      const long long* b = e + t * 16;
      printf("  t=%d mma: sdpfree %6lld S'dP' issued %6lld pds %6lld dQ/dV/dK issued %6lld | softmax: S rdy %6lld S,dP in regs %6lld "
             "mma2 ok %6lld P,dS pub %6lld | dq: rdy %6lld read %6lld reduced %6lld\n",
             t, b[0] ? b[0] - t0 : 0, b[1] ? b[1] - t0 : 0, b[2] - t0, b[3] - t0, b[4] - t0, b[5] - t0, b[6] - t0, b[7] - t0,
             b[8] - t0, b[9] - t0, b[10] - t0);
    }
  }
  return 0;
}
