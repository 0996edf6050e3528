// Tuning tool: prints the softmax / MMA timeline of the attention forward kernel.  Needs a library built with
// -DRV_ATTN_TIMELINE (tools/build_variant.sh timeline -DRV_ATTN_TIMELINE=1), run with LD_LIBRARY_PATH=build/var_timeline.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../include/radvlm_b200.h"

int main() {
  const int tiles = 37, heads = 16, seq = 729, seq_pad = 768, hd = 72, hd_pad = 80;
  const size_t nq = (size_t)tiles * heads * seq_pad * hd_pad;
  __nv_bfloat16 *q, *k, *v, *out;
  long long* tl;
  cudaMalloc(&q, nq * 2); cudaMalloc(&k, nq * 2); cudaMalloc(&v, nq * 2);
  cudaMemset(q, 0, nq * 2); cudaMemset(k, 0, nq * 2);
  radvlm_attention_prepare_vt(v, tiles, heads, seq, seq_pad, hd, hd_pad, nullptr);
  cudaMalloc(&out, (size_t)tiles * seq * heads * hd * 2);
  const int ctas = 296;
  cudaMalloc(&tl, ctas * 64 * 8);
  cudaMemset(tl, 0, ctas * 64 * 8);
  for (int rep = 0; rep < 2; ++rep)
    if (radvlm_attention_fwd_lse(q, k, v, out, (float*)tl, tiles, heads, seq, seq_pad, hd, hd_pad, 0.1f, nullptr)) {
      printf("error: %s\n", radvlm_last_error());
      return 1;
    }
  cudaDeviceSynchronize();
  std::vector<long long> h(ctas * 64);
  cudaMemcpy(h.data(), tl, h.size() * 8, cudaMemcpyDeviceToHost);
  // ping-pong kernel (default): [cta][group][8 * j + {S ready, S loaded, max, PV done, poly done, turn granted, turn passed, P published}], j < 4
  const bool pp = !(getenv("RADVLM_B200_ATTN") && getenv("RADVLM_B200_ATTN")[0] == '2');
  for (int c : {0, 100, 147}) {
    const long long* e = &h[c * 64];
    const long long t0 = e[0];
    if (pp) {
      printf("CTA %d, work item RV_ATTN_TIMELINE of the CTA (cycles from group A's first S ready)\n", c);
      for (int grp = 0; grp < 2; ++grp) {
        for (int j = 0; j < 3; ++j) {
          const long long* b = e + grp * 32 + 8 * j;
          printf("  %c j=%d  S_rdy %6lld  loaded %6lld  max %6lld  o_done %6lld  poly %6lld  turn %6lld  pass %6lld  P_pub %6lld\n",
                 'A' + grp, j, b[0] - t0, b[1] - t0, b[2] - t0, b[3] - t0, b[4] - t0, b[5] - t0, b[6] - t0, b[7] - t0);
        }
        const long long* b = e + grp * 32 + 24;
        printf("  %c item end (epilogue warps): O final %6lld  O read %6lld  stored %6lld | next item S_rdy j=0 %6lld  j=1 %6lld\n",
               'A' + grp, b[0] - t0, b[1] - t0, b[2] - t0, b[4] - t0, b[5] - t0);
      }
      continue;
    }
    printf("CTA %d, first work item (cycles from the first S ready)\n", c);
    for (int j = 0; j < 8; ++j)
      printf("  j=%d  S_rdy %6lld  loaded %6lld  xchg %6lld  o_free %6lld  exp_done %6lld  P_pub %6lld | S_iss %6lld  PV_iss %6lld\n", j,
             e[6 * j] - t0, e[6 * j + 1] - t0, e[6 * j + 2] - t0, e[6 * j + 3] - t0, e[6 * j + 4] - t0, e[6 * j + 5] - t0,
             e[48 + 2 * j] - t0, e[49 + 2 * j] - t0);
  }
  return 0;
}
