// C-ABI launchers for the tcgen05 GEMM (see gemm_sm100.cuh).
#include <cstdlib>
#include <cstring>

#include <vector>

#include "gemm2_sm100.cuh"
#include "gemm3_sm100.cuh"
#include "gemm4_sm100.cuh"
#include "gemm_sm100.cuh"
#include "host_util.h"

namespace rv {

template <int BN, int EPI>
static int launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
                            cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static thread_local bool configured = false;
  if (!configured) {
    RV_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, EPI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int num_tiles = ((args.M + kGemmBM - 1) / kGemmBM) * ((args.N + BN - 1) / BN);
  const int sms = device_sm_count();
  const int grid = num_tiles < sms ? num_tiles : sms;
  gemm_bf16_tn_kernel<BN, EPI><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, args);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

template <int BN, int EPI>
static int launch_gemm2_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
                             cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  static thread_local bool configured = false;
  if (!configured) {
    RV_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<BN, EPI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int num_tiles = ((args.M + 2 * kGemmBM - 1) / (2 * kGemmBM)) * ((args.N + BN - 1) / BN) *
                        (args.k_splits > 1 ? args.k_splits : 1);
  const int pairs = device_sm_count() / 2;
  const int grid = 2 * (num_tiles < pairs ? num_tiles : pairs);
  gemm_bf16_tn_2cta_kernel<BN, EPI><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, args);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// ---- scheduled variable-width pair kernel (gemm3_sm100.cuh) --------------------------------------------------------
// List scheduling, simulated once per (M, N, pairs): tiles in row-block-major order, each to the least loaded CTA pair
// (what a dynamic tile counter would do), a 128-wide tile charged 0.82 of a full one (measured; its MMAs last 64 cycles
// but its operands stream at the shared-memory / L2 limit).  Returns nullptr when the shape does not fit the parameter block.
static const GemmSched* gemm_sched_for(int M, int N, int pairs) {
  struct Item { int M, N, pairs; GemmSched s; };
  static thread_local std::vector<Item*> cache;
  for (const Item* it : cache)
    if (it->M == M && it->N == N && it->pairs == pairs) return &it->s;
  const int num_m = (M + 2 * kGemmBM - 1) / (2 * kGemmBM);
  const int num_n = (N + kSchedBN - 1) / kSchedBN;
  const int last_w = (N - (num_n - 1) * kSchedBN <= 128) ? 128 : kSchedBN;
  const long entries = static_cast<long>(num_m) * num_n;
  if (entries > kSchedMaxEntries || num_n > 32 || num_m > 2048 || pairs < 1) return nullptr;
  const int clusters = static_cast<int>(entries < pairs ? entries : pairs) < kSchedMaxClusters
                           ? static_cast<int>(entries < pairs ? entries : pairs) : kSchedMaxClusters;
  std::vector<std::vector<uint16_t>> lists(clusters);
  std::vector<long> load(clusters, 0);
  // cost of a 128-wide tile against 100 for a full one: 82 stand-alone (0.112 vs 0.137 ms, tools/test_gemm sched), about
  // 100 inside a K = 4304 launch (tools/gemm_timeline.cu: its MMAs take as long as a full tile's)
  static const int strip_cost = getenv("RADVLM_B200_STRIP_COST") ? atoi(getenv("RADVLM_B200_STRIP_COST")) : 82;
  for (int m = 0; m < num_m; ++m)
    for (int n = 0; n < num_n; ++n) {
      int best = 0;
      for (int c = 1; c < clusters; ++c)
        if (load[c] < load[best]) best = c;
      lists[best].push_back(static_cast<uint16_t>(m * 32 + n));
      load[best] += (n == num_n - 1 && last_w == 128) ? strip_cost : 100;
    }
  Item* it = new Item();
  it->M = M; it->N = N; it->pairs = pairs;
  int pos = 0;
  for (int c = 0; c <= kSchedMaxClusters; ++c) {
    it->s.off[c] = static_cast<uint16_t>(pos);
    if (c < clusters)
      for (uint16_t v : lists[c]) it->s.ent[pos++] = v;
  }
  if (cache.size() >= 64) { delete cache.front(); cache.erase(cache.begin()); }
  cache.push_back(it);
  return &it->s;
}

template <int EPI>
static int launch_gemm3_inst(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb64, const GemmArgs& args,
                             const GemmSched& sched, cudaStream_t stream) {
  static thread_local bool configured = false;
  if (!configured) {
    RV_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_2cta_sched_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Gemm3CfgT<EPI>::kSmemBytes));
    configured = true;
  }
  int clusters = 0;
  while (clusters < kSchedMaxClusters && sched.off[clusters + 1] > sched.off[clusters]) ++clusters;
  RV_CUDA(launch_kernel_pdl(gemm_bf16_tn_2cta_sched_kernel<EPI>, 2 * clusters, kGemmThreads, Gemm3CfgT<EPI>::kSmemBytes, stream,
                            pdl_enabled(), ta, tb, tb64, args, sched));
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

template <int EPI>
static int launch_gemm2_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
                           cudaStream_t stream) {
  switch (bn) {
    case 128: return launch_gemm2_inst<128, EPI>(ta, tb, args, stream);
    case 192: return launch_gemm2_inst<192, EPI>(ta, tb, args, stream);
    case 256: return launch_gemm2_inst<256, EPI>(ta, tb, args, stream);
  }
  set_error("block_n must be 128, 192 or 256 (got %d)", bn);
  return RADVLM_ERR_BAD_ARGUMENT;
}

template <int EPI>
static int launch_gemm_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
                          cudaStream_t stream) {
  switch (bn) {
    case 128: return launch_gemm_inst<128, EPI>(ta, tb, args, stream);
    case 192: return launch_gemm_inst<192, EPI>(ta, tb, args, stream);
    case 256: return launch_gemm_inst<256, EPI>(ta, tb, args, stream);
  }
  set_error("block_n must be 128, 192 or 256 (got %d)", bn);
  return RADVLM_ERR_BAD_ARGUMENT;
}

// Wave-quantisation aware tile-width choice: cost = waves * BN (MMA cycles per K step scale with BN).
#ifdef RV_GEMM_TIMELINE
static long long* g_gemm_timeline = nullptr;
#endif
static int g_gemm_mode = 0;  // 0 auto, 1 force single-CTA tiles, 2 force CTA-pair tiles
// scheduled variable-width tiles (gemm3_sm100.cuh) for the pair GEMMs whose A is K-major; RADVLM_B200_GEMM_SCHED=0
// keeps the round-robin 256 x BN kernel (A/B switch)
static bool gemm_sched_enabled() {
  static const bool on = !(getenv("RADVLM_B200_GEMM_SCHED") && atoi(getenv("RADVLM_B200_GEMM_SCHED")) == 0);
  return on;
}

int gemm_pick_block_n(int M, int N, int cta_group) {
  const int sms = (device_sm_count() > 0 ? device_sm_count() : 148) / cta_group;
  const int num_m = (M + cta_group * kGemmBM - 1) / (cta_group * kGemmBM);
  int best = 256;
  long best_cost = -1;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long tiles = static_cast<long>(num_m) * ((N + bn - 1) / bn);
    const long waves = (tiles + sms - 1) / sms;
    // per-tile cost ~ BN, penalised for narrow tiles (more L2->SMEM bytes per FLOP, the binding limit)
    const long cost = waves * bn * (bn == 256 ? 100 : (bn == 192 ? 108 : 135));
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

// MN-major B is fetched in 64-column boxes per CTA (BN/2 % 64 == 0): 256- or 128-wide tiles.  256 always: 128-wide
// tiles waste no padding at N = 1152 (9 x 128 against 5 x 256 = 1280 columns) but were measured 22 % slower on the
// data- and weight-gradient GEMMs of a 40-tile step (dgrad 19.0 -> 23.6 ms, wgrad 25.5 -> 31.3 ms): a 256 x 128 pair tile
// streams 96 B/clk of operands per SM, at the shared-memory limit.
int gemm_bmn_block_n(int N) {
  static const int force = getenv("RADVLM_B200_BMN_BN") ? atoi(getenv("RADVLM_B200_BMN_BN")) : 0;   // tuning only
  (void)N;
  return force == 128 ? 128 : 256;
}

// ---- two dependent GEMMs as one persistent kernel (gemm4_sm100.cuh): the projector and the ViT MLP block -------------
static const ChainSched* chain_sched_for(int rows, int n0_tiles, int n1_tiles, int last_w1, int k0_slabs, int k1_slabs,
                                         int pairs, int chunk) {
  struct Item { int key[8]; ChainSched s; };
  static thread_local std::vector<Item*> cache;
  const int key[8] = {rows, n0_tiles, n1_tiles, last_w1, k0_slabs, k1_slabs, pairs, chunk};
  for (const Item* it : cache)
    if (!memcmp(it->key, key, sizeof(key))) return &it->s;
  const int num_m = (rows + 2 * kGemmBM - 1) / (2 * kGemmBM);
  const long entries = static_cast<long>(num_m) * (n0_tiles + n1_tiles);
  if (entries > kChainMaxEntries || n0_tiles > 32 || n1_tiles > 32 || num_m > 1023 || pairs < 1) return nullptr;
  const int clusters = static_cast<int>(entries < pairs ? entries : pairs) < kSchedMaxClusters
                           ? static_cast<int>(entries < pairs ? entries : pairs) : kSchedMaxClusters;
  std::vector<std::vector<uint16_t>> lists(clusters);
  std::vector<long> load(clusters, 0);
  auto deal = [&](int phase, int m0, int m1) {   // tiles of row blocks [m0, m1) of one phase, row-block-major
    const int nt = phase ? n1_tiles : n0_tiles;
    for (int m = m0; m < m1 && m < num_m; ++m)
      for (int n = 0; n < nt; ++n) {
        int best = 0;
        for (int c = 1; c < clusters; ++c)
          if (load[c] < load[best]) best = c;
        lists[best].push_back(static_cast<uint16_t>(m * 64 + phase * 32 + n));
        const bool strip = phase == 1 && n == nt - 1 && last_w1 == 128;
        load[best] += (phase ? k1_slabs : k0_slabs) * (strip ? 82 : 100) + 400;   // MMA time + a tile's fixed cost
      }
  };
  const int chunks = (num_m + chunk - 1) / chunk;
  for (int c = 0; c <= chunks; ++c) {   // phase 0 runs one chunk ahead of phase 1
    if (c < chunks) deal(0, c * chunk, (c + 1) * chunk);
    if (c > 0) deal(1, (c - 1) * chunk, c * chunk);
  }
  Item* it = new Item();
  memcpy(it->key, key, sizeof(key));
  int pos = 0;
  for (int c = 0; c <= kSchedMaxClusters; ++c) {
    it->s.off[c] = static_cast<uint16_t>(pos);
    if (c < clusters)
      for (uint16_t v : lists[c]) it->s.ent[pos++] = v;
  }
  if (cache.size() >= 16) { delete cache.front(); cache.erase(cache.begin()); }
  cache.push_back(it);
  return &it->s;
}

template <int EPI1, int EPI2>
static int launch_chain_inst(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& th, const CUtensorMap& tw2,
                             const CUtensorMap& tw2h, const ChainArgs& ca, const ChainSched& sched, cudaStream_t stream) {
  static thread_local bool configured = false;
  if (!configured) {
    RV_CUDA(cudaFuncSetAttribute(gemm_chain_kernel<EPI1, EPI2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Gemm3Cfg::kSmemBytes));
    configured = true;
  }
  int clusters = 0;
  while (clusters < kSchedMaxClusters && sched.off[clusters + 1] > sched.off[clusters]) ++clusters;
  gemm_chain_kernel<EPI1, EPI2><<<2 * clusters, kGemmThreads, Gemm3Cfg::kSmemBytes, stream>>>(tx, tw1, th, tw2, tw2h, ca, sched);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// H = epi0(X W1^T ...) [g0.M, g0.N] bf16 (= g0.out, row pitch g0.ldo), then epi1(H W2^T ...) described by g1 (g1.K == g0.N).
// ready: g0.M / 256 (rounded up) unsigned ints of scratch.  Returns RADVLM_ERR_UNSUPPORTED_SHAPE when the chained kernel
// does not cover the case (the caller then launches the two GEMMs separately).
int gemm_chain_dispatch(const void* X, const void* W1, const void* W2, const GemmArgs& g0, int epi0, const GemmArgs& g1,
                        int epi1, void* ready, cudaStream_t stream) {
  int st = require_sm100();
  if (st != RADVLM_OK) return st;
  static const int chunk_env = getenv("RADVLM_B200_CHAIN_CHUNK") ? atoi(getenv("RADVLM_B200_CHAIN_CHUNK")) : 0;
  const int pairs = device_sm_count() / 2;
  const int rows = g0.M;
  if (ready == nullptr || g1.M != rows || g1.K != g0.N || (g0.K & 7) || (g0.N & 7) || (g1.N & 7) || (g0.ldo & 7) ||
      rows < 4 * kGemmBM || g_gemm_mode == 1)
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  const int n0_tiles = (g0.N + kSchedBN - 1) / kSchedBN, n1_tiles = (g1.N + kSchedBN - 1) / kSchedBN;
  const int last_w1 = (g1.N - (n1_tiles - 1) * kSchedBN <= 128) ? 128 : kSchedBN;
  const int k0 = (g0.K + kGemmBK - 1) / kGemmBK, k1 = (g1.K + kGemmBK - 1) / kGemmBK;
  const ChainSched* sched = chain_sched_for(rows, n0_tiles, n1_tiles, last_w1, k0, k1, pairs, chunk_env > 0 ? chunk_env : 16);
  if (sched == nullptr) return RADVLM_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tx, tw1, th, tw2, tw2h;
  if ((st = make_tmap_bf16_2d(&tx, X, g0.K, rows, static_cast<uint64_t>(g0.K) * 2, kGemmBK, kGemmBM, CU_TENSOR_MAP_SWIZZLE_128B))) return st;
  if ((st = make_tmap_bf16_2d(&tw1, W1, g0.K, g0.N, static_cast<uint64_t>(g0.K) * 2, kGemmBK, kSchedBN / 2, CU_TENSOR_MAP_SWIZZLE_128B))) return st;
  if ((st = make_tmap_bf16_2d(&th, g0.out, g0.N, rows, static_cast<uint64_t>(g0.ldo) * 2, kGemmBK, kGemmBM, CU_TENSOR_MAP_SWIZZLE_128B))) return st;
  if ((st = make_tmap_bf16_2d(&tw2, W2, g1.K, g1.N, static_cast<uint64_t>(g1.K) * 2, kGemmBK, kSchedBN / 2, CU_TENSOR_MAP_SWIZZLE_128B))) return st;
  if ((st = make_tmap_bf16_2d(&tw2h, W2, g1.K, g1.N, static_cast<uint64_t>(g1.K) * 2, kGemmBK, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return st;
  ChainArgs ca{};
  ca.g[0] = g0;
  ca.g[1] = g1;
  ca.ready = static_cast<unsigned int*>(ready);
  ca.ready_target = static_cast<unsigned int>(n0_tiles) * 2u * kGemmEpiWarps;
  const int num_m = (rows + 2 * kGemmBM - 1) / (2 * kGemmBM);
  RV_CUDA(cudaMemsetAsync(ready, 0, static_cast<size_t>(num_m) * sizeof(unsigned int), stream));
#define RV_CHAIN(E0, E1) if (epi0 == E0 && epi1 == E1) return launch_chain_inst<E0, E1>(tx, tw1, th, tw2, tw2h, ca, *sched, stream)
  RV_CHAIN(EPI_GELU_ERF_BF16, EPI_BIAS_BF16);
  RV_CHAIN(EPI_GELU_ERF_BF16, EPI_BIAS_F16);
  RV_CHAIN(EPI_GELU_ERF_BF16, EPI_BIAS_F32);
  RV_CHAIN(EPI_GELU_TANH_BF16, EPI_RESID_F32);
#undef RV_CHAIN
  return RADVLM_ERR_UNSUPPORTED_SHAPE;
}

int projector_chain_dispatch(const void* X, const void* W1, const float* b1, void* H, const void* W2, const float* b2,
                             void* out, int out_dtype, int rows, int in_dim, int hidden, void* ready,
                             cudaStream_t stream) {
  static const bool off = getenv("RADVLM_B200_PROJ") && !strcmp(getenv("RADVLM_B200_PROJ"), "split");
  if (off) return RADVLM_ERR_UNSUPPORTED_SHAPE;
  GemmArgs g0{}, g1{};
  g0.M = rows; g0.N = hidden; g0.K = in_dim; g0.bias = b1; g0.out = H; g0.ldo = hidden;
  g1.M = rows; g1.N = hidden; g1.K = hidden; g1.bias = b2; g1.out = out; g1.ldo = hidden;
  const int epi1 = out_dtype == RADVLM_DT_BF16 ? EPI_BIAS_BF16 : (out_dtype == RADVLM_DT_F16 ? EPI_BIAS_F16 :
                   (out_dtype == RADVLM_DT_F32 ? EPI_BIAS_F32 : -1));
  RV_CHECK_ARG(epi1 >= 0, "projector: out_dtype must be bf16, f16 or f32");
  return gemm_chain_dispatch(X, W1, W2, g0, EPI_GELU_ERF_BF16, g1, epi1, ready, stream);
}

int gemm_ln_part_slots(int M, int N) {
  if (!gemm_sched_enabled() || g_gemm_mode == 1 || M < 4 * kGemmBM || (N & 7) != 0) return 0;
  if (gemm_sched_for(M, N, device_sm_count() / 2) == nullptr) return 0;
  return 2 * ((N + kSchedBN - 1) / kSchedBN);
}

int gemm_dispatch(const void* A, int64_t lda, const void* W, int64_t ldw, const GemmArgs& args_in,
                  int epilogue, int block_n, cudaStream_t stream) {
  int st = require_sm100();
  if (st != RADVLM_OK) return st;
  GemmArgs args = args_in;
#ifdef RV_GEMM_TIMELINE
  args.timeline = g_gemm_timeline;
#endif
  RV_CHECK_ARG(A != nullptr && W != nullptr, "gemm: null operand");
  RV_CHECK_ARG(args.M > 0 && args.N > 0 && args.K > 0, "gemm: bad shape M=%d N=%d K=%d", args.M,
               args.N, args.K);
  RV_CHECK_ARG(lda >= (args.a_mn ? args.M : args.K) && ldw >= (args.b_mn ? args.N : args.K),
               "gemm: row pitch smaller than the stored row length");
  RV_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0, "gemm: operand pitches must be multiples of 8 elements");
  RV_CHECK_ARG(epilogue != EPI_DELTA_BF16 || (args.out2 != nullptr && args.aux16 != nullptr && args.out != nullptr &&
                                               (args.N & 7) == 0 && (args.ldo & 7) == 0),
               "gemm: the bf16 branch epilogue needs out, out2, aux16 and N, ldo multiples of 8");
  RV_CHECK_ARG((epilogue != EPI_GELU_TANH_DUAL_BF16 && epilogue != EPI_MUL_BF16) || args.out2 != nullptr,
               "gemm: the dual-output / GELU-backward epilogues need out2");
  const bool general = args.a_mn || args.b_mn || args.k_splits > 1;
  RV_CHECK_ARG(args.k_splits <= 1 || epilogue == EPI_ATOMIC_F32, "gemm: split-K needs the atomic fp32 epilogue");
  // CTA pairs (256 x BN tiles) whenever there is enough work to fill the 74 pairs; the MN-major / split-K paths
  // exist in the pair kernel only
  const bool pair = general || block_n < 0 || g_gemm_mode == 2 || (g_gemm_mode == 0 && args.M >= 4 * kGemmBM);
  int bn = block_n > 0 ? block_n : gemm_pick_block_n(args.M, args.N, pair ? 2 : 1);
  if (args.b_mn && block_n <= 0) bn = gemm_bmn_block_n(args.N);
  if (args.k_splits > 1) {  // no empty split: (splits - 1) * ceil(slabs / splits) < slabs
    const int slabs = (args.K + kGemmBK - 1) / kGemmBK;
    int sp = args.k_splits < slabs ? args.k_splits : slabs;
    while (sp > 1 && (sp - 1) * ((slabs + sp - 1) / sp) >= slabs) --sp;
    args.k_splits = sp;
  }
  // block_n < 0 forces the scheduled kernel (tests), block_n > 0 the round-robin one
  const GemmSched* sched = nullptr;
  if (pair && !args.a_mn && args.k_splits <= 1 && epilogue != EPI_ATOMIC_F32 &&
      (block_n < 0 || (block_n == 0 && gemm_sched_enabled())))
    sched = gemm_sched_for(args.M, args.N, device_sm_count() / 2);
  RV_CHECK_ARG(block_n >= 0 || sched != nullptr, "gemm: the scheduled kernel does not cover this shape / layout");
  CUtensorMap ta, tb;
  if (!args.a_mn)
    st = make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(args.K), static_cast<uint64_t>(args.M),
                           static_cast<uint64_t>(lda) * 2, kGemmBK, kGemmBM, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    st = make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(args.M), static_cast<uint64_t>(args.K),
                           static_cast<uint64_t>(lda) * 2, 64, kGemmBK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != RADVLM_OK) return st;
  if (!args.b_mn)
    st = make_tmap_bf16_2d(&tb, W, static_cast<uint64_t>(args.K), static_cast<uint64_t>(args.N),
                           static_cast<uint64_t>(ldw) * 2, kGemmBK,
                           static_cast<uint32_t>(sched ? 128 : (pair ? bn / 2 : bn)), CU_TENSOR_MAP_SWIZZLE_128B);
  else
    st = make_tmap_bf16_2d(&tb, W, static_cast<uint64_t>(args.N), static_cast<uint64_t>(args.K),
                           static_cast<uint64_t>(ldw) * 2, 64, kGemmBK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != RADVLM_OK) return st;
  if (sched != nullptr) {
    CUtensorMap tb64 = tb;   // 128-wide tiles: each CTA of the pair fetches 64 rows of W
    if (!args.b_mn) {
      st = make_tmap_bf16_2d(&tb64, W, static_cast<uint64_t>(args.K), static_cast<uint64_t>(args.N),
                             static_cast<uint64_t>(ldw) * 2, kGemmBK, 64, CU_TENSOR_MAP_SWIZZLE_128B);
      if (st != RADVLM_OK) return st;
    }
    switch (epilogue) {
      case EPI_BIAS_BF16: return launch_gemm3_inst<EPI_BIAS_BF16>(ta, tb, tb64, args, *sched, stream);
      case EPI_GELU_TANH_BF16: return launch_gemm3_inst<EPI_GELU_TANH_BF16>(ta, tb, tb64, args, *sched, stream);
      case EPI_GELU_ERF_BF16: return launch_gemm3_inst<EPI_GELU_ERF_BF16>(ta, tb, tb64, args, *sched, stream);
      case EPI_RESID_F32: return launch_gemm3_inst<EPI_RESID_F32>(ta, tb, tb64, args, *sched, stream);
      case EPI_POS_F32: return launch_gemm3_inst<EPI_POS_F32>(ta, tb, tb64, args, *sched, stream);
      case EPI_QKV_SPLIT: return launch_gemm3_inst<EPI_QKV_SPLIT>(ta, tb, tb64, args, *sched, stream);
      case EPI_BIAS_F32: return launch_gemm3_inst<EPI_BIAS_F32>(ta, tb, tb64, args, *sched, stream);
      case EPI_GELU_TANH_DUAL_BF16: return launch_gemm3_inst<EPI_GELU_TANH_DUAL_BF16>(ta, tb, tb64, args, *sched, stream);
      case EPI_BIAS_F16: return launch_gemm3_inst<EPI_BIAS_F16>(ta, tb, tb64, args, *sched, stream);
      case EPI_MUL_BF16: return launch_gemm3_inst<EPI_MUL_BF16>(ta, tb, tb64, args, *sched, stream);
      case EPI_DELTA_BF16: return launch_gemm3_inst<EPI_DELTA_BF16>(ta, tb, tb64, args, *sched, stream);
    }
    set_error("gemm: unknown epilogue %d", epilogue);
    return RADVLM_ERR_BAD_ARGUMENT;
  }
  if (pair) {
    switch (epilogue) {
      case EPI_BIAS_BF16: return launch_gemm2_bn<EPI_BIAS_BF16>(bn, ta, tb, args, stream);
      case EPI_GELU_TANH_BF16: return launch_gemm2_bn<EPI_GELU_TANH_BF16>(bn, ta, tb, args, stream);
      case EPI_GELU_ERF_BF16: return launch_gemm2_bn<EPI_GELU_ERF_BF16>(bn, ta, tb, args, stream);
      case EPI_RESID_F32: return launch_gemm2_bn<EPI_RESID_F32>(bn, ta, tb, args, stream);
      case EPI_POS_F32: return launch_gemm2_bn<EPI_POS_F32>(bn, ta, tb, args, stream);
      case EPI_QKV_SPLIT: return launch_gemm2_bn<EPI_QKV_SPLIT>(bn, ta, tb, args, stream);
      case EPI_BIAS_F32: return launch_gemm2_bn<EPI_BIAS_F32>(bn, ta, tb, args, stream);
      case EPI_ATOMIC_F32: return launch_gemm2_bn<EPI_ATOMIC_F32>(bn, ta, tb, args, stream);
      case EPI_GELU_TANH_DUAL_BF16: return launch_gemm2_bn<EPI_GELU_TANH_DUAL_BF16>(bn, ta, tb, args, stream);
      case EPI_BIAS_F16: return launch_gemm2_bn<EPI_BIAS_F16>(bn, ta, tb, args, stream);
      case EPI_MUL_BF16: return launch_gemm2_bn<EPI_MUL_BF16>(bn, ta, tb, args, stream);
      case EPI_DELTA_BF16: return launch_gemm2_bn<EPI_DELTA_BF16>(bn, ta, tb, args, stream);
    }
    set_error("gemm: unknown epilogue %d", epilogue);
    return RADVLM_ERR_BAD_ARGUMENT;
  }
  switch (epilogue) {
    case EPI_BIAS_BF16: return launch_gemm_bn<EPI_BIAS_BF16>(bn, ta, tb, args, stream);
    case EPI_GELU_TANH_BF16: return launch_gemm_bn<EPI_GELU_TANH_BF16>(bn, ta, tb, args, stream);
    case EPI_GELU_ERF_BF16: return launch_gemm_bn<EPI_GELU_ERF_BF16>(bn, ta, tb, args, stream);
    case EPI_RESID_F32: return launch_gemm_bn<EPI_RESID_F32>(bn, ta, tb, args, stream);
    case EPI_POS_F32: return launch_gemm_bn<EPI_POS_F32>(bn, ta, tb, args, stream);
    case EPI_QKV_SPLIT: return launch_gemm_bn<EPI_QKV_SPLIT>(bn, ta, tb, args, stream);
    case EPI_BIAS_F32: return launch_gemm_bn<EPI_BIAS_F32>(bn, ta, tb, args, stream);
    case EPI_ATOMIC_F32: return launch_gemm_bn<EPI_ATOMIC_F32>(bn, ta, tb, args, stream);
    case EPI_GELU_TANH_DUAL_BF16: return launch_gemm_bn<EPI_GELU_TANH_DUAL_BF16>(bn, ta, tb, args, stream);
    case EPI_BIAS_F16: return launch_gemm_bn<EPI_BIAS_F16>(bn, ta, tb, args, stream);
    case EPI_MUL_BF16: return launch_gemm_bn<EPI_MUL_BF16>(bn, ta, tb, args, stream);
    case EPI_DELTA_BF16: return launch_gemm_bn<EPI_DELTA_BF16>(bn, ta, tb, args, stream);
  }
  set_error("gemm: unknown epilogue %d", epilogue);
  return RADVLM_ERR_BAD_ARGUMENT;
}

}  // namespace rv

// Host-only view of the tile schedule of the scheduled GEMM kernel (no CUDA call: usable in CPU tests).  pairs: CTA
// pairs of the device (74 on B200).  tiles_per_pair[pairs] (may be NULL) receives the number of tiles of each pair;
// *n_tiles the total, *max_load / *min_load the largest / smallest summed tile cost (a 256-wide tile = 100, a 128-wide
// one = 82).  Every (row block, column tile) appears exactly once, in row-block-major order inside every pair's list
// (checked here: returns RADVLM_ERR_BAD_ARGUMENT otherwise).
extern "C" int radvlm_gemm_schedule_stats(int M, int N, int pairs, int* tiles_per_pair, int* n_tiles, int* max_load,
                                          int* min_load) {
  using namespace rv;
  RV_CHECK_ARG(M > 0 && N > 0 && pairs > 0 && pairs <= kSchedMaxClusters, "schedule_stats: bad arguments");
  const GemmSched* s = gemm_sched_for(M, N, pairs);
  if (s == nullptr) {
    set_error("schedule_stats: shape [%d, %d] is not covered by the scheduled kernel", M, N);
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  const int num_m = (M + 2 * kGemmBM - 1) / (2 * kGemmBM);
  const int num_n = (N + kSchedBN - 1) / kSchedBN;
  const bool strip = (N - (num_n - 1) * kSchedBN) <= 128;
  std::vector<char> seen(static_cast<size_t>(num_m) * num_n, 0);
  int total = 0, lo = 1 << 30, hi = 0;
  for (int c = 0; c < pairs; ++c) {
    int load = 0, prev = -1;
    for (int e = s->off[c]; e < s->off[c + 1]; ++e) {
      const int v = s->ent[e], m = v >> 5, n = v & 31;
      RV_CHECK_ARG(m < num_m && n < num_n && !seen[static_cast<size_t>(m) * num_n + n] && v > prev,
                   "schedule_stats: corrupt schedule (pair %d entry %d)", c, e);
      seen[static_cast<size_t>(m) * num_n + n] = 1;
      prev = v;
      load += (strip && n == num_n - 1) ? 82 : 100;
      ++total;
    }
    if (tiles_per_pair) tiles_per_pair[c] = s->off[c + 1] - s->off[c];
    if (s->off[c + 1] > s->off[c]) { lo = load < lo ? load : lo; hi = load > hi ? load : hi; }
  }
  RV_CHECK_ARG(total == num_m * num_n, "schedule_stats: %d tiles scheduled, %d expected", total, num_m * num_n);
  if (n_tiles) *n_tiles = total;
  if (max_load) *max_load = hi;
  if (min_load) *min_load = lo;
  return RADVLM_OK;
}

extern "C" int radvlm_gemm_set_mode(int mode) {
  if (mode < 0 || mode > 2) {
    rv::set_error("gemm mode must be 0 (auto), 1 (single-CTA tiles) or 2 (CTA-pair tiles)");
    return RADVLM_ERR_BAD_ARGUMENT;
  }
  rv::g_gemm_mode = mode;
  return RADVLM_OK;
}

extern "C" int radvlm_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N,
                                int K, const float* bias, int epilogue, void* out, int64_t ldo,
                                const float* aux, int aux_period, int block_n, void* stream) {
  using namespace rv;
  RV_CHECK_ARG(epilogue != EPI_QKV_SPLIT, "use radvlm_gemm_qkv_split for the QKV epilogue");
  RV_CHECK_ARG(epilogue != EPI_GELU_TANH_DUAL_BF16 && epilogue != EPI_MUL_BF16,
               "the dual-output GELU / GELU-backward epilogues are internal to the training path");
  RV_CHECK_ARG(out != nullptr && ldo >= N, "gemm: bad output (ldo=%lld N=%d)", (long long)ldo, N);
  RV_CHECK_ARG((ldo % 8) == 0, "gemm: ldo must be a multiple of 8 elements");
  RV_CHECK_ARG(epilogue != EPI_RESID_F32 || aux != nullptr, "gemm: residual epilogue needs aux");
  RV_CHECK_ARG(epilogue != EPI_POS_F32 || (aux != nullptr && aux_period > 0 && (N % 4) == 0),
               "gemm: position epilogue needs aux/aux_period and N % 4 == 0");
  GemmArgs a{};
  a.M = M; a.N = N; a.K = K;
  a.bias = bias;
  a.out = out;
  a.ldo = static_cast<int>(ldo);
  a.aux = aux;
  a.aux_period = aux_period;
  return gemm_dispatch(A, lda, W, ldw, a, epilogue, block_n, static_cast<cudaStream_t>(stream));
}

// tuning builds (-DRV_GEMM_TIMELINE, tools/gemm_timeline.cu): where the instrumented kernel writes its clock64 stamps
extern "C" void radvlm_gemm_set_timeline(long long* p) {
#ifdef RV_GEMM_TIMELINE
  rv::g_gemm_timeline = p;
#else
  (void)p;
#endif
}

extern "C" int radvlm_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                                   const float* bias, const void* ln_stats, const float* ln_s, int epilogue, void* out,
                                   int64_t ldo, void* stream) {
  using namespace rv;
  RV_CHECK_ARG(epilogue == EPI_BIAS_BF16 || epilogue == EPI_GELU_TANH_BF16 || epilogue == EPI_GELU_ERF_BF16,
               "gemm_ln: the LayerNorm fold exists for the bf16 epilogues (and the QKV head split inside the tower)");
  RV_CHECK_ARG(out != nullptr && ldo >= N && (ldo % 8) == 0 && (N % 8) == 0, "gemm_ln: bad output (ldo=%lld N=%d)",
               (long long)ldo, N);
  RV_CHECK_ARG(ln_stats != nullptr && ln_s != nullptr, "gemm_ln: null statistics / row sums");
  GemmArgs a{};
  a.M = M; a.N = N; a.K = K;
  a.bias = bias;
  a.out = out;
  a.ldo = static_cast<int>(ldo);
  a.ln_stats = static_cast<const float2*>(ln_stats);
  a.ln_s = ln_s;
  return gemm_dispatch(A, lda, W, ldw, a, epilogue, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_gemm_bf16_ex(const void* A, int64_t lda, int a_layout, const void* W, int64_t ldw, int b_layout,
                                   int M, int N, int K, const float* bias, int epilogue, void* out, int64_t ldo,
                                   const float* aux, int aux_period, int k_splits, void* stream) {
  using namespace rv;
  RV_CHECK_ARG(epilogue != EPI_QKV_SPLIT, "use radvlm_gemm_qkv_split for the QKV epilogue");
  RV_CHECK_ARG(epilogue != EPI_GELU_TANH_DUAL_BF16 && epilogue != EPI_MUL_BF16,
               "the dual-output GELU / GELU-backward epilogues are internal to the training path");
  RV_CHECK_ARG(out != nullptr && ldo >= N && (ldo % 8) == 0, "gemm: bad output (ldo=%lld N=%d)", (long long)ldo, N);
  RV_CHECK_ARG((a_layout | b_layout) >= 0 && a_layout <= 1 && b_layout <= 1, "gemm: layouts are 0 or 1");
  RV_CHECK_ARG(epilogue != EPI_RESID_F32 || aux != nullptr, "gemm: residual epilogue needs aux");
  RV_CHECK_ARG(epilogue != EPI_POS_F32 || (aux != nullptr && aux_period > 0 && (N % 4) == 0),
               "gemm: position epilogue needs aux/aux_period and N % 4 == 0");
  RV_CHECK_ARG(epilogue != EPI_ATOMIC_F32 || ((N % 4) == 0 && bias == nullptr), "gemm: atomic epilogue: N % 4 == 0, no bias");
  GemmArgs a{};
  a.M = M; a.N = N; a.K = K;
  a.bias = bias;
  a.out = out;
  a.ldo = static_cast<int>(ldo);
  a.aux = aux;
  a.aux_period = aux_period;
  a.a_mn = a_layout; a.b_mn = b_layout; a.k_splits = k_splits;
  return gemm_dispatch(A, lda, W, ldw, a, epilogue, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_gemm_qkv_split(const void* A, int64_t lda, const void* W, int64_t ldw, int M,
                                     int K, const float* bias, void* q, void* k, void* vt, int seq,
                                     int seq_pad, int heads, int hd, int hd_pad, int block_n,
                                     void* stream) {
  using namespace rv;
  RV_CHECK_ARG(q && k && vt, "qkv: null output");
  RV_CHECK_ARG(seq > 0 && seq_pad >= seq && hd_pad >= hd && (hd % 8) == 0 && (hd_pad % 8) == 0 &&
                   (seq_pad % 8) == 0 && (M % seq) == 0,
               "qkv: bad geometry seq=%d seq_pad=%d hd=%d hd_pad=%d M=%d", seq, seq_pad, hd, hd_pad, M);
  GemmArgs a{};
  a.M = M; a.N = 3 * heads * hd; a.K = K;
  a.bias = bias;
  a.q = static_cast<__nv_bfloat16*>(q);
  a.k = static_cast<__nv_bfloat16*>(k);
  a.vt = static_cast<__nv_bfloat16*>(vt);
  a.seq = seq; a.seq_pad = seq_pad; a.heads = heads; a.hd = hd; a.hd_pad = hd_pad;
  return gemm_dispatch(A, lda, W, ldw, a, EPI_QKV_SPLIT, block_n, static_cast<cudaStream_t>(stream));
}
