// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   * A (activations, row-major [M,K]) and W (nn.Linear weight, row-major [N,K]) are both K-major,
//     fetched by TMA (128B swizzle, 64-element K slabs) into a multi-stage shared-memory ring.
//   * One elected thread issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1); fp32 accumulators
//     live in TMEM, double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
//   * Four epilogue warps read TMEM with tcgen05.ld (one output row per thread) and apply the fused
//     epilogue: bias, GELU (tanh / erf), fp32 residual add, position-embedding add, or the QKV
//     head-split scatter (Q, K, V -> [tile,head,seq_pad,hd_pad]).
//
// Replaces the cuBLASLt / ATen elementwise launches behind nn.Linear / nn.Conv2d(k=14,s=14) in
//   finetuning/llava/model/multimodal_encoder/siglip_encoder.py:156-173,192-194,207-209,237,252-254
//   finetuning/llava/model/multimodal_projector/builder.py:44-48
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"
#include "gemm_args.h"

namespace rv {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmEpiWarps = 8;   // two warps per TMEM lane quadrant, each takes half of the columns
constexpr int kGemmThreads = 64 + 32 * kGemmEpiWarps;  // warp0 TMA, warp1 MMA, warps 2..9 epilogue

constexpr int kGemmStageWarpBytes = 4096;  // per epilogue warp: 32 rows x 128 B transpose buffer
constexpr int kGemmStagingBytes = kGemmEpiWarps * kGemmStageWarpBytes;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 128) ? 6 : 4;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kGemmStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// (-mean, rstd) of output row `row` when LayerNorm is folded into this GEMM (gemm_args.h); (0, 1) otherwise.  Issued by
// the epilogue warps BEFORE they wait for the accumulator, so the loads cost nothing on the critical path.
__device__ __forceinline__ void gemm_ln_row_stats(const GemmArgs& a, int row, float& nmean, float& rstd) {
  nmean = 0.f;
  rstd = 1.f;
  if (a.ln_s == nullptr || row >= a.M) return;
  if (a.ln_stats != nullptr) {
    const float2 st = __ldg(a.ln_stats + row);
    nmean = -st.x;
    rstd = st.y;
  } else if (a.ln_part != nullptr) {
    const float2* p = a.ln_part + static_cast<size_t>(row) * a.ln_slots;
    float s = 0.f, q = 0.f;
    for (int i = 0; i < a.ln_slots; ++i) {
      const float2 v = p[i];   // written by the previous kernel's epilogue (plain load: not read-only for the program)
      s += v.x;
      q += v.y;
    }
    const float mean = s * a.ln_inv_dim;
    nmean = -mean;
    rstd = rsqrtf(fmaxf(fmaf(-mean, mean, q * a.ln_inv_dim), 0.f) + a.ln_eps);
  }
}

template <int EPI>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmArgs& a, int row, int col0,
                                                    const uint32_t* acc) {
  // acc: 32 fp32 accumulators for (row, col0 .. col0+31)
  if (row >= a.M || col0 >= a.N) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  const bool full = (col0 + 32 <= a.N);

  if (a.ln_s != nullptr) {  // LayerNorm folded into the GEMM (gemm_args.h)
    float nmean, rstd;
    gemm_ln_row_stats(a, row, nmean, rstd);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < a.N) v[j] = rstd * fmaf(nmean, __ldg(a.ln_s + col0 + j), v[j]);
  }
  if (a.bias != nullptr) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(a.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x;
        v[4 * j + 1] += b.y;
        v[4 * j + 2] += b.z;
        v[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < a.N) v[j] += __ldg(a.bias + col0 + j);
    }
  }

  if constexpr (EPI == EPI_GELU_TANH_DUAL_BF16) {
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(a.out2) + static_cast<size_t>(row) * a.ldo + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float y, dy;
      gelu_tanh_both_f(v[j], y, dy);
      if (col0 + j < a.N) o2[j] = __float2bfloat16_rn(dy);
      v[j] = y;
    }
  }
  if constexpr (EPI == EPI_GELU_TANH_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
  }
  if constexpr (EPI == EPI_GELU_ERF_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  }

  if constexpr (EPI == EPI_MUL_BF16) {
    const __nv_bfloat16* m = reinterpret_cast<const __nv_bfloat16*>(a.out2) + static_cast<size_t>(row) * a.ldo + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < a.N) v[j] = __bfloat162float(__float2bfloat16_rn(v[j])) * __bfloat162float(m[j]);
  }
  if constexpr (EPI == EPI_BIAS_F16) {
    __half* o = reinterpret_cast<__half*>(a.out) + static_cast<size_t>(row) * a.ldo + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < a.N) o[j] = __float2half_rn(v[j]);
  } else if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_TANH_BF16 || EPI == EPI_GELU_ERF_BF16 ||
                EPI == EPI_GELU_TANH_DUAL_BF16 || EPI == EPI_MUL_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(row) * a.ldo + col0;
    if (full) {
      uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 p;
        p.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        p.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        p.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        p.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        o4[j] = p;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < a.N) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else if constexpr (EPI == EPI_DELTA_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(row) * a.ldo + col0;
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(a.out2) + static_cast<size_t>(row) * a.ldo + col0;
    const __nv_bfloat16* x = a.aux16 + static_cast<size_t>(row) * a.ldo + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < a.N) {
        const float r = __bfloat162float(x[j]);
        o[j] = __float2bfloat16_rn(v[j]);
        o2[j] = __float2bfloat16_rn(r + v[j]);
      }
  } else if constexpr (EPI == EPI_ATOMIC_F32) {
    float* o = reinterpret_cast<float*>(a.out) + static_cast<size_t>(row) * a.ldo + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < a.N) atomicAdd(o + j, v[j]);
  } else if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_POS_F32 || EPI == EPI_BIAS_F32) {
    float* o = reinterpret_cast<float*>(a.out) + static_cast<size_t>(row) * a.ldo + col0;
    const float* x = nullptr;
    if constexpr (EPI == EPI_RESID_F32) x = a.aux + static_cast<size_t>(row) * a.ldo + col0;
    if constexpr (EPI == EPI_POS_F32) x = a.aux + static_cast<size_t>(row % a.aux_period) * a.N + col0;
    if (full && !(EPI == EPI_RESID_F32 && a.aux16 != nullptr)) {
      float4* o4 = reinterpret_cast<float4*>(o);
      if constexpr (EPI != EPI_BIAS_F32) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        float4 xr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) xr[j] = x4[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 * j + 0] += xr[j].x;
          v[4 * j + 1] += xr[j].y;
          v[4 * j + 2] += xr[j].z;
          v[4 * j + 3] += xr[j].w;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o4[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < a.N) {
          float r = v[j];
          if constexpr (EPI != EPI_BIAS_F32) r += x[j];
          if constexpr (EPI == EPI_RESID_F32)
            if (a.aux16 != nullptr) r += __bfloat162float(a.aux16[static_cast<size_t>(row) * a.ldo + col0 + j]);
          o[j] = r;
        }
    }
  } else if constexpr (EPI == EPI_QKV_SPLIT) {
    const int D = a.heads * a.hd;
    const int tile = row / a.seq;
    const int t = row - tile * a.seq;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = col0 + 8 * g;
      if (c >= a.N) break;
      const int which = c / D;
      const int rem = c - which * D;
      const int head = rem / a.hd;
      const int d = rem - head * a.hd;
      const size_t th = static_cast<size_t>(tile) * a.heads + head;
      __nv_bfloat16* base = (which == 0 ? a.q : (which == 1 ? a.k : a.vt)) + (th * a.seq_pad + t) * a.hd_pad + d;
      uint4 p;
      p.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
      p.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
      p.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
      p.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
      *reinterpret_cast<uint4*>(base) = p;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Coalesced epilogues.  tcgen05.ld hands each thread one accumulator ROW; writing that straight to global
// memory makes every 16-byte access of a warp hit 32 different rows (32 LSU wavefronts per instruction),
// which made the fp32 residual epilogue slower than the K=1152 main loop.  Instead each epilogue warp
// transposes its 32 x 128 B block through a private XOR-swizzled shared-memory buffer and then touches
// global memory with 8 lanes per row (4 full 128-byte lines per instruction).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_store_row(uint32_t stage, int lane, const uint32_t* w /*32 words*/) {
  const uint32_t my = stage + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t addr = my + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * j]), "r"(w[4 * j + 1]),
                 "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                 : "memory");
  }
}
__device__ __forceinline__ uint4 stage_load_vec(uint32_t stage, int rr, int v) {
  uint4 x;
  const uint32_t addr = stage + static_cast<uint32_t>(rr) * 128u + (static_cast<uint32_t>(v ^ (rr & 7)) << 4);
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(addr));
  return x;
}

// fp32 outputs (EPI_RESID_F32 / EPI_POS_F32 / EPI_BIAS_F32): one 32-column chunk, requires N % 4 == 0
// kStats: also accumulate, per output row this thread touches (8 of them: rows i * 4 + lane / 8), the sum and the sum of
// squares of the values it stores (ps / pq); the caller reduces them over the 8 lanes that share a row once per tile.
// kResidSmem (EPI_DELTA_BF16, gemm3_sm100.cuh): the bf16 residual chunk has been prefetched into shared memory at
// `resid_smem` as [32 rows][64 B] (rows beyond M zero-filled) instead of being loaded from a.aux16 here.
template <int EPI, bool kStats = false, bool kResidSmem = false>
__device__ __forceinline__ void gemm_epilogue_f32_staged(const GemmArgs& a, int row0, int col0, const uint32_t* acc,
                                                         uint32_t stage, int lane, f32x2* ps2 = nullptr,
                                                         f32x2* pq2 = nullptr, uint32_t resid_smem = 0) {
  stage_store_row(stage, lane, acc);
  __syncwarp();
  const int v = lane & 7;
  const int gcol = col0 + 4 * v;
  if (gcol < a.N) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(a.bias + gcol));
    const f32x2 b01 = f2_make(b.x, b.y), b23 = f2_make(b.z, b.w);
    float4 xr[8];
    uint2 x16[(EPI == EPI_RESID_F32 || EPI == EPI_DELTA_BF16) ? 8 : 1];   // the bf16 addend / residual (aux16)
    const bool has16 = (EPI == EPI_DELTA_BF16) || (EPI == EPI_RESID_F32 && a.aux16 != nullptr);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int grow = row0 + i * 4 + (lane >> 3);
      xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_DELTA_BF16) {
        x16[i] = make_uint2(0u, 0u);
        if constexpr (kResidSmem) {
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x16[i].x), "=r"(x16[i].y)
                       : "r"(resid_smem + static_cast<uint32_t>(i * 4 + (lane >> 3)) * 64u + static_cast<uint32_t>(v) * 8u));
        } else {
          if (has16 && grow < a.M) x16[i] = *reinterpret_cast<const uint2*>(a.aux16 + static_cast<size_t>(grow) * a.ldo + gcol);
        }
      }
      if (grow < a.M) {
        if constexpr (EPI == EPI_RESID_F32)
          xr[i] = *reinterpret_cast<const float4*>(a.aux + static_cast<size_t>(grow) * a.ldo + gcol);
        if constexpr (EPI == EPI_POS_F32)
          xr[i] = __ldg(reinterpret_cast<const float4*>(a.aux + static_cast<size_t>(grow % a.aux_period) * a.N + gcol));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = i * 4 + (lane >> 3);
      const int grow = row0 + rr;
      if (grow < a.M) {
        const uint4 t = stage_load_vec(stage, rr, v);
        // packed pairs (FADD2 / FFMA2): see gemm_epilogue_bf16_staged
        f32x2 o01 = f2_add(f2_add(f2_make(__uint_as_float(t.x), __uint_as_float(t.y)), b01), f2_make(xr[i].x, xr[i].y));
        f32x2 o23 = f2_add(f2_add(f2_make(__uint_as_float(t.z), __uint_as_float(t.w)), b23), f2_make(xr[i].z, xr[i].w));
        if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_DELTA_BF16) {
          if (has16) {
            const f32x2 a01 = f2_make(__uint_as_float(x16[i].x << 16), __uint_as_float(x16[i].x & 0xFFFF0000u));
            const f32x2 a23 = f2_make(__uint_as_float(x16[i].y << 16), __uint_as_float(x16[i].y & 0xFFFF0000u));
            if constexpr (EPI == EPI_DELTA_BF16) {   // the branch itself goes out in bf16, the stream copy is residual + branch
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(grow) * a.ldo + gcol) =
                  make_uint2(f2_pack_bf16(o01), f2_pack_bf16(o23));
            }
            o01 = f2_add(o01, a01);
            o23 = f2_add(o23, a23);
          }
        }
        float4 o;
        f2_get(o01, o.x, o.y);
        f2_get(o23, o.z, o.w);
        float* dst = reinterpret_cast<float*>(a.out) + static_cast<size_t>(grow) * a.ldo + gcol;
        if constexpr (kStats) {   // ps / pq hold (even column, odd column) partial sums as packed pairs
          ps2[i] = f2_add(ps2[i], f2_add(o01, o23));
          pq2[i] = f2_fma(o01, o01, f2_fma(o23, o23, pq2[i]));
        }
        if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_POS_F32 || EPI == EPI_DELTA_BF16) {
          if (a.out2 != nullptr)  // bf16 copy of the new residual stream (8 lanes x 8 B = 64 contiguous bytes per row)
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out2) + static_cast<size_t>(grow) * a.ldo + gcol) =
                make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
        if constexpr (EPI == EPI_ATOMIC_F32) {   // one 16-byte vector reduction instead of four scalar atomics
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                       : "memory");
        } else if constexpr (EPI != EPI_DELTA_BF16) {
          *reinterpret_cast<float4*>(dst) = o;
        }
      }
    }
  }
  __syncwarp();
}

// bf16 outputs (EPI_BIAS_BF16 / EPI_GELU_*): two 32-column chunks = 64 columns = 128 B per row, requires N % 8 == 0
// kLn: LayerNorm folded into the GEMM (gemm_args.h).  A template parameter, not a run-time test: a (uniform) branch
// inside the unrolled column loop kept ptxas from hoisting the bias / row-sum loads, and the 32 serialised L1 round
// trips per 64 columns made the epilogue of the K = 1152 GEMMs their critical path (QKV + 17 %, fc1 + 14 % in-step).
template <int EPI, bool kLn>
__device__ __forceinline__ void gemm_epilogue_bf16_staged(const GemmArgs& a, int row0, int col0, const uint32_t* acc0,
                                                          const uint32_t* acc1, uint32_t stage, int lane,
                                                          float ln_nmean = 0.f, float ln_rstd = 1.f) {
  // EPI_GELU_TANH_DUAL_BF16 produces two tiles from the same accumulators: gelu(u) -> out and gelu'(u) -> out2 (one tanh
  // for both); they go through the staging buffer one after the other.
  constexpr int kPasses = (EPI == EPI_GELU_TANH_DUAL_BF16) ? 2 : 1;
  uint32_t pk[32], pk2[kPasses == 2 ? 32 : 1];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t* acc = h ? acc1 : acc0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col0 + 32 * h + 4 * j;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.bias != nullptr && c < a.N) b = __ldg(reinterpret_cast<const float4*>(a.bias + c));
      // two packed pairs per 4 columns (FFMA2 / FADD2 / FMUL2): the epilogue's instruction count, not its latency, is
      // what the MMA pipeline of the SM feels
      f32x2 p0 = f2_make(__uint_as_float(acc[4 * j + 0]), __uint_as_float(acc[4 * j + 1]));
      f32x2 p1 = f2_make(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
      const f32x2 b0 = f2_make(b.x, b.y), b1 = f2_make(b.z, b.w);
      if constexpr (kLn) {  // rstd * (acc - mean * s[n]) + b'[n]
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < a.N) s4 = __ldg(reinterpret_cast<const float4*>(a.ln_s + c));
        const f32x2 nm = f2_dup(ln_nmean), rs = f2_dup(ln_rstd);
        p0 = f2_fma(rs, f2_fma(nm, f2_make(s4.x, s4.y), p0), b0);
        p1 = f2_fma(rs, f2_fma(nm, f2_make(s4.z, s4.w), p1), b1);
      } else {
        p0 = f2_add(p0, b0);
        p1 = f2_add(p1, b1);
      }
      if constexpr (EPI == EPI_GELU_TANH_BF16) {
        p0 = gelu_tanh_f2(p0);
        p1 = gelu_tanh_f2(p1);
      }
      if constexpr (EPI == EPI_GELU_TANH_DUAL_BF16) {
        f32x2 d0, d1;
        gelu_tanh_both_f2(p0, p0, d0);
        gelu_tanh_both_f2(p1, p1, d1);
        pk2[16 * h + 2 * j] = f2_pack_bf16(d0);
        pk2[16 * h + 2 * j + 1] = f2_pack_bf16(d1);
      }
      float v0, v1, v2, v3;
      f2_get(p0, v0, v1);
      f2_get(p1, v2, v3);
      if constexpr (EPI == EPI_GELU_ERF_BF16) {
        v0 = gelu_erf_f(v0); v1 = gelu_erf_f(v1); v2 = gelu_erf_f(v2); v3 = gelu_erf_f(v3);
      }
      if constexpr (EPI == EPI_BIAS_F16) {
        pk[16 * h + 2 * j] = pack_f16x2(v0, v1);
        pk[16 * h + 2 * j + 1] = pack_f16x2(v2, v3);
      } else {
        pk[16 * h + 2 * j] = pack_bf16x2(v0, v1);
        pk[16 * h + 2 * j + 1] = pack_bf16x2(v2, v3);
      }
    }
  }
  // EPI_MUL_BF16: all eight multiplier pieces of this lane are requested up front (one L2 round trip, overlapped with
  // the transposition) - loads issued inside the store loop would each wait behind the previous iteration's store
  uint4 mm[EPI == EPI_MUL_BF16 ? 8 : 1];
  if constexpr (EPI == EPI_MUL_BF16) {
    const int gc = col0 + 8 * (lane & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int grow = row0 + i * 4 + (lane >> 3);
      mm[i] = make_uint4(0u, 0u, 0u, 0u);
      if (gc < a.N && grow < a.M)
        mm[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.out2) +
                                                static_cast<size_t>(grow) * a.ldo + gc);
    }
  }
#pragma unroll
  for (int pass = 0; pass < kPasses; ++pass) {
    stage_store_row(stage, lane, pass == 0 ? pk : pk2);
    __syncwarp();
    const int v = lane & 7;
    const int gcol = col0 + 8 * v;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(pass == 0 ? a.out : a.out2);  // (any 16-bit type)
    if (gcol < a.N) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3);
        const int grow = row0 + rr;
        if (grow < a.M) {
          uint4 t = stage_load_vec(stage, rr, v);
          if constexpr (EPI == EPI_MUL_BF16) {
            // out = bf16(acc) * m, m = out2 read coalesced (the 16-byte piece the result is stored to).  Used as
            // dL/du = dL/da * gelu'(u): dL/da is rounded to bf16 first, like a stand-alone GELU backward pass would see it.
            const uint32_t mw[4] = {mm[i].x, mm[i].y, mm[i].z, mm[i].w};
            uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tw[q] = pack_bf16x2(__uint_as_float(tw[q] << 16) * __uint_as_float(mw[q] << 16),
                                  __uint_as_float(tw[q] & 0xFFFF0000u) * __uint_as_float(mw[q] & 0xFFFF0000u));
            t = make_uint4(tw[0], tw[1], tw[2], tw[3]);
          }
          *reinterpret_cast<uint4*>(dst + static_cast<size_t>(grow) * a.ldo + gcol) = t;
        }
      }
    }
    __syncwarp();
  }
}

// QKV head-split scatter, coalesced: 64 columns (two 32-column chunks) of 32 rows go through the swizzled staging
// buffer; afterwards 8 lanes own one row and write its eight 16-byte pieces, which are contiguous in the destination
// ([tile, head, seq_pad, hd_pad] rows of q / k / v) except at a head boundary.  The direct row-per-thread form made every
// warp store touch 32 half-used 32-byte sectors.  Requires hd % 8 == 0 (a 16-byte piece never straddles heads).
template <bool kLn>
__device__ __forceinline__ void gemm_epilogue_qkv_staged(const GemmArgs& a, int row0, int col0, const uint32_t* acc0,
                                                         const uint32_t* acc1, uint32_t stage, int lane,
                                                         float ln_nmean = 0.f, float ln_rstd = 1.f) {
  uint32_t pk[32];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t* acc = h ? acc1 : acc0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col0 + 32 * h + 4 * j;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.bias != nullptr && c < a.N) b = __ldg(reinterpret_cast<const float4*>(a.bias + c));
      f32x2 p0 = f2_make(__uint_as_float(acc[4 * j + 0]), __uint_as_float(acc[4 * j + 1]));
      f32x2 p1 = f2_make(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
      const f32x2 b0 = f2_make(b.x, b.y), b1 = f2_make(b.z, b.w);
      if constexpr (kLn) {  // LayerNorm folded into the GEMM: rstd * (acc - mean * s[n]) + b'[n]
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < a.N) s4 = __ldg(reinterpret_cast<const float4*>(a.ln_s + c));
        const f32x2 nm = f2_dup(ln_nmean), rs = f2_dup(ln_rstd);
        p0 = f2_fma(rs, f2_fma(nm, f2_make(s4.x, s4.y), p0), b0);
        p1 = f2_fma(rs, f2_fma(nm, f2_make(s4.z, s4.w), p1), b1);
      } else {
        p0 = f2_add(p0, b0);
        p1 = f2_add(p1, b1);
      }
      pk[16 * h + 2 * j] = f2_pack_bf16(p0);
      pk[16 * h + 2 * j + 1] = f2_pack_bf16(p1);
    }
  }
  stage_store_row(stage, lane, pk);
  __syncwarp();
  const int v = lane & 7;
  const int c = col0 + 8 * v;
  if (c < a.N) {
    const int D = a.heads * a.hd;
    const int which = c / D;
    const int rem = c - which * D;
    const int head = rem / a.hd;
    const int d = rem - head * a.hd;
    __nv_bfloat16* dst = (which == 0 ? a.q : (which == 1 ? a.k : a.vt)) + d;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = i * 4 + (lane >> 3);
      const int grow = row0 + rr;
      if (grow < a.M) {
        const int tile = grow / a.seq;
        const int t = grow - tile * a.seq;
        const size_t th = static_cast<size_t>(tile) * a.heads + head;
        *reinterpret_cast<uint4*>(dst + (th * a.seq_pad + t) * a.hd_pad) = stage_load_vec(stage, rr, v);
      }
    }
  }
  __syncwarp();
}

// Drain NCOLS accumulator columns of this thread's row: two tcgen05.ld in flight per wait.
template <int EPI, int NCOLS>
__device__ __forceinline__ void gemm_epilogue_drain(const GemmArgs& args, int row, int col_base, uint32_t t_row,
                                                    uint32_t stage, int lane, int ln_slot = -1,
                                                    float ln_nmean = 0.f, float ln_rstd = 1.f) {
  static_assert(NCOLS % 32 == 0, "column span must be a multiple of 32");
  constexpr bool kF32 = (EPI == EPI_RESID_F32 || EPI == EPI_POS_F32 || EPI == EPI_BIAS_F32 || EPI == EPI_ATOMIC_F32 ||
                         EPI == EPI_DELTA_BF16);   // epilogues that work on fp32 values after the transposition
  constexpr bool kBf16 = (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_TANH_BF16 || EPI == EPI_GELU_ERF_BF16 ||
                          EPI == EPI_GELU_TANH_DUAL_BF16 || EPI == EPI_BIAS_F16 ||
                          EPI == EPI_MUL_BF16);  // 16-bit outputs
  const bool staged = ((args.N & 7) == 0) && ((args.ldo & 7) == 0);  // vector validity == column validity
  const int row0 = row - lane;
  // ln_nmean / ln_rstd: this thread's row statistics when LayerNorm is folded into the GEMM (gemm_ln_row_stats)
  constexpr bool kLnCapable = (EPI == EPI_QKV_SPLIT || EPI == EPI_GELU_TANH_BF16 || EPI == EPI_BIAS_BF16 ||
                               EPI == EPI_GELU_ERF_BF16);
  const bool ln = kLnCapable && args.ln_s != nullptr;
  // LayerNorm fold, producer side: row sums of the new residual stream over this warp's columns (gemm_args.h: ln_part)
  constexpr bool kStatsCapable = (EPI == EPI_RESID_F32 || EPI == EPI_POS_F32 || EPI == EPI_DELTA_BF16);
  const bool stats = kStatsCapable && staged && args.ln_part != nullptr && ln_slot >= 0;
  f32x2 ps[kStatsCapable ? 8 : 1], pq[kStatsCapable ? 8 : 1];   // (even column, odd column) partial sums
  if constexpr (kStatsCapable) {
#pragma unroll
    for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0ull;
  }
#pragma unroll 1
  for (int c = 0; c < NCOLS; c += 64) {
    uint32_t r0[32], r1[32];
    const bool two = (c + 32 < NCOLS);
    tmem_ld_x32(t_row + c, r0);
    if (two) tmem_ld_x32(t_row + c + 32, r1);
    tmem_wait_ld();
    if (kF32 && staged) {
      if constexpr (kF32) {
        if constexpr (kStatsCapable) {
          if (stats) {
            gemm_epilogue_f32_staged<EPI, true>(args, row0, col_base + c, r0, stage, lane, ps, pq);
            if (two) gemm_epilogue_f32_staged<EPI, true>(args, row0, col_base + c + 32, r1, stage, lane, ps, pq);
            continue;
          }
        }
        gemm_epilogue_f32_staged<EPI>(args, row0, col_base + c, r0, stage, lane);
        if (two) gemm_epilogue_f32_staged<EPI>(args, row0, col_base + c + 32, r1, stage, lane);
      }
    } else if (kBf16 && staged && two) {
      if constexpr (kBf16) {
        if constexpr (kLnCapable) {
          if (ln) gemm_epilogue_bf16_staged<EPI, true>(args, row0, col_base + c, r0, r1, stage, lane, ln_nmean, ln_rstd);
          else gemm_epilogue_bf16_staged<EPI, false>(args, row0, col_base + c, r0, r1, stage, lane);
        } else {
          gemm_epilogue_bf16_staged<EPI, false>(args, row0, col_base + c, r0, r1, stage, lane);
        }
      }
    } else if (EPI == EPI_QKV_SPLIT && two && (args.N & 7) == 0 && (args.hd & 7) == 0) {
      if constexpr (EPI == EPI_QKV_SPLIT) {
        if (ln) gemm_epilogue_qkv_staged<true>(args, row0, col_base + c, r0, r1, stage, lane, ln_nmean, ln_rstd);
        else gemm_epilogue_qkv_staged<false>(args, row0, col_base + c, r0, r1, stage, lane);
      }
    } else {
      gemm_epilogue_chunk<EPI>(args, row, col_base + c, r0);
      if (two) gemm_epilogue_chunk<EPI>(args, row, col_base + c + 32, r1);
    }
  }
  if constexpr (kStatsCapable) {
    if (stats) {  // the 8 lanes that share a row (lane bits 0..2 = column group) add up, lane & 7 == 0 writes the slot
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s0, s1, q0, q1;
        f2_get(ps[i], s0, s1);
        f2_get(pq[i], q0, q1);
        float sum = s0 + s1, sq = q0 + q1;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        const int grow = row0 + i * 4 + (lane >> 3);
        if ((lane & 7) == 0 && grow < args.M)
          args.ln_part[static_cast<size_t>(grow) * args.ln_slots + ln_slot] = make_float2(sum, sq);
      }
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base + kStages * Cfg::kStageBytes;  // epilogue transpose buffers
  const uint32_t bar_base = stage_base + kGemmStagingBytes;
  // barrier map (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (args.M + kGemmBM - 1) / kGemmBM;
  const int num_n = (args.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (args.K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kGemmEpiWarps);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n;
        const int n_blk = tile - m_blk * num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kGemmBK, m_blk * kGemmBM);
          tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kGemmBK, n_blk * BN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop converged (all operands are warp-uniform); elect.sync inside the asm blocks picks
    // the lane that issues (umma_bf16_ss_elect, common.cuh).
    {
      constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t desc0 = make_smem_desc(smem_base, 1024, kLayoutSw128);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = desc0 + static_cast<uint64_t>((stage * Cfg::kStageBytes) >> 4);
          const uint64_t bdesc = adesc + static_cast<uint64_t>(Cfg::kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_bf16_ss_elect(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_elect(empty_bar(stage));
          if (kb == num_k - 1) umma_commit_elect(tfull_bar(acc));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue (4 warps) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n;
      const int n_blk = tile - m_blk * num_n;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * kGemmBM + quad * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + half * (BN / 2));
      const int col_base = n_blk * BN + half * (BN / 2);
      float ln_nmean, ln_rstd;
      gemm_ln_row_stats(args, row, ln_nmean, ln_rstd);
      gemm_epilogue_drain<EPI, BN / 2>(args, row, col_base, t_row,
                                       stage_base + static_cast<uint32_t>(warp - 2) * kGemmStageWarpBytes, lane, -1,
                                       ln_nmean, ln_rstd);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace rv
