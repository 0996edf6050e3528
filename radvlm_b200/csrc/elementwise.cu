// HBM-bound helper kernels of the tower: LayerNorm (fp32 residual stream -> bf16 GEMM operand),
// patch im2col (pixel tiles -> patch-major bf16 rows for the patch-embed GEMM), fp32 -> bf16 cast.
//
//   LayerNorm : siglip_encoder.py:264,266,287,296 (nn.LayerNorm(1152, eps=1e-6))
//   im2col    : siglip_encoder.py:156-171 (Conv2d k=14 s=14 "valid" == GEMM over 14x14x3 patches;
//               the 6 trailing rows / columns of the 384-px tile are not covered by any patch)
#include "common.cuh"
#include "host_util.h"

#include <cuda_fp16.h>

namespace rv {

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row kept in registers (two-pass mean / variance like ATen).
// ---------------------------------------------------------------------------------------------
template <int NV>  // float4 vectors per lane
__global__ void __launch_bounds__(256)
layernorm_f32_to_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                             const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows,
                             int D, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * D);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    v[i] = (idx < nvec) ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(warp) * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float4 g = __ldg(g4 + idx);
      const float4 b = __ldg(b4 + idx);
      uint2 o;
      o.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      o.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      yr[idx] = o;
    }
  }
}

int layernorm_launch(const float* x, const float* gamma, const float* beta, void* y, int rows, int D,
                     float eps, cudaStream_t stream) {
  RV_CHECK_ARG(x && gamma && beta && y && rows > 0, "layernorm: bad arguments");
  if ((D % 4) != 0 || D > 12 * 128) {
    set_error("layernorm: D=%d unsupported (need D %% 4 == 0 and D <= 1536)", D);
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  const int nv = (D / 4 + 31) / 32;
  __nv_bfloat16* yy = static_cast<__nv_bfloat16*>(y);
  if (nv <= 3)
    layernorm_f32_to_bf16_kernel<3><<<blocks, threads, 0, stream>>>(x, gamma, beta, yy, rows, D, eps);
  else if (nv <= 9)
    layernorm_f32_to_bf16_kernel<9><<<blocks, threads, 0, stream>>>(x, gamma, beta, yy, rows, D, eps);
  else
    layernorm_f32_to_bf16_kernel<12><<<blocks, threads, 0, stream>>>(x, gamma, beta, yy, rows, D, eps);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// Row statistics of the bf16 copy of the residual stream: (mean, rstd) per row, for the LayerNorm that is folded into
// the QKV / fc1 GEMM (gemm_args.h: ln_stats).  Reads 2 B per element (the LayerNorm kernel above moves 6); the
// statistics are those of the values the tensor core will multiply, so the fold's mean subtraction is exact for them.
// One warp per row, two passes over registers like the kernel above.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
ln_row_stats_bf16_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ stats, int rows, int D, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(warp) * D);
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (idx < nvec) u = xr[idx];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[i][2 * q] = __uint_as_float(w[q] << 16);
      v[i][2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
      sum += v[i][2 * q] + v[i][2 * q + 1];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) stats[warp] = make_float2(mean, rsqrtf(sq / static_cast<float>(D) + eps));
}

// (sum, sum of squares) partials written by the residual GEMM epilogues ([rows][slots] float2) -> (mean, rstd) per row
__global__ void __launch_bounds__(256)
ln_finalize_stats_kernel(const float2* __restrict__ part, float2* __restrict__ stats, int rows, int slots, float inv_d,
                         float eps) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float2* p = part + static_cast<size_t>(r) * slots;
  float s = 0.f, q = 0.f;
  for (int i = 0; i < slots; ++i) {
    const float2 v = p[i];
    s += v.x;
    q += v.y;
  }
  const float mean = s * inv_d;
  const float var = fmaxf(fmaf(-mean, mean, q * inv_d), 0.f);
  stats[r] = make_float2(mean, rsqrtf(var + eps));
}

int ln_finalize_stats_launch(const void* part, void* stats, int rows, int slots, int D, float eps, cudaStream_t stream) {
  RV_CHECK_ARG(part && stats && rows > 0 && slots > 0 && D > 0, "ln_finalize_stats: bad arguments");
  ln_finalize_stats_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(static_cast<const float2*>(part),
                                                                     static_cast<float2*>(stats), rows, slots,
                                                                     1.0f / static_cast<float>(D), eps);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

int ln_row_stats_launch(const void* x, void* stats, int rows, int D, float eps, cudaStream_t stream) {
  RV_CHECK_ARG(x && stats && rows > 0, "ln_row_stats: bad arguments");
  if ((D % 8) != 0 || D > 6 * 256) {
    set_error("ln_row_stats: D=%d unsupported (need D %% 8 == 0 and D <= 1536)", D);
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  const int nv = (D / 8 + 31) / 32;
  const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(x);
  float2* st = static_cast<float2*>(stats);
  if (nv <= 2)
    ln_row_stats_bf16_kernel<2><<<blocks, threads, 0, stream>>>(xx, st, rows, D, eps);
  else if (nv <= 5)
    ln_row_stats_bf16_kernel<5><<<blocks, threads, 0, stream>>>(xx, st, rows, D, eps);
  else
    ln_row_stats_bf16_kernel<6><<<blocks, threads, 0, stream>>>(xx, st, rows, D, eps);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (tower output -> projector A operand)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_f32_to_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, size_t nvec) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < nvec; i += stride) {
    const float4 v = x[i];
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    y[i] = o;
  }
}

int cast_f32_bf16_launch(const float* x, void* y, size_t n, cudaStream_t stream) {
  RV_CHECK_ARG(x && y && (n % 4) == 0, "cast: bad arguments (n must be a multiple of 4)");
  const size_t nvec = n / 4;
  size_t blocks = (nvec + 255) / 256;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return RADVLM_OK;
  cast_f32_to_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(y), nvec);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// im2col for the 14x14 stride-14 patch embedding.
//   in : pixels [n, C, S, S]  (fp32 / bf16 / fp16)
//   out: A [n * P * P, Kpad] bf16,  column k = c*ps*ps + ky*ps + kx  (== conv weight.flatten(1)),
//        columns [C*ps*ps, Kpad) zero.
// One CTA per (patch row py, tile): the 3 x 14 x (P*14) strip is read with coalesced row loads.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256)
im2col_patch_kernel(const T* __restrict__ px, __nv_bfloat16* __restrict__ out, int C, int S, int ps,
                    int P, int Kpad) {
  const int py = blockIdx.x;
  const int tile = blockIdx.y;
  const int K = C * ps * ps;
  const int strip_w = P * ps;  // 378
  const T* base = px + static_cast<size_t>(tile) * C * S * S;
  __nv_bfloat16* orow = out + (static_cast<size_t>(tile) * P * P + static_cast<size_t>(py) * P) * Kpad;
  // real columns
  const int total = C * ps * strip_w;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int x = i % strip_w;
    const int cy = i / strip_w;  // c*ps + ky
    const int c = cy / ps;
    const int ky = cy - c * ps;
    const float v = to_f32<T>(base[(static_cast<size_t>(c) * S + (py * ps + ky)) * S + x]);
    const int pxi = x / ps;
    const int kx = x - pxi * ps;
    orow[static_cast<size_t>(pxi) * Kpad + cy * ps + kx] = __float2bfloat16_rn(v);
  }
  // zero padding columns
  const int padw = Kpad - K;
  for (int i = threadIdx.x; i < P * padw; i += blockDim.x) {
    const int pxi = i / padw;
    orow[static_cast<size_t>(pxi) * Kpad + K + (i - pxi * padw)] = __float2bfloat16_rn(0.f);
  }
}

int im2col_launch(const void* pixels, int dtype, void* out, int n_tiles, int C, int S, int ps,
                  int Kpad, cudaStream_t stream) {
  RV_CHECK_ARG(pixels && out && n_tiles > 0, "im2col: bad arguments");
  const int P = S / ps;
  RV_CHECK_ARG(P > 0 && Kpad >= C * ps * ps && (Kpad % 8) == 0, "im2col: bad geometry");
  dim3 grid(P, n_tiles);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (dtype) {
    case RADVLM_DT_F32:
      im2col_patch_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(pixels), o, C, S, ps, P, Kpad);
      break;
    case RADVLM_DT_BF16:
      im2col_patch_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pixels), o, C, S, ps, P, Kpad);
      break;
    case RADVLM_DT_F16:
      im2col_patch_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(pixels), o, C, S, ps, P, Kpad);
      break;
    default:
      set_error("im2col: unknown dtype %d", dtype);
      return RADVLM_ERR_BAD_ARGUMENT;
  }
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// V padding for the attention kernel: zero everything, then a one in column `hd` of every valid key
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
v_ones_column_kernel(__nv_bfloat16* __restrict__ v, int hd, int hd_pad, int seq, int seq_pad) {
  __nv_bfloat16* base = v + static_cast<size_t>(blockIdx.x) * seq_pad * hd_pad + hd;
  for (int t = threadIdx.x; t < seq; t += blockDim.x) base[static_cast<size_t>(t) * hd_pad] = __float2bfloat16(1.0f);
}

// Padding of per-layer q / k / v slots (training keeps them for the backward): only the pad rows (t >= seq) and pad
// columns (d >= hd) are written - the QKV epilogue fills the rest - and column `hd` of every valid V row gets `v_one`
// (1 for the forward's row sums, 0 for the backward, whose dP = dO V^T must not see it).  q / k may be null.
__global__ void __launch_bounds__(256)
qkv_pad_prepare_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k, __nv_bfloat16* __restrict__ v, int hd,
                       int hd_pad, int seq, int seq_pad, float v_one) {
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const __nv_bfloat16 one = __float2bfloat16(v_one);
  const uint4 first = make_uint4(static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&one)), 0, 0, 0);
  const int row_vecs = hd_pad / 8, pad0 = hd / 8;
  for (int t = threadIdx.x; t < seq_pad; t += blockDim.x) {
    const size_t off = (static_cast<size_t>(blockIdx.x) * seq_pad + t) * hd_pad;
    const int v0 = t < seq ? pad0 : 0;
    for (int i = v0; i < row_vecs; ++i) {
      if (q != nullptr) reinterpret_cast<uint4*>(q + off)[i] = zero;
      if (k != nullptr) reinterpret_cast<uint4*>(k + off)[i] = zero;
      reinterpret_cast<uint4*>(v + off)[i] = (t < seq && i == pad0) ? first : zero;
    }
  }
}

int qkv_pad_prepare_launch(void* q, void* k, void* vt, int tiles, int heads, int seq, int seq_pad, int hd, int hd_pad,
                           float v_one, cudaStream_t stream) {
  RV_CHECK_ARG(vt && tiles > 0 && heads > 0 && hd < hd_pad && (hd % 8) == 0 && (hd_pad % 8) == 0 && seq <= seq_pad,
               "qkv_pad_prepare: bad arguments");
  qkv_pad_prepare_kernel<<<tiles * heads, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(q), static_cast<__nv_bfloat16*>(k),
                                                            static_cast<__nv_bfloat16*>(vt), hd, hd_pad, seq, seq_pad, v_one);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

int attention_prepare_vt_launch(void* vt, int tiles, int heads, int seq, int seq_pad, int hd, int hd_pad,
                                cudaStream_t stream) {
  RV_CHECK_ARG(vt && tiles > 0 && heads > 0 && hd < hd_pad && seq <= seq_pad, "prepare_vt: bad arguments");
  const size_t bytes = static_cast<size_t>(tiles) * heads * hd_pad * seq_pad * 2;
  RV_CUDA(cudaMemsetAsync(vt, 0, bytes, stream));
  v_ones_column_kernel<<<tiles * heads, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(vt), hd, hd_pad, seq, seq_pad);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

}  // namespace rv

extern "C" int radvlm_attention_prepare_vt(void* vt, int tiles, int heads, int seq, int seq_pad, int hd, int hd_pad,
                                           void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::attention_prepare_vt_launch(vt, tiles, heads, seq, seq_pad, hd, hd_pad, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_layernorm_f32_bf16(const float* x, const float* gamma, const float* beta, void* y,
                                         int rows, int D, float eps, void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::layernorm_launch(x, gamma, beta, y, rows, D, eps, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_ln_row_stats_bf16(const void* x, void* stats, int rows, int D, float eps, void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::ln_row_stats_launch(x, stats, rows, D, eps, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_cast_f32_bf16(const float* x, void* y, size_t n, void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::cast_f32_bf16_launch(x, y, n, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_patch_im2col(const void* pixels, int dtype, void* out, int n_tiles, int channels,
                                   int image_size, int patch_size, int k_pad, void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::im2col_launch(pixels, dtype, out, n_tiles, channels, image_size, patch_size, k_pad,
                           static_cast<cudaStream_t>(stream));
}
