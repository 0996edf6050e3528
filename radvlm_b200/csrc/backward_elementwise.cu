// HBM-bound kernels of the backward pass (training mode, mm_tunable_parts = vision tower + projector):
//
//   colsum            : bias gradients of every nn.Linear (db = sum over rows of dY)
//   gelu_fwd_bwd      : a = gelu(u), du = da * gelu'(u) in one pass (tanh form: SigLipMLP, siglip_encoder.py:247-254;
//                       erf form: mm_projector, builder.py:44-46)
//   layernorm_bwd     : nn.LayerNorm backward (siglip_encoder.py:264,266), dx added to the fp32 residual gradient,
//                       dgamma / dbeta reduced per block in shared memory, then one atomic per column per block
//   pos_embed_grad    : gradient of the position table = sum over tiles of d(hidden0)  (siglip_encoder.py:173)
#include "common.cuh"
#include "host_util.h"
#include "internal.h"

namespace rv {

// ---------------------------------------------------------------------------------------------
// column sums of a bf16 [rows, cols] matrix, accumulated into fp32 out[cols]
// ---------------------------------------------------------------------------------------------
constexpr int kColsumRows = 256;  // rows per block

__global__ void __launch_bounds__(128)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int rows, int cols, int ld, float* __restrict__ out) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= cols) return;
  const int r0 = blockIdx.y * kColsumRows;
  const int r1 = min(rows, r0 + kColsumRows);
  float s0 = 0.f, s1 = 0.f;
  const __nv_bfloat16* p = x + static_cast<size_t>(r0) * ld + c;
  int r = r0;
  for (; r + 8 <= r1; r += 8, p += 8 * static_cast<size_t>(ld)) {   // eight independent row loads in flight
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = *reinterpret_cast<const uint32_t*>(p + static_cast<size_t>(u) * ld);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s0 += __uint_as_float(w[u] << 16);
      s1 += __uint_as_float(w[u] & 0xFFFF0000u);
    }
  }
  for (; r < r1; ++r, p += ld) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
    s0 += __uint_as_float(w << 16);
    s1 += __uint_as_float(w & 0xFFFF0000u);
  }
  atomicAdd(out + c, s0);
  if (c + 1 < cols) atomicAdd(out + c + 1, s1);
}

int colsum_bf16_launch(const void* x, int rows, int cols, int ld, float* out, cudaStream_t stream) {
  RV_CHECK_ARG(x && out && rows > 0 && cols > 0 && (cols % 2) == 0 && (ld % 2) == 0, "colsum: bad arguments");
  dim3 grid((cols / 2 + 127) / 128, (rows + kColsumRows - 1) / kColsumRows);
  colsum_bf16_kernel<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), rows, cols, ld, out);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// GELU forward + backward, bf16, 8 elements per thread
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gelu_tanh_both(float x, float& y, float& dy) {
  const float k0 = 0.7978845608028654f, c = 0.044715f;
  const float x2 = x * x;
  const float t = tanhf(k0 * x * fmaf(c, x2, 1.0f));
  y = 0.5f * x * (1.0f + t);
  dy = 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * fmaf(3.0f * c, x2, 1.0f);
}
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.7071067811865476f));
  y = x * cdf;
  dy = cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

template <int KIND>  // 0 tanh, 1 erf
__global__ void __launch_bounds__(256)
gelu_fwd_bwd_kernel(const uint4* __restrict__ u, uint4* __restrict__ da_du, uint4* __restrict__ a, size_t nvec) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < nvec; i += stride) {
    const uint4 uv = u[i], gv = da_du[i];
    const uint32_t uw[4] = {uv.x, uv.y, uv.z, uv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t ao[4], go[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y0, y1, d0, d1;
      if (KIND == 0) {
        gelu_tanh_both(__uint_as_float(uw[j] << 16), y0, d0);
        gelu_tanh_both(__uint_as_float(uw[j] & 0xFFFF0000u), y1, d1);
      } else {
        gelu_erf_both(__uint_as_float(uw[j] << 16), y0, d0);
        gelu_erf_both(__uint_as_float(uw[j] & 0xFFFF0000u), y1, d1);
      }
      ao[j] = pack_bf16x2(y0, y1);
      go[j] = pack_bf16x2(d0 * __uint_as_float(gw[j] << 16), d1 * __uint_as_float(gw[j] & 0xFFFF0000u));
    }
    a[i] = make_uint4(ao[0], ao[1], ao[2], ao[3]);
    da_du[i] = make_uint4(go[0], go[1], go[2], go[3]);
  }
}

// u: pre-activation (bf16); da_du: on entry dL/da, on exit dL/du; a: gelu(u).  n % 8 == 0.
int gelu_fwd_bwd_launch(const void* u, void* da_du, void* a, size_t n, int erf_form, cudaStream_t stream) {
  RV_CHECK_ARG(u && da_du && a && (n % 8) == 0, "gelu_fwd_bwd: bad arguments");
  const size_t nvec = n / 8;
  size_t blocks = (nvec + 255) / 256;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return RADVLM_OK;
  if (erf_form)
    gelu_fwd_bwd_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const uint4*>(u), static_cast<uint4*>(da_du), static_cast<uint4*>(a), nvec);
  else
    gelu_fwd_bwd_kernel<0><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const uint4*>(u), static_cast<uint4*>(da_du), static_cast<uint4*>(a), nvec);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward, two HBM-bound passes (D <= 1536, D % 4 == 0):
//   (1) one warp per row:  xhat = (x - mean) * rstd,  g = dy * gamma,
//                          dres += rstd * (g - mean(g) - xhat * mean(g * xhat));   (mean, rstd) kept per row
//   (2) column sums:       dgamma += sum_r dy * xhat,  dbeta += sum_r dy   (threads own columns, strips of rows;
//                          coalesced row reads, two atomics per column per strip)
// A single fused pass (per-row parameter contributions reduced in registers or shared-memory atomics) measured
// 2.9x / 2.1x slower than the row pass alone: the reduction, not the traffic, bound it.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dy,
                        float* __restrict__ dres, __nv_bfloat16* __restrict__ dres_bf16, float2* __restrict__ stats, int rows,
                        int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = D >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float inv_d = 1.0f / static_cast<float>(D);
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * D);
  // Only the row (as xhat) stays in registers; g = dy * gamma is formed twice from dy / gamma (second read hits L1 / L2):
  // ~60 registers instead of ~100 doubles the resident warps of this latency-bound pass.
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
  const float mean = sum * inv_d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * inv_d + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const uint2 d = dyr[idx];
      const float4 gm = __ldg(g4 + idx);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
      const float gx = __uint_as_float(d.x << 16) * gm.x, gy = __uint_as_float(d.x & 0xFFFF0000u) * gm.y;
      const float gz = __uint_as_float(d.y << 16) * gm.z, gw = __uint_as_float(d.y & 0xFFFF0000u) * gm.w;
      m1 += (gx + gy) + (gz + gw);
      m2 += (gx * v[i].x + gy * v[i].y) + (gz * v[i].z + gw * v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m1 += __shfl_xor_sync(0xffffffffu, m1, o);
    m2 += __shfl_xor_sync(0xffffffffu, m2, o);
  }
  m1 *= inv_d;
  m2 *= inv_d;
  float4* dr = reinterpret_cast<float4*>(dres + static_cast<size_t>(row) * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const uint2 d = dyr[idx];
      const float4 gm = __ldg(g4 + idx);
      float4 o = dr[idx];
      o.x += rstd * (__uint_as_float(d.x << 16) * gm.x - m1 - v[i].x * m2);
      o.y += rstd * (__uint_as_float(d.x & 0xFFFF0000u) * gm.y - m1 - v[i].y * m2);
      o.z += rstd * (__uint_as_float(d.y << 16) * gm.z - m1 - v[i].z * m2);
      o.w += rstd * (__uint_as_float(d.y & 0xFFFF0000u) * gm.w - m1 - v[i].w * m2);
      dr[idx] = o;
      if (dres_bf16 != nullptr)  // bf16 copy of the updated residual gradient: the next dgrad / wgrad GEMMs read it
        reinterpret_cast<uint2*>(dres_bf16 + static_cast<size_t>(row) * D)[idx] =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

constexpr int kLnParamRows = 128;  // rows per block of the parameter-gradient pass

__global__ void __launch_bounds__(128)
layernorm_bwd_params_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                            const float2* __restrict__ stats, float* __restrict__ dgamma, float* __restrict__ dbeta,
                            int rows, int D) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= D) return;
  const int r0 = blockIdx.y * kLnParamRows;
  const int r1 = min(rows, r0 + kLnParamRows);
  float g0 = 0.f, g1 = 0.f, b0 = 0.f, b1 = 0.f;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {   // four independent row loads in flight per thread
    float2 st[4], xv[4];
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      st[u] = __ldg(stats + r + u);
      xv[u] = *reinterpret_cast<const float2*>(x + static_cast<size_t>(r + u) * D + c);
      w[u] = *reinterpret_cast<const uint32_t*>(dy + static_cast<size_t>(r + u) * D + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d0 = __uint_as_float(w[u] << 16), d1 = __uint_as_float(w[u] & 0xFFFF0000u);
      g0 += d0 * (xv[u].x - st[u].x) * st[u].y;
      g1 += d1 * (xv[u].y - st[u].x) * st[u].y;
      b0 += d0;
      b1 += d1;
    }
  }
  for (; r < r1; ++r) {
    const float2 st = __ldg(stats + r);
    const float2 xv = *reinterpret_cast<const float2*>(x + static_cast<size_t>(r) * D + c);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(dy + static_cast<size_t>(r) * D + c);
    const float d0 = __uint_as_float(w << 16), d1 = __uint_as_float(w & 0xFFFF0000u);
    g0 += d0 * (xv.x - st.x) * st.y;
    g1 += d1 * (xv.y - st.x) * st.y;
    b0 += d0;
    b1 += d1;
  }
  atomicAdd(dgamma + c, g0);
  atomicAdd(dgamma + c + 1, g1);
  atomicAdd(dbeta + c, b0);
  atomicAdd(dbeta + c + 1, b1);
}

// stats: scratch of rows * 8 bytes
int layernorm_bwd_launch(const float* x, const float* gamma, const void* dy, float* dres, float* dgamma, float* dbeta,
                         void* stats, int rows, int D, float eps, cudaStream_t stream, void* dres_bf16) {
  RV_CHECK_ARG(x && gamma && dy && dres && stats && rows > 0, "layernorm_bwd: bad arguments");
  if ((D % 4) != 0 || D > 12 * 128) {
    set_error("layernorm_bwd: D=%d unsupported (need D %% 4 == 0 and D <= 1536)", D);
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  const int nv = (D / 4 + 31) / 32;
  const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(dy);
  float2* st = static_cast<float2*>(stats);
  if (nv <= 3)
    layernorm_bwd_dx_kernel<3><<<blocks, threads, 0, stream>>>(x, gamma, d, dres, static_cast<__nv_bfloat16*>(dres_bf16), st, rows, D, eps);
  else if (nv <= 9)
    layernorm_bwd_dx_kernel<9><<<blocks, threads, 0, stream>>>(x, gamma, d, dres, static_cast<__nv_bfloat16*>(dres_bf16), st, rows, D, eps);
  else
    layernorm_bwd_dx_kernel<12><<<blocks, threads, 0, stream>>>(x, gamma, d, dres, static_cast<__nv_bfloat16*>(dres_bf16), st, rows, D, eps);
  RV_CUDA(cudaGetLastError());
  if (dgamma != nullptr && dbeta != nullptr) {
    dim3 grid((D / 2 + 127) / 128, (rows + kLnParamRows - 1) / kLnParamRows);
    layernorm_bwd_params_kernel<<<grid, 128, 0, stream>>>(x, d, st, dgamma, dbeta, rows, D);
    RV_CUDA(cudaGetLastError());
  }
  return RADVLM_OK;
}

// ---------------------------------------------------------------------------------------------
// dpos[t, c] += sum over tiles of dh[tile * T + t, c]   (fp32)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pos_embed_grad_kernel(const float4* __restrict__ dh, float4* __restrict__ dpos, int tiles, size_t per_tile_vec) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per_tile_vec) return;
  float4 acc = dpos[i];
  for (int t = 0; t < tiles; ++t) {
    const float4 v = dh[static_cast<size_t>(t) * per_tile_vec + i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  dpos[i] = acc;
}

int pos_embed_grad_launch(const float* dh, float* dpos, int tiles, int T, int D, cudaStream_t stream) {
  RV_CHECK_ARG(dh && dpos && tiles > 0 && (D % 4) == 0, "pos_embed_grad: bad arguments");
  const size_t per_tile_vec = static_cast<size_t>(T) * D / 4;
  pos_embed_grad_kernel<<<static_cast<unsigned>((per_tile_vec + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dh), reinterpret_cast<float4*>(dpos), tiles, per_tile_vec);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

}  // namespace rv

extern "C" int radvlm_colsum_bf16(const void* x, int rows, int cols, int ld, float* out, void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::colsum_bf16_launch(x, rows, cols, ld, out, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_gelu_fwd_bwd_bf16(const void* u, void* da_du, void* a, int64_t n, int erf_form, void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::gelu_fwd_bwd_launch(u, da_du, a, static_cast<size_t>(n), erf_form, static_cast<cudaStream_t>(stream));
}

extern "C" int radvlm_layernorm_bwd(const float* x, const float* gamma, const void* dy, float* dres, float* dgamma,
                                    float* dbeta, void* row_stats_scratch, int rows, int D, float eps, void* stream) {
  int st = rv::require_sm100();
  if (st != RADVLM_OK) return st;
  return rv::layernorm_bwd_launch(x, gamma, dy, dres, dgamma, dbeta, row_stats_scratch, rows, D, eps,
                                  static_cast<cudaStream_t>(stream), nullptr);
}
