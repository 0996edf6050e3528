// Merge + splice gather kernel (HBM-bound): one warp produces one output row of the padded
// [B, max_len, H] inputs_embeds tensor, reading its source exactly once.
//
//   spatial unpad / anyres_max pool / newline / base prepend   llava_arch.py:350-412 (unpad_image 127-159)
//   text embedding gather + interleave with image tokens       llava_arch.py:449-493
//   truncate / pad / stack, labels, attention_mask, position   llava_arch.py:495-531
//
// Row -> source resolution: warp-uniform binary search in the (sorted) segment table, then closed-form
// index arithmetic.  The bilinear branch follows ATen upsample_bilinear2d
// (align_corners=False): scale = in/out in fp32, src = max(scale*(i+0.5)-0.5, 0), 4-tap lerp in fp32.
#include <cuda_fp16.h>

#include <cstring>

#include "common.cuh"
#include "host_util.h"

namespace rv {

template <typename T>
struct Vec16;  // 16-byte vector of T <-> float[]
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                      pack_bf16x2(f[6], f[7]));
  }
};
template <>
struct Vec16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 p = __half22float2(h);
      f[2 * i] = p.x;
      f[2 * i + 1] = p.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

struct MergeSpliceArgs {
  const uint4* features;
  const uint4* newline;
  const uint4* embed;
  int nvec;  // 16-byte vectors per row
  int T;     // tokens per tile (729)
  int S;     // patches per side (27)
  const int64_t* input_ids;
  const int64_t* labels_in;
  const int32_t* text_src;
  const radvlm_splice_segment* segments;
  int n_segments;
  const radvlm_merge_image* images;
  int64_t total_rows;
  uint4* out;
  uint4* out_peers[RADVLM_MAX_PEERS];  // scatter form: the slot of this rank in every rank's gathered buffer
  int n_peers;
  int64_t* out_labels;
  uint8_t* out_mask;
  int64_t* out_pos;
  int64_t ignore_index;
};

constexpr int kMsUnroll = 7;

__device__ __forceinline__ void copy_row(const uint4* __restrict__ src, uint4* __restrict__ dst, int nvec,
                                         int lane) {
  int i = lane;
  for (; i + 32 * (kMsUnroll - 1) < nvec; i += 32 * kMsUnroll) {
    uint4 v[kMsUnroll];
#pragma unroll
    for (int u = 0; u < kMsUnroll; ++u) v[u] = ld_stream(src + i + 32 * u);
#pragma unroll
    for (int u = 0; u < kMsUnroll; ++u) st_stream(dst + i + 32 * u, v[u]);
  }
  for (; i < nvec; i += 32) st_stream(dst + i, ld_stream(src + i));
}

// Destination row(s) of one output row.  kScatter = false: the local inputs_embeds tensor.  kScatter = true (fused
// merge + all-gather): the same row of every rank's gathered buffer, written with plain stores through
// peer-mapped pointers (NVLink / NVSwitch), so the source is read once and no separate collective moves the tokens.
template <bool kScatter>
struct RowDst {
  uint4* p[kScatter ? RADVLM_MAX_PEERS : 1];
  int n;
  __device__ __forceinline__ void st(int i, const uint4& v) const {
    if (kScatter) {
#pragma unroll
      for (int d = 0; d < RADVLM_MAX_PEERS; ++d)
        if (d < n) st_stream(p[d] + i, v);
    } else {
      st_stream(p[0] + i, v);
    }
  }
};
template <bool kScatter>
__device__ __forceinline__ void copy_row(const uint4* __restrict__ src, const RowDst<kScatter>& dst, int nvec, int lane) {
  int i = lane;
  for (; i + 32 * (kMsUnroll - 1) < nvec; i += 32 * kMsUnroll) {
    uint4 v[kMsUnroll];
#pragma unroll
    for (int u = 0; u < kMsUnroll; ++u) v[u] = ld_stream(src + i + 32 * u);
#pragma unroll
    for (int u = 0; u < kMsUnroll; ++u) dst.st(i + 32 * u, v[u]);
  }
  for (; i < nvec; i += 32) dst.st(i, ld_stream(src + i));
}

// feature row (in 16-byte vectors) of grid position (R, C) of the un-cropped S*gh x S*gw map
__device__ __forceinline__ size_t grid_src_row(const radvlm_merge_image& im, int R, int C, int S, int T) {
  const int tr = R / S, tc = C / S;
  return static_cast<size_t>(im.tile_base + 1 + tr * im.grid_w + tc) * T + (R - tr * S) * S + (C - tc * S);
}


// Video samples (llava_arch.py:171-190 get_2dPool with stride 2, :222-249 add_token_per_grid / add_token_per_frame,
// :310-349): `grid_w` frames of S x S tokens are pooled to out_h x out_w (bilinear: ceil(S/2), ATen
// upsample_bilinear2d align_corners=False; average / max: floor(S/2), 2x2 windows) and laid out frame-major with a
// newline token per pooled row (GRID), per frame (FRAME), once at the end (ONE) or not at all (NONE).
struct VideoTaps {
  int newline;    // the token is the image_newline row
  size_t row[4];  // feature rows of the (up to) four taps
  float ly1, lx1; // bilinear weights of the second row / column
};
__device__ __forceinline__ VideoTaps video_source(const radvlm_merge_image& im, int t, int S, int T) {
  VideoTaps v;
  v.newline = 0;
  v.ly1 = v.lx1 = 0.f;
  const int per_frame = im.out_h * im.out_w;
  int f, r, c;
  if (im.reserved == RADVLM_NEWLINE_GRID) {
    const int stride = im.out_h * (im.out_w + 1);
    f = t / stride;
    const int u = t - f * stride;
    r = u / (im.out_w + 1);
    c = u - r * (im.out_w + 1);
    if (c == im.out_w) v.newline = 1;
  } else if (im.reserved == RADVLM_NEWLINE_FRAME) {
    f = t / (per_frame + 1);
    const int u = t - f * (per_frame + 1);
    if (u == per_frame) v.newline = 1;
    r = u / im.out_w;
    c = u - r * im.out_w;
  } else {
    f = t / per_frame;
    if (f >= im.grid_w) v.newline = 1;  // the single trailing newline (ONE)
    const int u = t - f * per_frame;
    r = u / im.out_w;
    c = u - r * im.out_w;
  }
  if (v.newline) {
    v.row[0] = v.row[1] = v.row[2] = v.row[3] = 0;
    return v;
  }
  int y0, y1, x0, x1;
  if (im.pool == RADVLM_POOL_BILINEAR) {
    const float sh = static_cast<float>(S) / static_cast<float>(im.out_h);
    const float sw = static_cast<float>(S) / static_cast<float>(im.out_w);
    float sy = sh * (static_cast<float>(r) + 0.5f) - 0.5f;
    float sx = sw * (static_cast<float>(c) + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    y0 = static_cast<int>(sy);
    x0 = static_cast<int>(sx);
    y1 = y0 + (y0 < S - 1 ? 1 : 0);
    x1 = x0 + (x0 < S - 1 ? 1 : 0);
    v.ly1 = sy - static_cast<float>(y0);
    v.lx1 = sx - static_cast<float>(x0);
  } else {  // 2x2 window
    y0 = 2 * r; y1 = y0 + 1;
    x0 = 2 * c; x1 = x0 + 1;
  }
  const size_t base = static_cast<size_t>(im.tile_base + f) * T;
  v.row[0] = base + y0 * S + x0;
  v.row[1] = base + y0 * S + x1;
  v.row[2] = base + y1 * S + x0;
  v.row[3] = base + y1 * S + x1;
  return v;
}

template <typename T, bool kScatter>
__global__ void __launch_bounds__(256)
merge_splice_kernel(const MergeSpliceArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp0; row < a.total_rows; row += nwarps) {
    // warp-uniform binary search: last segment with dst_row <= row
    int lo = 0, hi = a.n_segments - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(&a.segments[mid].dst_row) <= row) lo = mid; else hi = mid - 1;
    }
    const radvlm_splice_segment seg = a.segments[lo];
    const int off = static_cast<int>(row - seg.dst_row);
    RowDst<kScatter> dst;
    if (kScatter) {
      dst.n = a.n_peers;
#pragma unroll
      for (int d = 0; d < RADVLM_MAX_PEERS; ++d) dst.p[d] = (d < a.n_peers ? a.out_peers[d] : a.out) + static_cast<size_t>(row) * a.nvec;
    } else {
      dst.n = 1;
      dst.p[0] = a.out + static_cast<size_t>(row) * a.nvec;
    }
    int64_t label = a.ignore_index;
    uint8_t mask = 1;
    int64_t pos = seg.pos0 + off;

    if (seg.kind == RADVLM_SEG_PAD || off >= seg.length) {
      for (int i = lane; i < a.nvec; i += 32) dst.st(i, make_uint4(0, 0, 0, 0));
      mask = 0;
      pos = 0;
    } else if (seg.kind == RADVLM_SEG_TEXT) {
      const int32_t sp = __ldg(a.text_src + seg.src_off + off);
      const int64_t tok = __ldg(a.input_ids + sp);
      if (a.labels_in != nullptr) label = __ldg(a.labels_in + sp);
      copy_row(a.embed + static_cast<size_t>(tok) * a.nvec, dst, a.nvec, lane);
    } else {
      const radvlm_merge_image im = a.images[seg.image];
      const int t = seg.src_off + off;
      if (im.mode == RADVLM_MERGE_FLAT) {
        copy_row(a.features + (static_cast<size_t>(im.tile_base) * a.T + t) * a.nvec, dst, a.nvec, lane);
      } else if (im.mode == RADVLM_MERGE_VIDEO) {
        const VideoTaps v = video_source(im, t, a.S, a.T);
        if (v.newline) {
          copy_row(a.newline, dst, a.nvec, lane);
        } else {
          const uint4* p00 = a.features + v.row[0] * a.nvec;
          const uint4* p01 = a.features + v.row[1] * a.nvec;
          const uint4* p10 = a.features + v.row[2] * a.nvec;
          const uint4* p11 = a.features + v.row[3] * a.nvec;
          const float ly0 = 1.f - v.ly1, lx0 = 1.f - v.lx1;
          for (int i = lane; i < a.nvec; i += 32) {
            float f00[Vec16<T>::N], f01[Vec16<T>::N], f10[Vec16<T>::N], f11[Vec16<T>::N], o[Vec16<T>::N];
            Vec16<T>::unpack(ld_stream(p00 + i), f00);
            Vec16<T>::unpack(ld_stream(p01 + i), f01);
            Vec16<T>::unpack(ld_stream(p10 + i), f10);
            Vec16<T>::unpack(ld_stream(p11 + i), f11);
#pragma unroll
            for (int e = 0; e < Vec16<T>::N; ++e) {
              if (im.pool == RADVLM_POOL_BILINEAR)
                o[e] = ly0 * (lx0 * f00[e] + v.lx1 * f01[e]) + v.ly1 * (lx0 * f10[e] + v.lx1 * f11[e]);
              else if (im.pool == RADVLM_POOL_AVERAGE)
                o[e] = (((f00[e] + f01[e]) + f10[e]) + f11[e]) * 0.25f;
              else
                o[e] = fmaxf(fmaxf(f00[e], f01[e]), fmaxf(f10[e], f11[e]));
            }
            dst.st(i, Vec16<T>::pack(o));
          }
        }
      } else if (im.mode == RADVLM_MERGE_SINGLE) {
        if (t < a.T) copy_row(a.features + (static_cast<size_t>(im.tile_base) * a.T + t) * a.nvec, dst, a.nvec, lane);
        else copy_row(a.newline, dst, a.nvec, lane);
      } else if (!(im.reserved & RADVLM_ANYRES_NO_BASE) && t < a.T) {  // base tile
        copy_row(a.features + (static_cast<size_t>(im.tile_base) * a.T + t) * a.nvec, dst, a.nvec, lane);
      } else {
        const int u = t - ((im.reserved & RADVLM_ANYRES_NO_BASE) ? 0 : a.T);
        const int row_w = im.out_w + ((im.reserved & RADVLM_ANYRES_NO_NEWLINE) ? 0 : 1);
        const int r = u / row_w;
        const int c = u - r * row_w;
        if (c == im.out_w) {
          copy_row(a.newline, dst, a.nvec, lane);
        } else if (im.pool == RADVLM_POOL_MAX) {  // 'maxpool2x2': 2 x 2 windows of the whole grid (llava_arch.py:376-380)
          const uint4* p00 = a.features + grid_src_row(im, 2 * r, 2 * c, a.S, a.T) * a.nvec;
          const uint4* p01 = a.features + grid_src_row(im, 2 * r, 2 * c + 1, a.S, a.T) * a.nvec;
          const uint4* p10 = a.features + grid_src_row(im, 2 * r + 1, 2 * c, a.S, a.T) * a.nvec;
          const uint4* p11 = a.features + grid_src_row(im, 2 * r + 1, 2 * c + 1, a.S, a.T) * a.nvec;
          for (int i = lane; i < a.nvec; i += 32) {
            float f00[Vec16<T>::N], f01[Vec16<T>::N], f10[Vec16<T>::N], f11[Vec16<T>::N], o[Vec16<T>::N];
            Vec16<T>::unpack(ld_stream(p00 + i), f00);
            Vec16<T>::unpack(ld_stream(p01 + i), f01);
            Vec16<T>::unpack(ld_stream(p10 + i), f10);
            Vec16<T>::unpack(ld_stream(p11 + i), f11);
#pragma unroll
            for (int e = 0; e < Vec16<T>::N; ++e) o[e] = fmaxf(fmaxf(f00[e], f01[e]), fmaxf(f10[e], f11[e]));
            dst.st(i, Vec16<T>::pack(o));
          }
        } else if (!im.pool) {
          const size_t srow = grid_src_row(im, r + im.crop_r0, c + im.crop_c0, a.S, a.T);
          copy_row(a.features + srow * a.nvec, dst, a.nvec, lane);
        } else {
          const float sh = static_cast<float>(im.crop_h) / static_cast<float>(im.out_h);
          const float sw = static_cast<float>(im.crop_w) / static_cast<float>(im.out_w);
          float sy = sh * (static_cast<float>(r) + 0.5f) - 0.5f;
          float sx = sw * (static_cast<float>(c) + 0.5f) - 0.5f;
          sy = sy < 0.f ? 0.f : sy;
          sx = sx < 0.f ? 0.f : sx;
          const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
          const int y1 = y0 + (y0 < im.crop_h - 1 ? 1 : 0), x1 = x0 + (x0 < im.crop_w - 1 ? 1 : 0);
          const float ly1 = sy - static_cast<float>(y0), lx1 = sx - static_cast<float>(x0);
          const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
          const uint4* p00 = a.features + grid_src_row(im, y0 + im.crop_r0, x0 + im.crop_c0, a.S, a.T) * a.nvec;
          const uint4* p01 = a.features + grid_src_row(im, y0 + im.crop_r0, x1 + im.crop_c0, a.S, a.T) * a.nvec;
          const uint4* p10 = a.features + grid_src_row(im, y1 + im.crop_r0, x0 + im.crop_c0, a.S, a.T) * a.nvec;
          const uint4* p11 = a.features + grid_src_row(im, y1 + im.crop_r0, x1 + im.crop_c0, a.S, a.T) * a.nvec;
          for (int i = lane; i < a.nvec; i += 32) {
            float f00[Vec16<T>::N], f01[Vec16<T>::N], f10[Vec16<T>::N], f11[Vec16<T>::N], o[Vec16<T>::N];
            Vec16<T>::unpack(ld_stream(p00 + i), f00);
            Vec16<T>::unpack(ld_stream(p01 + i), f01);
            Vec16<T>::unpack(ld_stream(p10 + i), f10);
            Vec16<T>::unpack(ld_stream(p11 + i), f11);
#pragma unroll
            for (int e = 0; e < Vec16<T>::N; ++e)
              o[e] = ly0 * (lx0 * f00[e] + lx1 * f01[e]) + ly1 * (lx0 * f10[e] + lx1 * f11[e]);
            dst.st(i, Vec16<T>::pack(o));
          }
        }
      }
    }
    if (lane == 0) {
      if (a.out_labels) a.out_labels[row] = label;
      if (a.out_mask) a.out_mask[row] = mask;
      if (a.out_pos) a.out_pos[row] = pos;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Backward of the gather above (autograd of llava_arch.py:350-531): one warp per output row scatters its gradient
// to where the forward read it from.  Feature / newline gradients are fp32 and accumulated with atomics (a pooled
// row feeds four sources, the newline feeds many rows); text rows are copied to a compact [n_text, H] buffer in
// text_src order (the caller index_adds them into the embedding gradient).
// ---------------------------------------------------------------------------------------------
struct MergeSpliceBwdArgs {
  const uint4* dout;
  int nvec, T, S;
  const radvlm_splice_segment* segments;
  int n_segments;
  const radvlm_merge_image* images;
  int64_t total_rows;
  float* dfeat;     // [tiles*T, H] fp32
  float* dnewline;  // [H] fp32
  uint4* dtext;     // [n_text, H] (dtype of dout) or nullptr
  const uint4* features;  // forward input (max pooling only: the gradient goes to the arg-max of every window)
};

// Max-pooling backward: every channel's gradient goes to the first maximum of its 2 x 2 window in row-major window
// order (ATen max_pool2d_with_indices keeps the first value that is not smaller; NaN wins).
template <typename T>
__device__ __forceinline__ void scatter_row_max(const uint4* __restrict__ src, const uint4* __restrict__ feat,
                                                float* __restrict__ dfeat, const size_t* rows, int nvec, int lane) {
  const size_t H = static_cast<size_t>(nvec) * Vec16<T>::N;
  for (int i = lane; i < nvec; i += 32) {
    float g[Vec16<T>::N], f[4][Vec16<T>::N];
    Vec16<T>::unpack(ld_stream(src + i), g);
#pragma unroll
    for (int k = 0; k < 4; ++k) Vec16<T>::unpack(ld_stream(feat + rows[k] * nvec + i), f[k]);
#pragma unroll
    for (int e = 0; e < Vec16<T>::N; ++e) {
      int best = 0;
      float bv = f[0][e];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (f[k][e] > bv || f[k][e] != f[k][e]) { best = k; bv = f[k][e]; }
      atomicAdd(dfeat + rows[best] * H + static_cast<size_t>(i) * Vec16<T>::N + e, g[e]);
    }
  }
}

template <typename T>
__device__ __forceinline__ void scatter_row(const uint4* __restrict__ src, float* __restrict__ dst, float wgt, int nvec,
                                            int lane) {
  for (int i = lane; i < nvec; i += 32) {
    float f[Vec16<T>::N];
    Vec16<T>::unpack(ld_stream(src + i), f);
#pragma unroll
    for (int e = 0; e < Vec16<T>::N; ++e) atomicAdd(dst + static_cast<size_t>(i) * Vec16<T>::N + e, wgt * f[e]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
merge_splice_bwd_kernel(const MergeSpliceBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const size_t H = static_cast<size_t>(a.nvec) * Vec16<T>::N;
  for (int64_t row = warp0; row < a.total_rows; row += nwarps) {
    int lo = 0, hi = a.n_segments - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(&a.segments[mid].dst_row) <= row) lo = mid; else hi = mid - 1;
    }
    const radvlm_splice_segment seg = a.segments[lo];
    const int off = static_cast<int>(row - seg.dst_row);
    const uint4* src = a.dout + static_cast<size_t>(row) * a.nvec;
    if (seg.kind == RADVLM_SEG_PAD || off >= seg.length) continue;
    if (seg.kind == RADVLM_SEG_TEXT) {
      if (a.dtext != nullptr) copy_row(src, a.dtext + static_cast<size_t>(seg.src_off + off) * a.nvec, a.nvec, lane);
      continue;
    }
    const radvlm_merge_image im = a.images[seg.image];
    const int t = seg.src_off + off;
    if (im.mode == RADVLM_MERGE_VIDEO) {
      const VideoTaps v = video_source(im, t, a.S, a.T);
      if (v.newline) {
        scatter_row<T>(src, a.dnewline, 1.f, a.nvec, lane);
      } else if (im.pool == RADVLM_POOL_MAX) {
        if (a.features != nullptr) scatter_row_max<T>(src, a.features, a.dfeat, v.row, a.nvec, lane);
      } else if (im.pool == RADVLM_POOL_BILINEAR) {
        const float ly0 = 1.f - v.ly1, lx0 = 1.f - v.lx1;
        scatter_row<T>(src, a.dfeat + v.row[0] * H, ly0 * lx0, a.nvec, lane);
        scatter_row<T>(src, a.dfeat + v.row[1] * H, ly0 * v.lx1, a.nvec, lane);
        scatter_row<T>(src, a.dfeat + v.row[2] * H, v.ly1 * lx0, a.nvec, lane);
        scatter_row<T>(src, a.dfeat + v.row[3] * H, v.ly1 * v.lx1, a.nvec, lane);
      } else if (im.pool == RADVLM_POOL_AVERAGE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) scatter_row<T>(src, a.dfeat + v.row[k] * H, 0.25f, a.nvec, lane);
      }
    } else if (im.mode == RADVLM_MERGE_FLAT) {
      scatter_row<T>(src, a.dfeat + (static_cast<size_t>(im.tile_base) * a.T + t) * H, 1.f, a.nvec, lane);
    } else if (im.mode == RADVLM_MERGE_SINGLE) {
      if (t < a.T) scatter_row<T>(src, a.dfeat + (static_cast<size_t>(im.tile_base) * a.T + t) * H, 1.f, a.nvec, lane);
      else scatter_row<T>(src, a.dnewline, 1.f, a.nvec, lane);
    } else if (!(im.reserved & RADVLM_ANYRES_NO_BASE) && t < a.T) {
      scatter_row<T>(src, a.dfeat + (static_cast<size_t>(im.tile_base) * a.T + t) * H, 1.f, a.nvec, lane);
    } else {
      const int u = t - ((im.reserved & RADVLM_ANYRES_NO_BASE) ? 0 : a.T);
      const int row_w = im.out_w + ((im.reserved & RADVLM_ANYRES_NO_NEWLINE) ? 0 : 1);
      const int r = u / row_w;
      const int c = u - r * row_w;
      if (c == im.out_w) {
        scatter_row<T>(src, a.dnewline, 1.f, a.nvec, lane);
      } else if (im.pool == RADVLM_POOL_MAX) {
        if (a.features != nullptr) {
          const size_t rows[4] = {grid_src_row(im, 2 * r, 2 * c, a.S, a.T), grid_src_row(im, 2 * r, 2 * c + 1, a.S, a.T),
                                  grid_src_row(im, 2 * r + 1, 2 * c, a.S, a.T), grid_src_row(im, 2 * r + 1, 2 * c + 1, a.S, a.T)};
          scatter_row_max<T>(src, a.features, a.dfeat, rows, a.nvec, lane);
        }
      } else if (!im.pool) {
        scatter_row<T>(src, a.dfeat + grid_src_row(im, r + im.crop_r0, c + im.crop_c0, a.S, a.T) * H, 1.f, a.nvec, lane);
      } else {
        const float sh = static_cast<float>(im.crop_h) / static_cast<float>(im.out_h);
        const float sw = static_cast<float>(im.crop_w) / static_cast<float>(im.out_w);
        float sy = sh * (static_cast<float>(r) + 0.5f) - 0.5f;
        float sx = sw * (static_cast<float>(c) + 0.5f) - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
        const int y1 = y0 + (y0 < im.crop_h - 1 ? 1 : 0), x1 = x0 + (x0 < im.crop_w - 1 ? 1 : 0);
        const float ly1 = sy - static_cast<float>(y0), lx1 = sx - static_cast<float>(x0);
        const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
        scatter_row<T>(src, a.dfeat + grid_src_row(im, y0 + im.crop_r0, x0 + im.crop_c0, a.S, a.T) * H, ly0 * lx0, a.nvec, lane);
        scatter_row<T>(src, a.dfeat + grid_src_row(im, y0 + im.crop_r0, x1 + im.crop_c0, a.S, a.T) * H, ly0 * lx1, a.nvec, lane);
        scatter_row<T>(src, a.dfeat + grid_src_row(im, y1 + im.crop_r0, x0 + im.crop_c0, a.S, a.T) * H, ly1 * lx0, a.nvec, lane);
        scatter_row<T>(src, a.dfeat + grid_src_row(im, y1 + im.crop_r0, x1 + im.crop_c0, a.S, a.T) * H, ly1 * lx1, a.nvec, lane);
      }
    }
  }
}

}  // namespace rv

namespace rv {

static int merge_splice_launch(const void* features, const void* newline, const void* embed_table, int dtype, int hidden,
                               int tokens_per_tile, int patches_per_side, const int64_t* input_ids,
                               const int64_t* labels_in, const int32_t* text_src, const radvlm_splice_segment* segments,
                               int n_segments, const radvlm_merge_image* images, int n_images, int64_t total_rows,
                               void* out_embeds, void* const* out_peers, int n_peers, int max_ctas, int64_t* out_labels,
                               uint8_t* out_mask, int64_t* out_pos, int64_t ignore_index, void* stream) {
  int st = require_sm100();
  if (st) return st;
  RV_CHECK_ARG(segments && n_segments > 0 && (out_embeds || n_peers > 0) && total_rows > 0, "merge_splice: bad arguments");
  RV_CHECK_ARG(n_images == 0 || (images && features), "merge_splice: image table without features");
  RV_CHECK_ARG(n_peers >= 0 && n_peers <= RADVLM_MAX_PEERS, "merge_splice: %d destinations (at most %d)", n_peers,
               RADVLM_MAX_PEERS);
  const int esize = (dtype == RADVLM_DT_F32) ? 4 : 2;
  RV_CHECK_ARG((static_cast<long long>(hidden) * esize) % 16 == 0, "merge_splice: row bytes must be a multiple of 16");
  MergeSpliceArgs a;
  a.features = static_cast<const uint4*>(features);
  a.newline = static_cast<const uint4*>(newline);
  a.embed = static_cast<const uint4*>(embed_table);
  a.nvec = hidden * esize / 16;
  a.T = tokens_per_tile;
  a.S = patches_per_side;
  a.input_ids = input_ids;
  a.labels_in = labels_in;
  a.text_src = text_src;
  a.segments = segments;
  a.n_segments = n_segments;
  a.images = images;
  a.total_rows = total_rows;
  a.out = static_cast<uint4*>(out_embeds);
  a.n_peers = n_peers;
  for (int d = 0; d < RADVLM_MAX_PEERS; ++d) {
    a.out_peers[d] = d < n_peers ? static_cast<uint4*>(out_peers[d]) : nullptr;
    RV_CHECK_ARG(d >= n_peers || out_peers[d] != nullptr, "merge_splice: destination %d is null", d);
  }
  a.out_labels = out_labels;
  a.out_mask = out_mask;
  a.out_pos = out_pos;
  a.ignore_index = ignore_index;
  const int threads = 256;
  const int64_t want = (total_rows * 32 + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;  // 8 resident CTAs per SM
  if (max_ctas > 0 && max_ctas < cap) cap = max_ctas;
  const int blocks = static_cast<int>(want < cap ? want : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope ps(PROF_MERGE_SPLICE, s);
  const bool scatter = n_peers > 0;
#define RV_MS_LAUNCH(TYPE)                                                            \
  do {                                                                                \
    if (scatter) merge_splice_kernel<TYPE, true><<<blocks, threads, 0, s>>>(a);       \
    else merge_splice_kernel<TYPE, false><<<blocks, threads, 0, s>>>(a);              \
  } while (0)
  switch (dtype) {
    case RADVLM_DT_F32: RV_MS_LAUNCH(float); break;
    case RADVLM_DT_BF16: RV_MS_LAUNCH(__nv_bfloat16); break;
    case RADVLM_DT_F16: RV_MS_LAUNCH(__half); break;
    default: set_error("merge_splice: unknown dtype %d", dtype); return RADVLM_ERR_BAD_ARGUMENT;
  }
#undef RV_MS_LAUNCH
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// Cross-rank step barrier over peer memory: rank `rank` publishes `value` into slot [rank] of every rank's flag
// array (system-scope release after this stream's earlier work, i.e. after its scattered rows), then waits until
// all `n` slots of its own array have reached `value`.  The wait is bounded by wall time (%globaltimer, independent
// of the SM clock); timeout_ns <= 0 waits for ever.  A peer that does not arrive in time does NOT kill the context:
// the kernel records 1 + (first missing rank) in *status (host-pinned or device memory, may be null) and returns, and
// the host raises when it next looks (dist.PeerGather.check).
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_signal_wait_kernel(unsigned long long* const* flags_peers, unsigned long long* flags_local, int n,
                                        int rank, unsigned long long value, long long timeout_ns, int* status) {
  const int d = threadIdx.x;
  if (d >= n) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags_peers[d] + rank), "l"(value) : "memory");
  const unsigned long long t0 = global_timer_ns();
  unsigned long long seen = 0;
  unsigned spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags_local + d) : "memory");
    if (seen >= value) break;
    if ((++spins & 255u) == 0 && timeout_ns > 0 &&
        global_timer_ns() - t0 > static_cast<unsigned long long>(timeout_ns)) {
      if (status != nullptr) {
        atomicCAS(status, 0, 1 + d);   // first missing rank wins; a nonzero status is sticky until the host clears it
        __threadfence_system();
      }
      break;
    }
    __nanosleep(200);
  } while (true);
}

}  // namespace rv

extern "C" int radvlm_merge_splice(const void* features, const void* newline, const void* embed_table,
                                   int dtype, int hidden, int tokens_per_tile, int patches_per_side,
                                   const int64_t* input_ids, const int64_t* labels_in,
                                   const int32_t* text_src, const radvlm_splice_segment* segments,
                                   int n_segments, const radvlm_merge_image* images, int n_images,
                                   int64_t total_rows, void* out_embeds, int64_t* out_labels,
                                   uint8_t* out_mask, int64_t* out_pos, int64_t ignore_index,
                                   void* stream) {
  return rv::merge_splice_launch(features, newline, embed_table, dtype, hidden, tokens_per_tile, patches_per_side,
                                 input_ids, labels_in, text_src, segments, n_segments, images, n_images, total_rows,
                                 out_embeds, nullptr, 0, 0, out_labels, out_mask, out_pos, ignore_index, stream);
}

extern "C" int radvlm_merge_splice_scatter(const void* features, const void* newline, const void* embed_table,
                                           int dtype, int hidden, int tokens_per_tile, int patches_per_side,
                                           const int64_t* input_ids, const int64_t* labels_in,
                                           const int32_t* text_src, const radvlm_splice_segment* segments,
                                           int n_segments, const radvlm_merge_image* images, int n_images,
                                           int64_t total_rows, void* const* out_peers, int n_peers, int max_ctas,
                                           int64_t* out_labels, uint8_t* out_mask, int64_t* out_pos,
                                           int64_t ignore_index, void* stream) {
  using namespace rv;
  RV_CHECK_ARG(out_peers != nullptr && n_peers >= 1, "merge_splice_scatter: no destinations");
  return merge_splice_launch(features, newline, embed_table, dtype, hidden, tokens_per_tile, patches_per_side, input_ids,
                             labels_in, text_src, segments, n_segments, images, n_images, total_rows, nullptr, out_peers,
                             n_peers, max_ctas, out_labels, out_mask, out_pos, ignore_index, stream);
}

// ---- peer memory (one process per GPU, same node): cudaIpc handles exchanged by the host through torch.distributed
extern "C" int radvlm_peer_alloc(size_t bytes, void** ptr, uint8_t* handle64) {
  using namespace rv;
  RV_CHECK_ARG(bytes > 0 && ptr && handle64, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  RV_CUDA(cudaMalloc(&p, bytes));
  RV_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return RADVLM_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return RADVLM_OK;
}

extern "C" int radvlm_peer_open(const uint8_t* handle64, void** ptr) {
  using namespace rv;
  RV_CHECK_ARG(handle64 && ptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  RV_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RADVLM_OK;
}

extern "C" int radvlm_peer_close(void* ptr) {
  using namespace rv;
  if (ptr) RV_CUDA(cudaIpcCloseMemHandle(ptr));
  return RADVLM_OK;
}

extern "C" int radvlm_peer_free(void* ptr) {
  using namespace rv;
  if (ptr) RV_CUDA(cudaFree(ptr));
  return RADVLM_OK;
}

extern "C" int radvlm_peer_signal_wait(void* const* flags_peers_dev, void* flags_local, int n, int rank,
                                       unsigned long long value, double timeout_s, int* status, void* stream) {
  using namespace rv;
  int st = require_sm100();
  if (st) return st;
  RV_CHECK_ARG(flags_peers_dev && flags_local && n >= 1 && n <= RADVLM_MAX_PEERS && rank >= 0 && rank < n,
               "peer_signal_wait: bad arguments");
  const long long timeout_ns = timeout_s > 0 ? static_cast<long long>(timeout_s * 1e9) : 0;
  peer_signal_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<unsigned long long* const*>(flags_peers_dev), static_cast<unsigned long long*>(flags_local), n,
      rank, value, timeout_ns, status);
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}

// Copy-engine form of the exchange: this rank's finished [rows, H] slice is pushed into a peer's gathered buffer by a
// DMA engine over NVLink (no SM is used; the pointers are UVA, the driver routes by their owning devices).
extern "C" int radvlm_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  using namespace rv;
  RV_CHECK_ARG(dst && src && bytes > 0, "peer_copy: bad arguments");
  RV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return RADVLM_OK;
}

extern "C" int radvlm_merge_splice_backward(const void* d_out_embeds, int dtype, int hidden, int tokens_per_tile,
                                            int patches_per_side, const radvlm_splice_segment* segments, int n_segments,
                                            const radvlm_merge_image* images, int n_images, int64_t total_rows,
                                            float* d_features, float* d_newline, void* d_text, const void* features,
                                            void* stream) {
  using namespace rv;
  int st = require_sm100();
  if (st) return st;
  RV_CHECK_ARG(d_out_embeds && segments && n_segments > 0 && total_rows > 0, "merge_splice_backward: bad arguments");
  RV_CHECK_ARG(n_images == 0 || (images && d_features && d_newline), "merge_splice_backward: image table without gradients");
  const int esize = (dtype == RADVLM_DT_F32) ? 4 : 2;
  RV_CHECK_ARG((static_cast<long long>(hidden) * esize) % 16 == 0, "merge_splice_backward: row bytes must be a multiple of 16");
  MergeSpliceBwdArgs a;
  a.dout = static_cast<const uint4*>(d_out_embeds);
  a.nvec = hidden * esize / 16;
  a.T = tokens_per_tile;
  a.S = patches_per_side;
  a.segments = segments;
  a.n_segments = n_segments;
  a.images = images;
  a.total_rows = total_rows;
  a.dfeat = d_features;
  a.dnewline = d_newline;
  a.dtext = static_cast<uint4*>(d_text);
  a.features = static_cast<const uint4*>(features);
  const int threads = 256;
  const int64_t want = (total_rows * 32 + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  const int blocks = static_cast<int>(want < cap ? want : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case RADVLM_DT_F32: merge_splice_bwd_kernel<float><<<blocks, threads, 0, s>>>(a); break;
    case RADVLM_DT_BF16: merge_splice_bwd_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(a); break;
    case RADVLM_DT_F16: merge_splice_bwd_kernel<__half><<<blocks, threads, 0, s>>>(a); break;
    default: set_error("merge_splice_backward: unknown dtype %d", dtype); return RADVLM_ERR_BAD_ARGUMENT;
  }
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}
