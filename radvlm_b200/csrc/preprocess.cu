// Fused anyres preprocessing on the GPU (HBM-bound, integer arithmetic, bit-exact with Pillow):
//   uint8 HWC image -> aspect-preserving BICUBIC resize -> centred paste on a black canvas ->
//   384x384 crops (row-major) + the aspect-distorting 384x384 base tile -> rescale(1/255) ->
//   normalize((x-0.5)/0.5) -> CHW tiles [1+gw*gh, 3, 384, 384] in fp32 / bf16 / fp16.
//
// Replaces (reference, all on CPU in DataLoader workers):
//   process_anyres_image / resize_and_pad_image / divide_to_patches   mm_utils.py:152-210,243-293
//   SigLipImageProcessor.preprocess                                   siglip_encoder.py:47-67
// Third-party arithmetic restated: Pillow's 8bpc ImagingResample (pinned 10.3.0, verified against the
// 12.2.0 in this image): separable, horizontal pass first into a uint8 intermediate, 22-bit fixed-point
// coefficients from a float64 bicubic (a=-0.5) kernel of support 2*max(1, in/out), accumulator seeded
// with 1<<21, arithmetic >>22, clip to [0,255].  transformers.image_transforms.rescale/normalize:
// float32(float64(u8) * (1/255)), then (x - 0.5f) / 0.5f in fp32.
//
// Kernels (all batched over images through a device-resident descriptor table):
//   1. resample_coeffs_kernel : per image, 4 coefficient tables (x/y for the canvas resize, x/y for the base tile)
//   2. resample_h_kernel      : horizontal pass -> uint8 intermediates (canvas and base)
//   3. resample_v_tiles_kernel: vertical pass + paste + crop + normalise + CHW scatter
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "host_util.h"

namespace rv {

constexpr int kPrecisionBits = 22;
constexpr int kMaxKsize = 129;  // supports down-scales up to 32x

struct AxisLayout {
  int ksize;
  size_t bounds_off;  // int32 [2*out]
  size_t kk_off;      // int32 [out*ksize]
};

struct PreLayout {
  AxisLayout xm, ym, xb, yb;  // main (canvas) x / y, base tile x / y
  size_t tmp_main_off;        // uint8 [H, nw, C]
  size_t tmp_base_off;        // uint8 [H, S, C]
  size_t total;
};

__host__ __device__ inline int resample_ksize(int in, int out) {
  double fs = static_cast<double>(in) / static_cast<double>(out);
  if (fs < 1.0) fs = 1.0;
  const double support = 2.0 * fs;
  return static_cast<int>(ceil(support)) * 2 + 1;
}

__host__ __device__ inline size_t al16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }

__host__ __device__ inline PreLayout make_pre_layout(int W, int H, int C, int nw, int nh, int S) {
  PreLayout L;
  size_t off = 0;
  auto axis = [&](AxisLayout& a, int in, int out) {
    a.ksize = resample_ksize(in, out);
    a.bounds_off = off;
    off = al16(off + static_cast<size_t>(out) * 2 * 4);
    a.kk_off = off;
    off = al16(off + static_cast<size_t>(out) * a.ksize * 4);
  };
  axis(L.xm, W, nw);
  axis(L.ym, H, nh);
  axis(L.xb, W, S);
  axis(L.yb, H, S);
  L.tmp_main_off = off;
  off = al16(off + static_cast<size_t>(H) * nw * C);
  L.tmp_base_off = off;
  off = al16(off + static_cast<size_t>(H) * S * C);
  L.total = off;
  return L;
}

// ---------------------------------------------------------------------------------------------
// 1. coefficient tables (Pillow precompute_coeffs + normalize_coeffs_8bpc), fp64 without contraction
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double bicubic_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) {
    // ((a + 2.0) * x - (a + 3.0)) * x * x + 1, a = -0.5
    double t = __dsub_rn(__dmul_rn(1.5, x), 2.5);
    t = __dmul_rn(t, x);
    t = __dmul_rn(t, x);
    return __dadd_rn(t, 1.0);
  }
  if (x < 2.0) {
    // (((x - 5) * x + 8) * x - 4) * a
    double t = __dsub_rn(x, 5.0);
    t = __dadd_rn(__dmul_rn(t, x), 8.0);
    t = __dsub_rn(__dmul_rn(t, x), 4.0);
    return __dmul_rn(t, -0.5);
  }
  return 0.0;
}

__device__ void compute_axis_coeffs(int in, int out, int ksize, int xx, int32_t* bounds, int32_t* kk) {
  const double scale = __ddiv_rn(static_cast<double>(in), static_cast<double>(out));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);
  const double ss = __ddiv_rn(1.0, filterscale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in) xmax = in;
  xmax -= xmin;
  double w[kMaxKsize];
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    w[x] = bicubic_filter(arg);
    ww = __dadd_rn(ww, w[x]);
  }
  int32_t* k = kk + static_cast<size_t>(xx) * ksize;
  for (int x = 0; x < ksize; ++x) {
    int32_t c = 0;
    if (x < xmax) {
      double v = w[x];
      if (ww != 0.0) v = __ddiv_rn(v, ww);
      const double s = __dmul_rn(v, static_cast<double>(1 << kPrecisionBits));
      c = (v < 0.0) ? static_cast<int32_t>(__dadd_rn(-0.5, s)) : static_cast<int32_t>(__dadd_rn(0.5, s));
    }
    k[x] = c;
  }
  bounds[2 * xx] = xmin;
  bounds[2 * xx + 1] = xmax;
}

__global__ void __launch_bounds__(64)
resample_coeffs_kernel(const radvlm_preprocess_image* __restrict__ imgs, uint8_t* __restrict__ scratch,
                       int S) {
  const radvlm_preprocess_image im = imgs[blockIdx.z];
  const PreLayout L = make_pre_layout(im.width, im.height, im.channels, im.resized_w, im.resized_h, S);
  const int axis = blockIdx.y;
  int in, out;
  AxisLayout A;
  switch (axis) {
    case 0: in = im.width; out = im.resized_w; A = L.xm; break;
    case 1: in = im.height; out = im.resized_h; A = L.ym; break;
    case 2: in = im.width; out = S; A = L.xb; break;
    default: in = im.height; out = S; A = L.yb; break;
  }
  uint8_t* base = scratch + im.scratch_offset;
  for (int xx = blockIdx.x * blockDim.x + threadIdx.x; xx < out; xx += gridDim.x * blockDim.x)
    compute_axis_coeffs(in, out, A.ksize, xx, reinterpret_cast<int32_t*>(base + A.bounds_off),
                        reinterpret_cast<int32_t*>(base + A.kk_off));
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ---------------------------------------------------------------------------------------------
// 2+3 fused: one CTA produces a block of the output plane (canvas tiles: 16 rows x 128 columns, base tile: 32 x 64;
// smaller blocks when a batch holds large down-scales, e.g. 2544x3056 MIMIC images, so the source rectangle fits):
//   source rectangle -> shared memory with 16-byte loads -> horizontal pass -> uint8 intermediate (planar, shared
//   memory; the same rounding to uint8 Pillow does between its passes) -> vertical pass -> LUT normalise -> CHW
//   tile rows with 16-byte stores.  Same coefficient tables and the same integer arithmetic as the two-kernel
//   path below, which stays as the fall-back for geometries whose source rectangle does not fit shared memory even
//   with 8 x 16 blocks (down-scales beyond ~20x) or tile sizes that are not a multiple of 32.
// ---------------------------------------------------------------------------------------------
constexpr int kFuBX = 128, kFuThreads = 256;  // kFuBX: widest block = row pitch of the intermediate planes

struct FusedSmem {   // identical on host (sizing) and device (offsets)
  int max_nr;        // source / intermediate rows held per block
  int src_pitch;     // bytes per staged source row (16-byte multiple)
  int max_ksx;       // horizontal taps
  int cbx, cby;      // output block of the canvas target (columns, rows)
  int bbx, bby;      // output block of the base-tile target (its down-scale is larger: smaller blocks)
};
__host__ __device__ inline size_t fused_src_off() { return 0; }
__host__ __device__ inline size_t fused_inter_off(const FusedSmem& f) { return al16(static_cast<size_t>(f.max_nr) * f.src_pitch); }
__host__ __device__ inline size_t fused_kx_off(const FusedSmem& f) { return fused_inter_off(f) + al16(static_cast<size_t>(3) * f.max_nr * kFuBX); }
__host__ __device__ inline size_t fused_bx_off(const FusedSmem& f) { return fused_kx_off(f) + al16(static_cast<size_t>(kFuBX) * f.max_ksx * 4); }
__host__ __device__ inline size_t fused_shift_off(const FusedSmem& f) { return fused_bx_off(f) + 2 * kFuBX * 4; }
__host__ __device__ inline size_t fused_lut_off(const FusedSmem& f) { return fused_shift_off(f) + al16(static_cast<size_t>(f.max_nr) * 4); }
__host__ __device__ inline size_t fused_total(const FusedSmem& f) { return fused_lut_off(f) + 256 * 4; }

// rows / columns of the input an output range of n samples can touch (scale = in / out, ksize taps per sample)
__host__ __device__ inline int fused_span(int in, int out, int n) {
  if (in == out) return n;
  const double scale = static_cast<double>(in) / static_cast<double>(out);
  return static_cast<int>(ceil((n - 1) * scale)) + resample_ksize(in, out) + 1;
}

template <typename T>
__device__ __forceinline__ void store8(T* dst, const float* v);
template <>
__device__ __forceinline__ void store8<float>(float* dst, const float* v) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* dst, const float* v) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                              pack_bf16x2(v[6], v[7]));
}
template <>
__device__ __forceinline__ void store8<__half>(__half* dst, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename T>
__global__ void __launch_bounds__(kFuThreads)
resample_fused_kernel(const radvlm_preprocess_image* __restrict__ imgs, const uint8_t* __restrict__ src,
                      const uint8_t* __restrict__ scratch, T* __restrict__ tiles, int S, FusedSmem fs) {
  extern __shared__ __align__(16) uint8_t fsm[];
  uint8_t* src_s = fsm + fused_src_off();
  uint8_t* inter_s = fsm + fused_inter_off(fs);
  int32_t* kx_s = reinterpret_cast<int32_t*>(fsm + fused_kx_off(fs));
  int32_t* bx_s = reinterpret_cast<int32_t*>(fsm + fused_bx_off(fs));
  int32_t* shift_s = reinterpret_cast<int32_t*>(fsm + fused_shift_off(fs));
  float* lut = reinterpret_cast<float*>(fsm + fused_lut_off(fs));
  const int tid = threadIdx.x;

  const radvlm_preprocess_image im = imgs[blockIdx.z];
  const int canvas_w = im.grid_w * S, canvas_h = im.grid_h * S;
  const int nb_canvas = canvas_h / fs.cby;        // S % cby == 0: a block never straddles canvas and base rows
  const bool is_base = (static_cast<int>(blockIdx.y) >= nb_canvas);
  const int BX = is_base ? fs.bbx : fs.cbx, BY = is_base ? fs.bby : fs.cby;
  const int Y0 = is_base ? canvas_h + (static_cast<int>(blockIdx.y) - nb_canvas) * BY : static_cast<int>(blockIdx.y) * BY;
  const int X0 = blockIdx.x * BX;
  if (Y0 >= canvas_h + S) return;
  const int plane_w = is_base ? S : canvas_w;
  if (X0 >= plane_w) return;

  // rescale: float32(float64(u8) * (1/255));  normalize: (x - 0.5f) / 0.5f
  if (tid < 256) {
    const float r = __double2float_rn(__dmul_rn(static_cast<double>(tid), 1.0 / 255.0));
    lut[tid] = __fdiv_rn(__fsub_rn(r, 0.5f), 0.5f);
  }

  const PreLayout L = make_pre_layout(im.width, im.height, im.channels, im.resized_w, im.resized_h, S);
  const uint8_t* base = scratch + im.scratch_offset;
  const int C = im.channels, W = im.width, H = im.height;
  const int out_w = is_base ? S : im.resized_w, out_h = is_base ? S : im.resized_h;
  const int px = is_base ? 0 : im.paste_x, py = is_base ? canvas_h : im.paste_y;  // output -> resized coordinates
  const AxisLayout Ax = is_base ? L.xb : L.xm, Ay = is_base ? L.yb : L.ym;
  const int32_t* bxg = reinterpret_cast<const int32_t*>(base + Ax.bounds_off);
  const int32_t* kxg = reinterpret_cast<const int32_t*>(base + Ax.kk_off);
  const int32_t* byg = reinterpret_cast<const int32_t*>(base + Ay.bounds_off);
  const int32_t* kyg = reinterpret_cast<const int32_t*>(base + Ay.kk_off);
  const bool h_skip = (out_w == W), v_skip = (out_h == H);

  // the part of this block that lies inside the pasted (resized) image
  const int oy_lo = max(Y0 - py, 0), oy_hi = min(Y0 - py + BY, out_h);
  const int ox_lo = max(X0 - px, 0), ox_hi = min(X0 - px + BX, out_w);
  const bool empty = (oy_lo >= oy_hi) || (ox_lo >= ox_hi);
  int r_lo = 0, nr = 0, c_lo = 0, nc = 0;
  const int nox = empty ? 0 : ox_hi - ox_lo;
  const int col0 = ox_lo - (X0 - px);  // intermediate column of ox_lo (columns are indexed by X - X0)
  if (!empty) {
    if (v_skip) { r_lo = oy_lo; nr = oy_hi - oy_lo; }
    else { r_lo = byg[2 * oy_lo]; nr = byg[2 * (oy_hi - 1)] + byg[2 * (oy_hi - 1) + 1] - r_lo; }
    if (h_skip) { c_lo = ox_lo; nc = nox; }
    else { c_lo = bxg[2 * ox_lo]; nc = bxg[2 * (ox_hi - 1)] + bxg[2 * (ox_hi - 1) + 1] - c_lo; }
    if (nr > fs.max_nr || nc * C + 32 > fs.src_pitch || Ax.ksize > fs.max_ksx) __trap();  // host sizing bug

    // ---- phase 0: horizontal coefficients of this block's columns + the source rectangle (16-byte loads)
    if (!h_skip) {
      for (int i = tid; i < nox * Ax.ksize; i += kFuThreads) kx_s[i] = kxg[static_cast<size_t>(ox_lo) * Ax.ksize + i];
      for (int i = tid; i < 2 * nox; i += kFuThreads) bx_s[i] = bxg[2 * ox_lo + i];
    }
    const uint8_t* img = src + im.src_offset;
    const uint8_t* img_end = img + static_cast<size_t>(W) * H * C;
    const int chunks = fs.src_pitch / 16;
    for (int i = tid; i < nr * chunks; i += kFuThreads) {
      const int r = i / chunks, ch = i - r * chunks;
      const uint8_t* row = img + (static_cast<size_t>(r_lo + r) * W + c_lo) * C;
      const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(row) & 15u);
      if (ch == 0) shift_s[r] = static_cast<int>(a);
      if (ch * 16 >= static_cast<int>(a) + nc * C) continue;  // past the row's bytes
      const uint8_t* g = row - a + ch * 16;
      uint4 v;
      if (g >= img && g + 16 <= img_end) {
        v = *reinterpret_cast<const uint4*>(g);
      } else {  // first / last bytes of the image: stay inside it
        uint8_t b[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) b[e] = (g + e >= img && g + e < img_end) ? g[e] : 0;
        v = *reinterpret_cast<const uint4*>(b);
      }
      *reinterpret_cast<uint4*>(src_s + static_cast<size_t>(r) * fs.src_pitch + ch * 16) = v;
    }
  }
  __syncthreads();

  // ---- phase 1: horizontal pass -> planar uint8 intermediate [channel][row][X - X0]
  if (!empty) {
    for (int i = tid; i < nr * nox; i += kFuThreads) {
      const int r = i / nox, o = i - r * nox;
      const uint8_t* rowp = src_s + static_cast<size_t>(r) * fs.src_pitch + shift_s[r];
      uint8_t* dst = inter_s + static_cast<size_t>(r) * kFuBX + col0 + o;
      const size_t plane = static_cast<size_t>(fs.max_nr) * kFuBX;
      if (h_skip) {
        const uint8_t* p = rowp + o * C;
        dst[0] = p[0];
        if (C == 3) { dst[plane] = p[1]; dst[2 * plane] = p[2]; }
      } else {
        const int xmin = bx_s[2 * o] - c_lo, cnt = bx_s[2 * o + 1];
        const int32_t* k = kx_s + o * Ax.ksize;
        if (C == 1) {
          int s0 = 1 << (kPrecisionBits - 1);
          for (int x = 0; x < cnt; ++x) s0 += static_cast<int>(rowp[xmin + x]) * k[x];
          dst[0] = clip8(s0);
        } else {
          int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
          const uint8_t* p = rowp + xmin * 3;
          for (int x = 0; x < cnt; ++x) {
            const int kv = k[x];
            s0 += static_cast<int>(p[3 * x]) * kv;
            s1 += static_cast<int>(p[3 * x + 1]) * kv;
            s2 += static_cast<int>(p[3 * x + 2]) * kv;
          }
          dst[0] = clip8(s0);
          dst[plane] = clip8(s1);
          dst[2 * plane] = clip8(s2);
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: vertical pass + paste (black outside) + normalise + CHW tile rows, 8 pixels per thread
  const size_t plane = static_cast<size_t>(fs.max_nr) * kFuBX;
  const int groups = BX / 8;
  for (int i = tid; i < BY * 3 * groups; i += kFuThreads) {
    const int g8 = i % groups;
    const int ch = (i / groups) % 3;
    const int y = i / (3 * groups);
    const int Y = Y0 + y, X = X0 + g8 * 8;
    if (X >= plane_w || Y >= canvas_h + S) continue;  // plane widths are multiples of 8
    const int oy = Y - py;
    int v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0;
    if (!empty && oy >= oy_lo && oy < oy_hi) {
      const uint8_t* pl = inter_s + (C == 1 ? 0 : ch) * plane + g8 * 8;
      if (v_skip) {
        const uint2 w = *reinterpret_cast<const uint2*>(pl + static_cast<size_t>(oy - r_lo) * kFuBX);
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[e] = (w.x >> (8 * e)) & 0xFF; v[4 + e] = (w.y >> (8 * e)) & 0xFF; }
      } else {
        const int ymin = byg[2 * oy] - r_lo, cnt = byg[2 * oy + 1];
        const int32_t* k = kyg + static_cast<size_t>(oy) * Ay.ksize;
        int acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 1 << (kPrecisionBits - 1);
        for (int t = 0; t < cnt; ++t) {
          const uint2 w = *reinterpret_cast<const uint2*>(pl + static_cast<size_t>(ymin + t) * kFuBX);
          const int kv = __ldg(k + t);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[e] += static_cast<int>((w.x >> (8 * e)) & 0xFF) * kv;
            acc[4 + e] += static_cast<int>((w.y >> (8 * e)) & 0xFF) * kv;
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = clip8(acc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {  // columns outside the pasted image are black
        const int ox = X + e - px;
        if (ox < ox_lo || ox >= ox_hi) v[e] = 0;
      }
    }
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = lut[v[e]];
    int tile, ty, tx;
    if (is_base) {
      tile = im.tile_base; ty = Y - canvas_h; tx = X;
    } else {
      const int gy = Y / S, gx = X / S;
      tile = im.tile_base + 1 + gy * im.grid_w + gx; ty = Y - gy * S; tx = X - gx * S;
    }
    store8<T>(tiles + (static_cast<size_t>(tile) * 3 + ch) * S * S + static_cast<size_t>(ty) * S + tx, f);
  }
}

// ---------------------------------------------------------------------------------------------
// 2. horizontal pass: src [H, W, C] -> tmp [H, out_w, C]   (z = image*2 + target; target 0 canvas, 1 base)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
resample_h_kernel(const radvlm_preprocess_image* __restrict__ imgs, const uint8_t* __restrict__ src,
                  uint8_t* __restrict__ scratch, int S) {
  const radvlm_preprocess_image im = imgs[blockIdx.z >> 1];
  const int target = blockIdx.z & 1;
  const int out_w = target ? S : im.resized_w;
  const int y = blockIdx.y;
  if (y >= im.height) return;
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_w) return;
  if (out_w == im.width) return;  // pass skipped (Pillow copies); the vertical pass reads the source
  const PreLayout L = make_pre_layout(im.width, im.height, im.channels, im.resized_w, im.resized_h, S);
  const AxisLayout A = target ? L.xb : L.xm;
  uint8_t* base = scratch + im.scratch_offset;
  const int32_t* bounds = reinterpret_cast<const int32_t*>(base + A.bounds_off);
  const int32_t* k = reinterpret_cast<const int32_t*>(base + A.kk_off) + static_cast<size_t>(xx) * A.ksize;
  const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
  const int C = im.channels;
  const uint8_t* row = src + im.src_offset + static_cast<size_t>(y) * im.width * C;
  uint8_t* out = base + (target ? L.tmp_base_off : L.tmp_main_off) +
                 (static_cast<size_t>(y) * out_w + xx) * C;
  if (C == 1) {
    int s0 = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < xmax; ++x) s0 += static_cast<int>(row[x + xmin]) * k[x];
    out[0] = clip8(s0);
  } else {
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
      const uint8_t* p = row + static_cast<size_t>(x + xmin) * 3;
      const int kv = k[x];
      s0 += static_cast<int>(p[0]) * kv;
      s1 += static_cast<int>(p[1]) * kv;
      s2 += static_cast<int>(p[2]) * kv;
    }
    out[0] = clip8(s0);
    out[1] = clip8(s1);
    out[2] = clip8(s2);
  }
}

// ---------------------------------------------------------------------------------------------
// 3. vertical pass + paste + crop + normalise + CHW tile scatter
//    grid: x = 128-pixel column chunks, y = canvas row (rows >= gh*S address the base tile), z = image
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void __launch_bounds__(128)
resample_v_tiles_kernel(const radvlm_preprocess_image* __restrict__ imgs, const uint8_t* __restrict__ src,
                        const uint8_t* __restrict__ scratch, T* __restrict__ tiles, int S) {
  __shared__ float lut[256];
  {
    // rescale: float32(float64(u8) * (1/255));  normalize: (x - 0.5f) / 0.5f
    for (int u = threadIdx.x; u < 256; u += blockDim.x) {
      const float r = __double2float_rn(__dmul_rn(static_cast<double>(u), 1.0 / 255.0));
      lut[u] = __fdiv_rn(__fsub_rn(r, 0.5f), 0.5f);
    }
  }
  __syncthreads();
  const radvlm_preprocess_image im = imgs[blockIdx.z];
  const int canvas_w = im.grid_w * S, canvas_h = im.grid_h * S;
  const int Y = blockIdx.y;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const bool is_base = (Y >= canvas_h);
  if (Y >= canvas_h + S) return;
  if (X >= (is_base ? S : canvas_w)) return;

  const PreLayout L = make_pre_layout(im.width, im.height, im.channels, im.resized_w, im.resized_h, S);
  const uint8_t* base = scratch + im.scratch_offset;
  const int C = im.channels;

  int tile, ty, tx;      // destination tile and in-tile coordinates
  int oy, ox;            // coordinates inside the resized image
  int out_w, out_h;      // resized image size of this target
  AxisLayout A;
  const uint8_t* hsrc;   // horizontally resampled intermediate (or the source when that pass is skipped)
  bool inside = true;
  if (is_base) {
    tile = im.tile_base;
    ty = Y - canvas_h;
    tx = X;
    oy = ty;
    ox = tx;
    out_w = S;
    out_h = S;
    A = L.yb;
    hsrc = (S == im.width) ? (src + im.src_offset) : (base + L.tmp_base_off);
  } else {
    const int gy = Y / S, gx = X / S;
    tile = im.tile_base + 1 + gy * im.grid_w + gx;
    ty = Y - gy * S;
    tx = X - gx * S;
    oy = Y - im.paste_y;
    ox = X - im.paste_x;
    out_w = im.resized_w;
    out_h = im.resized_h;
    inside = (oy >= 0 && oy < out_h && ox >= 0 && ox < out_w);
    A = L.ym;
    hsrc = (im.resized_w == im.width) ? (src + im.src_offset) : (base + L.tmp_main_off);
  }

  int v0 = 0, v1 = 0, v2 = 0;  // black canvas outside the pasted image
  if (inside) {
    if (out_h == im.height) {  // vertical pass skipped
      const uint8_t* p = hsrc + (static_cast<size_t>(oy) * out_w + ox) * C;
      v0 = p[0];
      if (C == 3) { v1 = p[1]; v2 = p[2]; }
    } else {
      const int32_t* bounds = reinterpret_cast<const int32_t*>(base + A.bounds_off);
      const int32_t* k = reinterpret_cast<const int32_t*>(base + A.kk_off) + static_cast<size_t>(oy) * A.ksize;
      const int ymin = bounds[2 * oy], ymax = bounds[2 * oy + 1];
      int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
      const size_t pitch = static_cast<size_t>(out_w) * C;
      const uint8_t* p = hsrc + static_cast<size_t>(ymin) * pitch + static_cast<size_t>(ox) * C;
      if (C == 1) {
        for (int y = 0; y < ymax; ++y) s0 += static_cast<int>(p[y * pitch]) * k[y];
        v0 = clip8(s0);
      } else {
        for (int y = 0; y < ymax; ++y) {
          const uint8_t* q = p + y * pitch;
          const int kv = k[y];
          s0 += static_cast<int>(q[0]) * kv;
          s1 += static_cast<int>(q[1]) * kv;
          s2 += static_cast<int>(q[2]) * kv;
        }
        v0 = clip8(s0);
        v1 = clip8(s1);
        v2 = clip8(s2);
      }
    }
    if (C == 1) { v1 = v0; v2 = v0; }
  }
  const size_t plane = static_cast<size_t>(S) * S;
  T* o = tiles + static_cast<size_t>(tile) * 3 * plane + static_cast<size_t>(ty) * S + tx;
  o[0] = cvt_out<T>(lut[v0]);
  o[plane] = cvt_out<T>(lut[v1]);
  o[2 * plane] = cvt_out<T>(lut[v2]);
}

}  // namespace rv

extern "C" size_t radvlm_preprocess_scratch_bytes(int width, int height, int channels, int resized_w,
                                                  int resized_h, int tile_size) {
  if (width <= 0 || height <= 0 || resized_w <= 0 || resized_h <= 0 || tile_size <= 0) return 0;
  return rv::make_pre_layout(width, height, channels, resized_w, resized_h, tile_size).total;
}

extern "C" int radvlm_preprocess_anyres(const uint8_t* src, const radvlm_preprocess_image* images_dev,
                                        const radvlm_preprocess_image* images_host, int n_images,
                                        int tile_size, void* tiles_out, int out_dtype, void* scratch,
                                        size_t scratch_bytes, void* stream) {
  using namespace rv;
  int st = require_sm100();
  if (st) return st;
  RV_CHECK_ARG(src && images_dev && images_host && n_images > 0 && tiles_out && scratch,
               "preprocess: bad arguments");
  int max_out = tile_size, max_h = 0, max_cw = tile_size, max_rows = 0;
  size_t need = 0;
  for (int i = 0; i < n_images; ++i) {
    const radvlm_preprocess_image& im = images_host[i];
    RV_CHECK_ARG(im.channels == 1 || im.channels == 3, "preprocess: image %d has %d channels (1 or 3)", i,
                 im.channels);
    RV_CHECK_ARG(im.width > 0 && im.height > 0 && im.grid_w > 0 && im.grid_h > 0 && im.resized_w > 0 &&
                     im.resized_h > 0,
                 "preprocess: image %d has a bad plan", i);
    const PreLayout L = make_pre_layout(im.width, im.height, im.channels, im.resized_w, im.resized_h, tile_size);
    if (L.xm.ksize > kMaxKsize || L.ym.ksize > kMaxKsize || L.xb.ksize > kMaxKsize || L.yb.ksize > kMaxKsize) {
      set_error("preprocess: image %d (%dx%d) needs a down-scale beyond the supported 32x", i, im.width, im.height);
      return RADVLM_ERR_UNSUPPORTED_SHAPE;
    }
    if (im.scratch_offset + L.total > need) need = im.scratch_offset + L.total;
    if (im.resized_w > max_out) max_out = im.resized_w;
    if (im.resized_h > max_out) max_out = im.resized_h;
    if (im.height > max_h) max_h = im.height;
    if (im.grid_w * tile_size > max_cw) max_cw = im.grid_w * tile_size;
    if ((im.grid_h + 1) * tile_size > max_rows) max_rows = (im.grid_h + 1) * tile_size;
  }
  if (need > scratch_bytes) {
    set_error("preprocess: scratch too small (%zu < %zu)", scratch_bytes, need);
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* scr = static_cast<uint8_t*>(scratch);
  // Block shapes and shared-memory needs of the fused kernel over the batch (both targets of every image): the first
  // candidate pair that fits twice per SM, else the first that fits at all, else the two-pass kernels.
  static const int kCand[5][4] = {{128, 16, 64, 32}, {128, 16, 32, 16}, {64, 16, 32, 16}, {64, 8, 32, 8}, {32, 8, 16, 8}};
  FusedSmem fs{};
  size_t fused_smem = 0;
  bool fused = false;
  static const bool force_two_pass = [] { const char* e = std::getenv("RADVLM_B200_PREPROCESS"); return e && e[0] == '2'; }();
  if (!force_two_pass && (tile_size % 32) == 0) {
    for (int budget_pass = 0; budget_pass < 2 && !fused; ++budget_pass) {
      const size_t budget = budget_pass == 0 ? 110 * 1024 : 200 * 1024;
      for (int c = 0; c < 5 && !fused; ++c) {
        FusedSmem f{1, 64, 1, kCand[c][0], kCand[c][1], kCand[c][2], kCand[c][3]};
        for (int i = 0; i < n_images; ++i) {
          const radvlm_preprocess_image& im = images_host[i];
          for (int target = 0; target < 2; ++target) {
            const int ow = target ? tile_size : im.resized_w, oh = target ? tile_size : im.resized_h;
            const int nr = fused_span(im.height, oh, target ? f.bby : f.cby);
            const int nc = fused_span(im.width, ow, target ? f.bbx : f.cbx);
            const int pitch = static_cast<int>(al16(static_cast<size_t>(nc) * im.channels + 32));
            const int ksx = resample_ksize(im.width, ow);
            if (nr > f.max_nr) f.max_nr = nr;
            if (pitch > f.src_pitch) f.src_pitch = pitch;
            if (ksx > f.max_ksx) f.max_ksx = ksx;
          }
        }
        if (fused_total(f) <= budget) {
          fs = f;
          fused_smem = fused_total(f);
          fused = true;
        }
      }
    }
  }
  ProfScope ps(PROF_PREPROCESS, s, fused ? 2 : 3);
  resample_coeffs_kernel<<<dim3((max_out + 63) / 64, 4, n_images), 64, 0, s>>>(images_dev, scr, tile_size);
  RV_CUDA(cudaGetLastError());
  if (fused) {
    const int gx_c = (max_cw + fs.cbx - 1) / fs.cbx, gx_b = (tile_size + fs.bbx - 1) / fs.bbx;
    const int gy = (max_rows - tile_size) / fs.cby + (tile_size + fs.bby - 1) / fs.bby;
    dim3 fgrid(gx_c > gx_b ? gx_c : gx_b, gy, n_images);
    const int smem = static_cast<int>(fused_smem);
#define RV_LAUNCH_FUSED(TYPE)                                                                                          \
  do {                                                                                                                 \
    RV_CUDA(cudaFuncSetAttribute(resample_fused_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
    resample_fused_kernel<TYPE><<<fgrid, kFuThreads, smem, s>>>(images_dev, src, scr, static_cast<TYPE*>(tiles_out),  \
                                                                 tile_size, fs);                                        \
  } while (0)
    switch (out_dtype) {
      case RADVLM_DT_F32: RV_LAUNCH_FUSED(float); break;
      case RADVLM_DT_BF16: RV_LAUNCH_FUSED(__nv_bfloat16); break;
      case RADVLM_DT_F16: RV_LAUNCH_FUSED(__half); break;
      default:
        set_error("preprocess: unknown out_dtype %d", out_dtype);
        return RADVLM_ERR_BAD_ARGUMENT;
    }
#undef RV_LAUNCH_FUSED
    RV_CUDA(cudaGetLastError());
    return RADVLM_OK;
  }
  resample_h_kernel<<<dim3((max_out + 127) / 128, max_h, n_images * 2), 128, 0, s>>>(images_dev, src, scr,
                                                                                   tile_size);
  RV_CUDA(cudaGetLastError());
  dim3 grid((max_cw + 127) / 128, max_rows, n_images);
  switch (out_dtype) {
    case RADVLM_DT_F32:
      resample_v_tiles_kernel<float><<<grid, 128, 0, s>>>(images_dev, src, scr, static_cast<float*>(tiles_out), tile_size);
      break;
    case RADVLM_DT_BF16:
      resample_v_tiles_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(images_dev, src, scr, static_cast<__nv_bfloat16*>(tiles_out), tile_size);
      break;
    case RADVLM_DT_F16:
      resample_v_tiles_kernel<__half><<<grid, 128, 0, s>>>(images_dev, src, scr, static_cast<__half*>(tiles_out), tile_size);
      break;
    default:
      set_error("preprocess: unknown out_dtype %d", out_dtype);
      return RADVLM_ERR_BAD_ARGUMENT;
  }
  RV_CUDA(cudaGetLastError());
  return RADVLM_OK;
}
