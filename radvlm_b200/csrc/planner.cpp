// Host-side planner (CPU, no CUDA): every integer decision of the encode path, bit-exact with the
// reference's Python float64 / int semantics.
//
//   select_best_resolution          finetuning/llava/mm_utils.py:119-149
//   get_anyres_image_grid_shape     finetuning/llava/mm_utils.py:213-240
//   resize_and_pad_image (geometry) finetuning/llava/mm_utils.py:152-188
//   unpad_image (window)            finetuning/llava/model/llava_arch.py:127-159
//   anyres_max pooling size         finetuning/llava/model/llava_arch.py:386-390
//   splice / truncate / pad layout  finetuning/llava/model/llava_arch.py:428-531
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "host_util.h"

namespace {

// CPython float floor division (Objects/floatobject.c float_floor_div / _float_div_mod):
// NOT floor(a / b) — e.g. 135 // 1.6666666666666667 == 80.0 while 135 / 1.666... rounds to 81.0.
double py_float_floordiv(double vx, double wx) {
  double mod = fmod(vx, wx);
  double div = (vx - mod) / wx;
  if (mod != 0.0) {
    if ((wx < 0) != (mod < 0)) {
      mod += wx;
      div -= 1.0;
    }
  }
  double floordiv;
  if (div != 0.0) {
    floordiv = floor(div);
    if (div - floordiv > 0.5) floordiv += 1.0;
  } else {
    floordiv = copysign(0.0, vx / wx);
  }
  return floordiv;
}

// mm_utils.py:119-149.  Candidates are (width, height); first best wins.
bool select_best_resolution(int W, int H, const int32_t* pin, int n, int* bw, int* bh) {
  bool found = false;
  long long max_eff = 0;
  double min_waste = INFINITY;
  for (int i = 0; i < n; ++i) {
    const int w = pin[2 * i], h = pin[2 * i + 1];
    const double sw = static_cast<double>(w) / static_cast<double>(W);
    const double sh = static_cast<double>(h) / static_cast<double>(H);
    const double scale = sw < sh ? sw : sh;  // Python min(): returns the first on ties, same value
    const long long dw = static_cast<long long>(static_cast<double>(W) * scale);
    const long long dh = static_cast<long long>(static_cast<double>(H) * scale);
    long long eff = dw * dh;
    const long long orig = static_cast<long long>(W) * H;
    if (orig < eff) eff = orig;
    const long long waste = static_cast<long long>(w) * h - eff;
    if (eff > max_eff || (eff == max_eff && static_cast<double>(waste) < min_waste)) {
      max_eff = eff;
      min_waste = static_cast<double>(waste);
      *bw = w;
      *bh = h;
      found = true;
    }
  }
  return found;
}

}  // namespace

extern "C" int radvlm_plan_select_best_resolution(int W, int H, const int32_t* pinpoints, int n,
                                                  int* best_w, int* best_h) {
  RV_CHECK_ARG(W > 0 && H > 0 && pinpoints && n > 0 && best_w && best_h, "plan: bad arguments");
  if (!select_best_resolution(W, H, pinpoints, n, best_w, best_h)) {
    // the reference returns None here and the caller's tuple-unpack raises
    rv::set_error("select_best_resolution: no candidate has a positive effective resolution");
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  return RADVLM_OK;
}

extern "C" int radvlm_plan_image(int W, int H, const int32_t* pinpoints, int n_pinpoints, int tile_size,
                                 int patches_per_side, int max_num_patches, radvlm_image_plan* out) {
  RV_CHECK_ARG(W > 0 && H > 0 && pinpoints && n_pinpoints > 0 && tile_size > 0 && patches_per_side > 0 && out,
               "plan_image: bad arguments");
  memset(out, 0, sizeof(*out));
  out->width = W;
  out->height = H;
  int bw = 0, bh = 0;
  if (!select_best_resolution(W, H, pinpoints, n_pinpoints, &bw, &bh)) {
    rv::set_error("select_best_resolution: no candidate has a positive effective resolution");
    return RADVLM_ERR_UNSUPPORTED_SHAPE;
  }
  out->best_w = bw;
  out->best_h = bh;
  out->grid_w = bw / tile_size;  // mm_utils.py:240
  out->grid_h = bh / tile_size;
  out->n_tiles = 1 + out->grid_w * out->grid_h;

  // resize_and_pad_image geometry (mm_utils.py:166-185)
  {
    const double scale_w = static_cast<double>(bw) / static_cast<double>(W);
    const double scale_h = static_cast<double>(bh) / static_cast<double>(H);
    int nw, nh;
    if (scale_w < scale_h) {
      nw = bw;
      const double c = ceil(static_cast<double>(H) * scale_w);
      nh = c < static_cast<double>(bh) ? static_cast<int>(c) : bh;
    } else {
      nh = bh;
      const double c = ceil(static_cast<double>(W) * scale_h);
      nw = c < static_cast<double>(bw) ? static_cast<int>(c) : bw;
    }
    out->resized_w = nw;
    out->resized_h = nh;
    out->paste_x = (bw - nw) / 2;
    out->paste_y = (bh - nh) / 2;
  }

  // unpad window in the (S*gh, S*gw) feature map (llava_arch.py:138-157)
  const int S = patches_per_side;
  const int ch = out->grid_h * S, cw = out->grid_w * S;
  int r0 = 0, c0 = 0, h = ch, w = cw;
  {
    const double orig_ar = static_cast<double>(W) / static_cast<double>(H);
    const double cur_ar = static_cast<double>(cw) / static_cast<double>(ch);
    if (orig_ar > cur_ar) {
      const double sf = static_cast<double>(cw) / static_cast<double>(W);
      const int new_h = static_cast<int>(static_cast<double>(H) * sf);
      const int pad = (ch - new_h) / 2;  // both operands non-negative: // == /
      r0 = pad;
      h = ch - 2 * pad;
    } else {
      const double sf = static_cast<double>(ch) / static_cast<double>(H);
      const int new_w = static_cast<int>(static_cast<double>(W) * sf);
      const int pad = (cw - new_w) / 2;
      c0 = pad;
      w = cw - 2 * pad;
    }
  }
  out->crop_r0 = r0;
  out->crop_c0 = c0;
  out->crop_h = h;
  out->crop_w = w;

  // anyres_max pooling (llava_arch.py:386-390); max_num_patches <= 0 means plain "anyres" (never pool)
  out->pool = 0;
  out->out_h = h;
  out->out_w = w;
  if (max_num_patches > 0) {
    const double times = sqrt(static_cast<double>(static_cast<long long>(h) * w) /
                              static_cast<double>(static_cast<long long>(max_num_patches) * S * S));
    if (times > 1.1) {
      out->pool = 1;
      out->out_h = static_cast<int>(py_float_floordiv(static_cast<double>(h), times));
      out->out_w = static_cast<int>(py_float_floordiv(static_cast<double>(w), times));
    }
  }
  out->n_tokens = S * S + out->out_h * (out->out_w + 1);
  return RADVLM_OK;
}

// ------------------------------------------------------------------------------------------------
// splice planner (llava_arch.py:428-531)
// ------------------------------------------------------------------------------------------------
extern "C" int radvlm_plan_splice(const int64_t* input_ids, const uint8_t* attention_mask, int B, int L,
                                  int image_token_index, const int32_t* image_tokens, int n_images,
                                  int64_t max_length /* <= 0: no truncation */, int left_pad,
                                  radvlm_splice_segment* segments, int segment_capacity, int* n_segments,
                                  int32_t* text_src, int text_capacity, int* n_text, int32_t* lengths,
                                  int* max_len_out) {
  RV_CHECK_ARG(input_ids && B > 0 && L > 0 && n_segments && n_text && lengths && max_len_out &&
                   (n_images == 0 || image_tokens),
               "plan_splice: bad arguments");
  struct Piece {
    int kind;  // 1 text, 2 image
    int len;
    int src_off;
    int image;
  };
  std::vector<std::vector<Piece>> per_sample(B);
  std::vector<int32_t> text;
  text.reserve(static_cast<size_t>(B) * L);
  int cur_image = 0;
  int max_len = 0;
  for (int b = 0; b < B; ++b) {
    std::vector<Piece>& pieces = per_sample[b];
    int n_img_tok = 0;
    for (int i = 0; i < L; ++i)
      if ((!attention_mask || attention_mask[static_cast<size_t>(b) * L + i]) &&
          input_ids[static_cast<size_t>(b) * L + i] == image_token_index)
        ++n_img_tok;
    if (n_img_tok == 0) {
      // text-only sample: consumes one image slot (llava_arch.py:452-459); an exhausted list raises
      if (cur_image >= n_images) {
        rv::set_error("IndexError: text-only sample %d needs image_features[%d] but only %d exist", b,
                      cur_image, n_images);
        return RADVLM_ERR_BAD_ARGUMENT;
      }
      ++cur_image;
    }
    long long total = 0;
    Piece cur{1, 0, static_cast<int>(text.size()), -1};
    for (int i = 0; i < L; ++i) {
      const size_t idx = static_cast<size_t>(b) * L + i;
      if (attention_mask && !attention_mask[idx]) continue;
      if (input_ids[idx] == image_token_index) {
        if (cur.len > 0) pieces.push_back(cur);
        total += cur.len;
        int img = cur_image;
        if (img >= n_images) img = cur_image - 1;  // IndexError fallback (llava_arch.py:478-481)
        if (img < 0 || img >= n_images) {
          rv::set_error("IndexError: sample %d needs image_features[%d] but only %d exist", b, cur_image,
                        n_images);
          return RADVLM_ERR_BAD_ARGUMENT;
        }
        ++cur_image;
        pieces.push_back(Piece{2, image_tokens[img], 0, img});
        total += image_tokens[img];
        cur = Piece{1, 0, static_cast<int>(text.size()), -1};
      } else {
        text.push_back(static_cast<int32_t>(idx));
        ++cur.len;
      }
    }
    if (cur.len > 0) pieces.push_back(cur);
    total += cur.len;
    if (max_length > 0 && total > max_length) total = max_length;  // x[:tokenizer_model_max_length]
    lengths[b] = static_cast<int32_t>(total);
    if (total > max_len) max_len = static_cast<int>(total);
  }
  // emit segments, clipped to lengths[b], plus zero padding
  int ns = 0;
  auto emit = [&](const radvlm_splice_segment& s) -> bool {
    if (s.length <= 0) return true;
    if (ns >= segment_capacity || segments == nullptr) {
      ++ns;
      return false;
    }
    segments[ns++] = s;
    return true;
  };
  bool fits = true;
  for (int b = 0; b < B; ++b) {
    const int len = lengths[b];
    const int pad = max_len - len;
    int64_t row = static_cast<int64_t>(b) * max_len;
    if (left_pad && pad > 0) {
      radvlm_splice_segment s{};
      s.dst_row = row; s.length = pad; s.kind = 0;
      fits &= emit(s);
      row += pad;
    }
    int pos = 0;
    for (const Piece& p : per_sample[b]) {
      if (pos >= len) break;
      int n = p.len;
      if (pos + n > len) n = len - pos;
      radvlm_splice_segment s{};
      s.dst_row = row; s.length = n; s.kind = p.kind; s.src_off = p.src_off; s.image = p.image; s.pos0 = pos;
      fits &= emit(s);
      row += n;
      pos += n;
    }
    if (!left_pad && pad > 0) {
      radvlm_splice_segment s{};
      s.dst_row = row; s.length = pad; s.kind = 0;
      fits &= emit(s);
    }
  }
  *n_segments = ns;
  *n_text = static_cast<int>(text.size());
  *max_len_out = max_len;
  if (!fits || (text_src == nullptr && !text.empty()) || static_cast<int>(text.size()) > text_capacity) {
    rv::set_error("plan_splice: output capacity too small (need %d segments, %zu text slots)", ns,
                  text.size());
    return RADVLM_ERR_WORKSPACE_TOO_SMALL;
  }
  if (!text.empty()) memcpy(text_src, text.data(), text.size() * sizeof(int32_t));
  return RADVLM_OK;
}
