"""TEST INFRASTRUCTURE ONLY (oracle) — pure-Python restatement of the reference's integer planning.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this; the product path (``radvlm_b200/``) never does.  Pinned against the real reference functions
by ``tests/test_oracle_pinned.py`` (container) and by ``tests/golden/planner_golden.json`` (everywhere).

Each function follows the cited reference lines statement by statement (Python float64 / int semantics).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

IGNORE_INDEX = -100       # finetuning/llava/constants.py:7
IMAGE_TOKEN_INDEX = -200  # finetuning/llava/constants.py:8


def default_pinpoints(patch: int = 384, lo: int = 1, hi: int = 6) -> List[List[int]]:
    """train.py:1583-1601 — "(1x1),...,(6x6)" expanded i-major."""
    return [[patch * i, patch * j] for i in range(lo, hi + 1) for j in range(lo, hi + 1)]


def select_best_resolution(original_size, possible_resolutions):
    """mm_utils.py:119-149."""
    original_width, original_height = original_size
    best_fit = None
    max_effective_resolution = 0
    min_wasted_resolution = float("inf")
    for width, height in possible_resolutions:
        scale = min(width / original_width, height / original_height)
        downscaled_width, downscaled_height = int(original_width * scale), int(original_height * scale)
        effective_resolution = min(downscaled_width * downscaled_height, original_width * original_height)
        wasted_resolution = (width * height) - effective_resolution
        if effective_resolution > max_effective_resolution or (
                effective_resolution == max_effective_resolution and wasted_resolution < min_wasted_resolution):
            max_effective_resolution = effective_resolution
            min_wasted_resolution = wasted_resolution
            best_fit = (width, height)
    return best_fit


def get_anyres_image_grid_shape(image_size, possible_resolutions, patch_size):
    """mm_utils.py:213-240 (list form of grid_pinpoints) -> (num_patch_width, num_patch_height)."""
    width, height = select_best_resolution(image_size, possible_resolutions)
    return width // patch_size, height // patch_size


def resize_and_pad_geometry(original_size, target_resolution):
    """mm_utils.py:152-188 -> (new_width, new_height, paste_x, paste_y)."""
    original_width, original_height = original_size
    target_width, target_height = target_resolution
    scale_w = target_width / original_width
    scale_h = target_height / original_height
    if scale_w < scale_h:
        new_width = target_width
        new_height = min(math.ceil(original_height * scale_w), target_height)
    else:
        new_height = target_height
        new_width = min(math.ceil(original_width * scale_h), target_width)
    paste_x = (target_width - new_width) // 2
    paste_y = (target_height - new_height) // 2
    return new_width, new_height, paste_x, paste_y


def unpad_window(original_size, current_height, current_width):
    """llava_arch.py:127-159 -> (r0, r1, c0, c1) slice bounds of the centre crop."""
    original_width, original_height = original_size
    original_aspect_ratio = original_width / original_height
    current_aspect_ratio = current_width / current_height
    if original_aspect_ratio > current_aspect_ratio:
        scale_factor = current_width / original_width
        new_height = int(original_height * scale_factor)
        padding = (current_height - new_height) // 2
        return padding, current_height - padding, 0, current_width
    scale_factor = current_height / original_height
    new_width = int(original_width * scale_factor)
    padding = (current_width - new_width) // 2
    return 0, current_height, padding, current_width - padding


def plan_image(image_size, possible_resolutions=None, tile_size=384, unit=27, max_num_patches: Optional[int] = 9):
    """All per-image integers of the path (llava_arch.py:350-406)."""
    if possible_resolutions is None:
        possible_resolutions = default_pinpoints(tile_size)
    W, H = image_size
    best = select_best_resolution(image_size, possible_resolutions)
    gw, gh = best[0] // tile_size, best[1] // tile_size
    nw, nh, px, py = resize_and_pad_geometry(image_size, best)
    ch, cw = gh * unit, gw * unit
    r0, r1, c0, c1 = unpad_window(image_size, ch, cw)
    h, w = r1 - r0, c1 - c0
    pool, oh, ow = False, h, w
    if max_num_patches:
        times = math.sqrt(h * w / (max_num_patches * unit ** 2))      # llava_arch.py:387
        if times > 1.1:
            pool = True
            oh, ow = int(h // times), int(w // times)                 # llava_arch.py:390
    return dict(width=W, height=H, best_w=best[0], best_h=best[1], grid_w=gw, grid_h=gh, resized_w=nw,
                resized_h=nh, paste_x=px, paste_y=py, n_tiles=1 + gw * gh, crop_r0=r0, crop_c0=c0, crop_h=h,
                crop_w=w, pool=int(pool), out_h=oh, out_w=ow, n_tokens=unit * unit + oh * (ow + 1))


def splice_layout(input_ids: Sequence[Sequence[int]], attention_mask, image_tokens: Sequence[int],
                  max_length: Optional[int], left_pad: bool, n_modalities: Optional[int] = None):
    """llava_arch.py:428-531 on plain Python lists.

    Returns dict(max_len, lengths, rows) where rows[b] is a list of per-position sources:
    ("text", b, i) | ("image", image_idx, token_idx) | ("pad",), already truncated and padded.
    Raises IndexError exactly where the reference's list indexing would.
    """
    B = len(input_ids)
    seqs = []
    cur_image_idx = 0
    n_images = len(image_tokens)
    for b in range(B):
        ids = input_ids[b]
        keep = [i for i in range(len(ids)) if attention_mask is None or attention_mask[b][i]]
        num_images = sum(1 for i in keep if ids[i] == IMAGE_TOKEN_INDEX)
        row = []
        if num_images == 0:
            if cur_image_idx >= n_images:
                raise IndexError("list index out of range")           # llava_arch.py:453 (uncaught)
            row = [("text", b, i) for i in keep]                       # + image_features[idx][0:0]
            cur_image_idx += 1
            seqs.append(row)
            continue
        for i in keep:
            if ids[i] == IMAGE_TOKEN_INDEX:
                idx = cur_image_idx
                if idx >= n_images:                                    # llava_arch.py:478-481
                    idx = cur_image_idx - 1
                    if idx < 0 or idx >= n_images:
                        raise IndexError("list index out of range")
                cur_image_idx += 1
                row.extend(("image", idx, t) for t in range(image_tokens[idx]))
            else:
                row.append(("text", b, i))
        seqs.append(row)
    if n_modalities is not None:
        seqs = seqs[:n_modalities]                                     # zip(new_input_embeds, modalities)
    if max_length:
        seqs = [s[:max_length] for s in seqs]                          # llava_arch.py:499-500
    max_len = max(len(s) for s in seqs)
    rows = []
    for s in seqs:
        pad = [("pad",)] * (max_len - len(s))
        rows.append(pad + s if left_pad else s + pad)
    return dict(max_len=max_len, lengths=[len(s) for s in seqs], rows=rows)


def merge_source(plan: dict, t: int, unit: int = 27):
    """Source of visual token ``t`` of one anyres image:
    ("feat", tile, token) | ("newline",) | ("bilinear", r, c) for pooled positions."""
    T = unit * unit
    if t < T:
        return ("feat", 0, t)
    u = t - T
    ow = plan["out_w"]
    r, c = divmod(u, ow + 1)
    if c == ow:
        return ("newline",)
    if plan["pool"]:
        return ("bilinear", r, c)
    R, Cc = r + plan["crop_r0"], c + plan["crop_c0"]
    return ("feat", 1 + (R // unit) * plan["grid_w"] + (Cc // unit), (R % unit) * unit + (Cc % unit))
