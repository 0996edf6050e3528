"""Host-side planning for the encode path (thin Python over the C planner in ``csrc/planner.cpp``).

Mirrors the reference's helper names and argument conventions (sizes are PIL ``(width, height)``):

  * ``select_best_resolution``         finetuning/llava/mm_utils.py:119-149
  * ``get_anyres_image_grid_shape``    finetuning/llava/mm_utils.py:213-240
  * ``plan_image``                     mm_utils.py:152-188 + llava_arch.py:127-159,386-390
  * ``plan_splice``                    llava_arch.py:428-531
"""
from __future__ import annotations

import ast
import ctypes as C
import re
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

IGNORE_INDEX = -100        # finetuning/llava/constants.py:7
IMAGE_TOKEN_INDEX = -200   # finetuning/llava/constants.py:8


def parse_grid_pinpoints(grid_pinpoints, patch_size: int) -> List[List[int]]:
    """Resolve ``image_grid_pinpoints`` to a list of ``[width, height]`` (mm_utils.py:225-238)."""
    if isinstance(grid_pinpoints, str) and "x" in grid_pinpoints:
        assert patch_size in [224, 336, 384, 448, 512], "patch_size should be in [224, 336, 384, 448, 512]"
        matches = re.findall(r"\((\d+)x(\d+)\)", grid_pinpoints)
        range_start = tuple(map(int, matches[0]))
        range_end = tuple(map(int, matches[-1]))
        grid = [(i, j) for i in range(range_start[0], range_end[0] + 1) for j in range(range_start[1], range_end[1] + 1)]
        return [[dim * patch_size for dim in pair] for pair in grid]
    if type(grid_pinpoints) is list:
        return grid_pinpoints
    return ast.literal_eval(grid_pinpoints)


def _pin_array(possible_resolutions: Sequence[Sequence[int]]):
    flat = [int(v) for pair in possible_resolutions for v in pair]
    return (C.c_int32 * len(flat))(*flat), len(flat) // 2


def select_best_resolution(original_size: Tuple[int, int], possible_resolutions) -> Optional[Tuple[int, int]]:
    lib = _lib.load()
    arr, n = _pin_array(possible_resolutions)
    bw, bh = C.c_int(0), C.c_int(0)
    st = lib.radvlm_plan_select_best_resolution(int(original_size[0]), int(original_size[1]), arr, n,
                                                C.byref(bw), C.byref(bh))
    if st == _lib.ERR_UNSUPPORTED_SHAPE:
        return None  # the reference returns None (best_fit never assigned)
    _lib.check(st)
    return bw.value, bh.value


def get_anyres_image_grid_shape(image_size, grid_pinpoints, patch_size: int) -> Tuple[int, int]:
    """Returns ``(num_patch_width, num_patch_height)`` like the reference (mm_utils.py:239-240)."""
    width, height = select_best_resolution(image_size, parse_grid_pinpoints(grid_pinpoints, patch_size))
    return width // patch_size, height // patch_size


def plan_image(image_size, grid_pinpoints, tile_size: int = 384, patches_per_side: int = 27,
               max_num_patches: int = 9) -> _lib.ImagePlan:
    lib = _lib.load()
    arr, n = _pin_array(parse_grid_pinpoints(grid_pinpoints, tile_size))
    plan = _lib.ImagePlan()
    _lib.check(lib.radvlm_plan_image(int(image_size[0]), int(image_size[1]), arr, n, int(tile_size),
                                     int(patches_per_side), int(max_num_patches or 0), C.byref(plan)))
    return plan


class SplicePlan:
    __slots__ = ("segments", "n_segments", "text_src", "n_text", "lengths", "max_len")


def plan_splice(input_ids: np.ndarray, attention_mask: Optional[np.ndarray], image_tokens: Sequence[int],
                max_length: Optional[int], left_pad: bool, image_token_index: int = IMAGE_TOKEN_INDEX) -> SplicePlan:
    """input_ids: int64 [B, L] (host); attention_mask: bool/uint8 [B, L] or None."""
    lib = _lib.load()
    ids = np.ascontiguousarray(input_ids, dtype=np.int64)
    B, L = ids.shape
    mask = None
    if attention_mask is not None:
        mask = np.ascontiguousarray(attention_mask, dtype=np.uint8)
        assert mask.shape == ids.shape
    n_images = len(image_tokens)
    tok = (C.c_int32 * max(n_images, 1))(*[int(t) for t in image_tokens])
    seg_cap = B * (2 * L + 3) + 4
    segments = (_lib.SpliceSegment * seg_cap)()
    text_cap = B * L
    text_src = np.empty(max(text_cap, 1), dtype=np.int32)
    lengths = np.empty(B, dtype=np.int32)
    n_seg, n_text, max_len = C.c_int(0), C.c_int(0), C.c_int(0)
    st = lib.radvlm_plan_splice(
        ids.ctypes.data, mask.ctypes.data if mask is not None else None, B, L, int(image_token_index),
        tok, n_images, int(max_length) if max_length else 0, 1 if left_pad else 0,
        segments, seg_cap, C.byref(n_seg), text_src.ctypes.data_as(C.POINTER(C.c_int32)), text_cap,
        C.byref(n_text), lengths.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(max_len))
    if st == _lib.ERR_BAD_ARGUMENT and _lib.last_error().startswith("IndexError"):
        raise IndexError(_lib.last_error())  # same exception type the reference's list indexing raises
    _lib.check(st)
    plan = SplicePlan()
    plan.segments = segments
    plan.n_segments = n_seg.value
    plan.text_src = text_src[: n_text.value]
    plan.n_text = n_text.value
    plan.lengths = lengths
    plan.max_len = max_len.value
    return plan
