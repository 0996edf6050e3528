"""TEST INFRASTRUCTURE ONLY (oracle) — numpy restatement of the reference's host preprocessing.

  process_anyres_image / resize_and_pad_image / divide_to_patches   mm_utils.py:152-210,243-293
  SigLipImageProcessor.preprocess                                   siglip_encoder.py:47-67

Third-party arithmetic that is not under /root/reference, restated from its published algorithm:
  * Pillow (pinned 10.3.0 in finetuning/requirements.txt:199; 12.2.0 installed here) ``Image.resize`` with the
    default BICUBIC filter on 8-bit images = ``ImagingResample`` (src/libImaging/Resample.c):
    ``precompute_coeffs`` (float64 bicubic a=-0.5, support 2*max(1, in/out)), ``normalize_coeffs_8bpc``
    (22-bit fixed point, round half away from zero), horizontal pass first into a uint8 image, then the
    vertical pass; accumulator seeded with 1<<21, arithmetic shift by 22, clip to [0,255].
  * transformers.image_transforms.rescale / normalize: float32(float64(u8) * (1/255)); (x-0.5f)/0.5f.
Pinned: ``tests/test_oracle_pinned.py`` checks this file bit-exactly against PIL / the reference's
``process_anyres_image`` in the container, and ``tests/golden/preprocess_golden.npz`` holds reference outputs.
"""
from __future__ import annotations

import math

import numpy as np

from . import planner_oracle as po

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (ksize, bounds [out,2], kk int32 [out,ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample along axis 0 of a uint8 array [in, ...] -> [out, ...]."""
    in_size = img.shape[0]
    _, bounds, kk = precompute_coeffs(in_size, out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = int(bounds[xx, 0]), int(bounds[xx, 1])
        k = kk[xx, :xmax].astype(np.int64)
        acc = np.tensordot(k, src[xmin:xmin + xmax], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bicubic(img: np.ndarray, size) -> np.ndarray:
    """``PIL.Image.resize((w, h))`` (BICUBIC) on a uint8 [H, W, C] array."""
    out_w, out_h = size
    H, W = img.shape[:2]
    if (out_w, out_h) == (W, H):
        return img.copy()
    cur = img
    if out_w != W:  # horizontal pass first
        cur = np.swapaxes(_resample_axis0(np.ascontiguousarray(np.swapaxes(cur, 0, 1)), out_w), 0, 1)
    if out_h != H:
        cur = _resample_axis0(np.ascontiguousarray(cur), out_h)
    return np.ascontiguousarray(cur)


def normalize_lut() -> np.ndarray:
    """siglip_encoder.py:55-62 on every uint8 value: rescale then normalize (fp32)."""
    u = np.arange(256, dtype=np.uint8)
    x = (u.astype(np.float64) * (1 / 255)).astype(np.float32)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)


def anyres_tiles_uint8(img: np.ndarray, possible_resolutions=None, tile: int = 384) -> np.ndarray:
    """mm_utils.py:275-291 on uint8 HWC: [1+gw*gh, tile, tile, 3], base tile first then row-major crops."""
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, axis=2)
    H, W = img.shape[:2]
    if possible_resolutions is None:
        possible_resolutions = po.default_pinpoints(tile)
    best = po.select_best_resolution((W, H), possible_resolutions)
    nw, nh, px, py = po.resize_and_pad_geometry((W, H), best)
    resized = pil_resize_bicubic(img, (nw, nh))
    canvas = np.zeros((best[1], best[0], 3), dtype=np.uint8)
    canvas[py:py + nh, px:px + nw] = resized
    tiles = [pil_resize_bicubic(img, (tile, tile))]
    for i in range(0, best[1], tile):
        for j in range(0, best[0], tile):
            tiles.append(canvas[i:i + tile, j:j + tile])
    return np.stack(tiles, axis=0)


def process_anyres_image(img: np.ndarray, possible_resolutions=None, tile: int = 384) -> np.ndarray:
    """Full reference preprocessing: uint8 HWC -> fp32 [n, 3, tile, tile]."""
    t = anyres_tiles_uint8(img, possible_resolutions, tile)
    return np.ascontiguousarray(normalize_lut()[t].transpose(0, 3, 1, 2))


def siglip_image_processor(img: np.ndarray, size: int = 384) -> np.ndarray:
    """SigLipImageProcessor.preprocess for one image (siglip_encoder.py:47-67): convert_to_rgb -> resize((size, size),
    BICUBIC; aspect is NOT preserved) -> rescale(1/255) -> normalize(0.5, 0.5) -> CHW fp32 [3, size, size]."""
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, axis=2)
    r = pil_resize_bicubic(img, (size, size))
    return np.ascontiguousarray(normalize_lut()[r].transpose(2, 0, 1))
