"""GPU replacements for the reference's host preprocessing (same names / argument meaning).

  * ``process_anyres_image``  finetuning/llava/mm_utils.py:243-293 (+ SigLipImageProcessor.preprocess,
    siglip_encoder.py:47-67) — runs resize / pad / tile / normalise in one fused CUDA pipeline,
    bit-exact with the PIL + numpy path.
  * ``process_images``        finetuning/llava/mm_utils.py:314-338 (anyres branch)
  * ``preprocess_anyres_batch`` — batched form used by the data-parallel encode path: raw uint8 images
    in, one ``[total_tiles, 3, 384, 384]`` tensor out (DataLoader workers only ship uint8 + (W,H)).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, planner

_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16}


def _to_uint8_hwc(image) -> torch.Tensor:
    """PIL image / numpy array / tensor -> contiguous uint8 tensor [H, W, 3] or [H, W] (grayscale)."""
    if isinstance(image, torch.Tensor):
        t = image
    elif isinstance(image, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(image))
    else:  # PIL.Image: the reference operates on RGB images (convert_to_rgb, siglip_encoder.py:55)
        t = torch.from_numpy(np.asarray(image.convert("RGB")).copy())
    if t.dtype != torch.uint8:
        raise TypeError("images must be uint8, got %s" % t.dtype)
    if t.dim() == 3 and t.shape[2] == 1:
        t = t[:, :, 0]
    if not (t.dim() == 2 or (t.dim() == 3 and t.shape[2] == 3)):
        raise ValueError("image must be [H,W,3] or [H,W], got %s" % (tuple(t.shape),))
    return t.contiguous()


_COPY_STREAMS = {}   # device -> the H2D stream of preprocess_anyres_batch


def preprocess_anyres_batch(images: Sequence, grid_pinpoints, device=None, dtype: torch.dtype = torch.float32,
                            tile_size: int = 384, patches_per_side: int = 27,
                            max_num_patches: Optional[int] = 9):
    """Returns (tiles [sum(1+gw*gh), 3, S, S], image_sizes [(W,H)], split_sizes, plans)."""
    lib = _lib.load()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("radvlm_b200: preprocessing runs on a CUDA device only (got %s); no CPU fallback" % device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    srcs = [_to_uint8_hwc(im) for im in images]
    n = len(srcs)
    plans, descs = [], (_lib.PreprocessImage * n)()
    scratch_off = tile_base = 0
    sizes, splits = [], []
    with torch.cuda.device(device):
        # host images go H2D straight from their own (ideally pinned) storage; device images are used in place.  The
        # copies run on a copy stream of their own, so a caller that enqueues ahead of the GPU (the encode path never
        # blocks the host) gets the next batch's transfer under the current batch's kernels instead of in between them.
        if any(t.device != device for t in srcs):
            cur = torch.cuda.current_stream(device)
            cs = _COPY_STREAMS.get(device)
            if cs is None:
                cs = _COPY_STREAMS[device] = torch.cuda.Stream(device)
            with torch.cuda.stream(cs):
                dev_srcs = [t if t.device == device else t.to(device, non_blocking=True) for t in srcs]
            cur.wait_stream(cs)
            for t, d in zip(srcs, dev_srcs):
                if d is not t:
                    d.record_stream(cur)     # allocated on the copy stream, consumed by the kernel on `cur`
        else:
            dev_srcs = srcs
        base_ptr = min(t.data_ptr() for t in dev_srcs)
        for i, t in enumerate(dev_srcs):
            H, W = int(t.shape[0]), int(t.shape[1])
            ch = 1 if t.dim() == 2 else 3
            p = planner.plan_image((W, H), grid_pinpoints, tile_size, patches_per_side, max_num_patches or 0)
            plans.append(p)
            d = descs[i]
            d.src_offset, d.scratch_offset = t.data_ptr() - base_ptr, scratch_off
            d.width, d.height, d.channels = W, H, ch
            d.grid_w, d.grid_h = p.grid_w, p.grid_h
            d.resized_w, d.resized_h, d.paste_x, d.paste_y = p.resized_w, p.resized_h, p.paste_x, p.paste_y
            d.tile_base = tile_base
            scratch_off += lib.radvlm_preprocess_scratch_bytes(W, H, ch, p.resized_w, p.resized_h, tile_size)
            tile_base += p.n_tiles
            sizes.append((W, H))
            splits.append(p.n_tiles)
        desc_host = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8)
        if torch.cuda.is_available():
            desc_host = desc_host.pin_memory()
        desc_dev = desc_host.to(device, non_blocking=True)
        scratch = torch.empty(max(scratch_off, 16), dtype=torch.uint8, device=device)
        tiles = torch.empty(tile_base, 3, tile_size, tile_size, dtype=dtype, device=device)
        _lib.check(lib.radvlm_preprocess_anyres(
            base_ptr, desc_dev.data_ptr(), descs, n, tile_size, tiles.data_ptr(), _DT[dtype],
            scratch.data_ptr(), scratch.numel(), torch.cuda.current_stream(device).cuda_stream))
    return tiles, sizes, splits, plans


def process_anyres_image(image, processor, grid_pinpoints, device=None, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Drop-in for mm_utils.py:243-293: one image -> [1+gw*gh, 3, 384, 384] (on the GPU)."""
    tile = 384
    if processor is not None:
        cs = getattr(processor, "crop_size", None)
        if isinstance(cs, dict) and "height" in cs:
            tile = int(cs["height"])
    pps = tile // 14
    tiles, _, _, _ = preprocess_anyres_batch([image], grid_pinpoints, device=device, dtype=dtype, tile_size=tile,
                                             patches_per_side=pps)
    return tiles


def process_images(images, image_processor, model_cfg, device=None, dtype: torch.dtype = torch.float32):
    """Drop-in for the anyres branch of mm_utils.py:314-338; returns a list of per-image tile tensors
    (or one stacked tensor when all images have the same tile count, like the reference)."""
    aspect = getattr(model_cfg, "image_aspect_ratio", None)
    if not (aspect == "anyres" or (aspect is not None and "anyres_max" in aspect)):
        raise NotImplementedError("radvlm_b200.process_images implements the anyres modes RadVLM trains with; got %r" % (aspect,))
    tiles, _, splits, _ = preprocess_anyres_batch(images, model_cfg.image_grid_pinpoints, device=device, dtype=dtype)
    new_images = list(torch.split(tiles, splits, dim=0))
    if all(x.shape == new_images[0].shape for x in new_images):
        return torch.stack(new_images, dim=0)
    return new_images


class SigLipImageProcessor:
    """Drop-in for siglip_encoder.py:34-67: same constructor attributes; ``preprocess(images, return_tensors)`` returns
    ``{"pixel_values": [n, 3, 384, 384]}`` (convert to RGB, aspect-distorting BICUBIC resize to ``size``, rescale 1/255,
    normalise with mean = std = 0.5), computed by the fused preprocessing kernel (it is the base tile of the anyres
    tiling), bit-exact with the reference.  Only the reference's constants are supported."""

    def __init__(self, image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), size=(384, 384), crop_size=None,
                 resample=3, rescale_factor=1 / 255, data_format="channels_first"):
        self.image_mean, self.image_std, self.size = image_mean, image_std, size
        self.resample, self.rescale_factor, self.data_format = resample, rescale_factor, data_format
        self.crop_size = crop_size if crop_size is not None else {"height": 384, "width": 384}
        if (tuple(image_mean) != (0.5, 0.5, 0.5) or tuple(image_std) != (0.5, 0.5, 0.5) or int(resample) != 3
                or abs(rescale_factor - 1 / 255) > 1e-12 or size[0] != size[1]):
            raise NotImplementedError("radvlm_b200.SigLipImageProcessor implements the reference's constants "
                                      "(mean = std = 0.5, BICUBIC, 1/255, square size)")

    def preprocess(self, images, return_tensors="pt", device=None, dtype: torch.dtype = torch.float32):
        if not isinstance(images, (list, tuple)):
            images = [images]
        side = int(self.size[0])
        tiles, _, splits, _ = preprocess_anyres_batch(list(images), [[side, side]], device=device, dtype=dtype,
                                                      tile_size=side, patches_per_side=side // 14)
        firsts, base = [], 0
        for n in splits:              # tile 0 of every image = the whole image resized to side x side
            firsts.append(base)
            base += n
        pixel_values = tiles[torch.tensor(firsts, device=tiles.device)]
        if return_tensors in (None, "np"):
            pixel_values = pixel_values.float().cpu().numpy()
            return {"pixel_values": list(pixel_values) if return_tensors is None else pixel_values}
        return {"pixel_values": pixel_values}
