"""DataLoader-side integration of the fused preprocessing.

The reference tiles every image on the CPU inside the DataLoader workers
(``LazySupervisedDataset.process_image``, finetuning/llava/train/train.py:1060-1099: PIL open -> ``process_anyres_image``
-> fp32 ``[1+gw*gh, 3, 384, 384]``, ~0.7 s and 17.7 MB per 1024^2 image) and the collator only flattens the
``(image, image_size, modality)`` triples (train.py:1269-1281).  Here the workers ship the decoded uint8 pixels
(3.1 MB per 1024^2 image) and the tiling runs on the GPU (``radvlm_preprocess_anyres``), bit-identical to the reference.

    dataset side   : ``process_image_raw``      replaces  ``process_image``            (same return triple, image = uint8 HWC)
    collator side  : ``collate_images``         replaces  train.py:1269-1281          (same batch keys)
    training step  : ``tile_batch_on_device``   before    ``model(**batch)``          (batch["images"] becomes the list of
                                                                                       per-image tile tensors the model expects)
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


def _as_uint8_hwc(image) -> torch.Tensor:
    """PIL image / numpy array / tensor -> uint8 [H, W, 3] (or [H, W] for grayscale, replicated to RGB by the kernel)."""
    if isinstance(image, torch.Tensor):
        t = image
    else:
        arr = np.asarray(image)
        if arr.dtype != np.uint8:
            raise TypeError("raw image must be uint8, got %s" % arr.dtype)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.dtype != torch.uint8 or t.dim() not in (2, 3) or (t.dim() == 3 and t.shape[2] != 3):
        raise ValueError("raw image must be uint8 [H, W] or [H, W, 3], got %s %s" % (t.dtype, tuple(t.shape)))
    return t.contiguous()


def process_image_raw(image_file: str, image_folder: Optional[str], image_aspect_ratio: str,
                      overwrite_image_aspect_ratio: Optional[str] = None):
    """Dataset-side replacement of ``process_image`` (train.py:1060-1099) for the anyres modes: decode only.
    Returns the reference's triple ``(image, image_size, "image")`` with ``image`` = uint8 ``[H, W, 3]`` pixels and
    ``image_size`` = PIL ``(W, H)``."""
    from PIL import Image
    aspect = overwrite_image_aspect_ratio if overwrite_image_aspect_ratio is not None else image_aspect_ratio
    if not (aspect == "anyres" or "anyres_max" in aspect):
        raise NotImplementedError("radvlm_b200.data handles the anyres / anyres_max_N modes RadVLM trains with; got %r" % aspect)
    path = os.path.join(image_folder, image_file) if image_folder else image_file
    try:
        image = Image.open(path).convert("RGB")
    except Exception as exn:   # same behaviour as the reference: report, re-raise
        print(f"Failed to open image {image_file}. Exception:", exn)
        raise exn
    return _as_uint8_hwc(image), image.size, "image"


def collate_images(instances: Sequence[Dict], batch: Dict) -> Dict:
    """Collator-side image plumbing (train.py:1269-1281), unchanged semantics: ``instance["image"]`` is a list of
    ``(image, image_size, modality)`` triples; the batch gets ``image_sizes``, ``modalities`` and ``images`` (here the raw
    uint8 pixels, pinned when possible so the H2D copy is asynchronous)."""
    if "image" in instances[0]:
        images = [instance["image"] for instance in instances]
        batch["image_sizes"] = [im[1] for im_list in images for im in im_list]
        batch["modalities"] = [im[2] for im_list in images for im in im_list]
        raw = [_as_uint8_hwc(im[0]) for im_list in images for im in im_list]
        if torch.cuda.is_available():
            raw = [t if t.is_pinned() else t.pin_memory() for t in raw]
        batch["images"] = raw
    return batch


def tile_batch_on_device(batch: Dict, grid_pinpoints, device=None, dtype: torch.dtype = torch.bfloat16) -> Dict:
    """Turn the raw uint8 images of a collated batch into the per-image tile tensors the model consumes
    (``prepare_inputs_labels_for_multimodal(images=[...], image_sizes=[...])``).  ``image_sizes`` is re-derived from the
    pixels and checked against what the dataset reported."""
    from . import mm_utils
    raw = batch.get("images")
    if not raw or raw[0].dtype != torch.uint8:
        return batch
    tiles, sizes, splits, _ = mm_utils.preprocess_anyres_batch(raw, grid_pinpoints, device=device, dtype=dtype)
    want = [tuple(int(v) for v in s) for s in batch.get("image_sizes", sizes)]
    if want != [tuple(s) for s in sizes]:
        raise ValueError("image_sizes %s do not match the decoded pixels %s" % (want, sizes))
    batch["images"] = list(torch.split(tiles, splits, dim=0))
    batch["image_sizes"] = sizes
    return batch


def raw_bytes_saved(image_sizes: Sequence, tile_counts: Sequence[int], tile: int = 384) -> Dict[str, int]:
    """Host->device bytes per batch: the reference's fp32 tiles vs raw uint8 pixels."""
    ref = sum(int(n) * 3 * tile * tile * 4 for n in tile_counts)
    raw = sum(int(w) * int(h) * 3 for (w, h) in image_sizes)
    return {"reference_fp32_tiles": ref, "raw_uint8": raw}
