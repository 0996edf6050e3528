"""Seeded input recipes shared by ``tests/golden/make_golden.py`` (reference side, container only) and the
tests (oracle / CUDA side, anywhere).  Fixtures store reference OUTPUTS; inputs are regenerated from here."""
from __future__ import annotations

import numpy as np
import torch

PINPOINTS = [[384 * i, 384 * j] for i in range(1, 7) for j in range(1, 7)]  # train.py:1583-1601

PARITY_SIZES = [(1024, 1024), (800, 1200), (3000, 1000), (1536, 1536), (2544, 3056), (384, 384), (500, 300)]


def planner_sizes():
    rng = np.random.default_rng(1234)
    sizes = list(PARITY_SIZES)
    sizes += [(64, 64), (1, 1), (383, 385), (385, 383), (2304, 2304), (4096, 64), (64, 4096), (2305, 2303),
              (1152, 1151), (1151, 1152), (767, 769), (769, 767), (1000, 3000), (3056, 2544), (135, 81)]
    for _ in range(300):
        sizes.append((int(rng.integers(64, 4097)), int(rng.integers(64, 4097))))
    for _ in range(60):  # near the pinpoint boundaries, where ties and int() truncation bite
        i, j = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        sizes.append((384 * i + int(rng.integers(-2, 3)), 384 * j + int(rng.integers(-2, 3))))
    return sizes


# ---------------------------------------------------------------------------------------------- merge / splice
MERGE_HIDDEN = 8
MERGE_VOCAB = 512


def _bf16_exact(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def merge_embed_table() -> torch.Tensor:
    g = torch.Generator().manual_seed(101)
    return _bf16_exact(torch.randn(MERGE_VOCAB, MERGE_HIDDEN, generator=g))


def merge_newline() -> torch.Tensor:
    g = torch.Generator().manual_seed(102)
    return _bf16_exact(torch.randn(MERGE_HIDDEN, generator=g))


def merge_cases():
    return {
        "c2": dict(seed=1, sizes=[(1024, 1024)], tiles=[10], lengths=[24], images_per_sample=[1], pad_to=24),
        "mixed": dict(seed=2, sizes=[(800, 1200), (384, 384), (3000, 1000), (500, 300), (2544, 3056), (384, 384)],
                      tiles=[13, 1, 13, 3, 31, 2], lengths=[40, 17, 65, 33, 9], images_per_sample=[1, 0, 2, 1, 1],
                      pad_to=72),
        "leftpad_trunc": dict(seed=3, sizes=[(1024, 1024), (500, 300)], tiles=[10, 3], lengths=[30, 50],
                              images_per_sample=[1, 1], pad_to=50, padding_side="left", max_length=5000),
        "anyres_nopool": dict(seed=4, sizes=[(1536, 1536)], tiles=[17], lengths=[12], images_per_sample=[1],
                              pad_to=12, aspect="anyres"),
        # the other spatial merge types of llava_arch.py:373-404 (RadVLM itself trains with spatial_unpad)
        "spatial_plain": dict(seed=5, sizes=[(1024, 1024), (500, 300), (384, 384)], tiles=[10, 3, 1], lengths=[20, 14, 9],
                              images_per_sample=[1, 1, 1], pad_to=20, merge_type="spatial"),
        "maxpool2x2": dict(seed=6, sizes=[(1024, 1024), (800, 400)], tiles=[10, 7], lengths=[18, 11],
                           images_per_sample=[1, 1], pad_to=18, merge_type="spatial_maxpool2x2"),
        "unpad_nobase": dict(seed=7, sizes=[(1024, 1024), (500, 300)], tiles=[10, 3], lengths=[15, 21],
                             images_per_sample=[1, 1], pad_to=21, merge_type="spatial_unpad_nobase"),
        "pad_2x2_unpad": dict(seed=8, sizes=[(500, 300), (640, 480)], tiles=[5, 5], lengths=[13, 17],
                              images_per_sample=[1, 1], pad_to=17, merge_type="spatial_unpad", aspect="pad"),
        "pad_2x2_plain_nobase": dict(seed=9, sizes=[(500, 300)], tiles=[5], lengths=[10], images_per_sample=[1],
                                     pad_to=10, merge_type="spatial_nobase", aspect="pad"),
    }


def video_cases():
    """Video / get_2dPool branch: `tiles` = frames for the entries whose modality is "video"."""
    return {
        "bilinear_grid": dict(seed=11, modalities=["video"], sizes=[(384, 384)], tiles=[4], lengths=[20],
                              images_per_sample=[1], pad_to=20, pool="bilinear", newline="grid"),
        "average_frame": dict(seed=12, modalities=["video"], sizes=[(384, 384)], tiles=[3], lengths=[16],
                              images_per_sample=[1], pad_to=16, pool="average", newline="frame"),
        "max_one_token": dict(seed=13, modalities=["video"], sizes=[(384, 384)], tiles=[2], lengths=[12],
                              images_per_sample=[1], pad_to=12, pool="max", newline="one_token"),
        "mixed_no_token": dict(seed=14, modalities=["video", "image", "image"], sizes=[(384, 384), (1024, 1024), (384, 384)],
                               tiles=[5, 10, 1], lengths=[30, 22, 9], images_per_sample=[1, 1, 1], pad_to=30,
                               pool="bilinear", newline="no_token"),
        "flat_average": dict(seed=15, modalities=["video"], sizes=[(384, 384)], tiles=[2], lengths=[10],
                             images_per_sample=[1], pad_to=10, pool="average", newline="grid", merge_type="flat"),
    }


def merge_features(case) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + case["seed"])
    return _bf16_exact(torch.randn(sum(case["tiles"]), 729, MERGE_HIDDEN, generator=g))


def merge_ids(case):
    """(input_ids [B,L], attention_mask [B,L] bool, labels [B,L]); right-padded raw batch like the collator's."""
    rng = np.random.default_rng(2000 + case["seed"])
    B, L = len(case["lengths"]), case["pad_to"]
    ids = np.zeros((B, L), dtype=np.int64)
    mask = np.zeros((B, L), dtype=bool)
    labels = np.full((B, L), -100, dtype=np.int64)
    for b, (n, k) in enumerate(zip(case["lengths"], case["images_per_sample"])):
        row = rng.integers(1, MERGE_VOCAB, size=n)
        if k:
            posn = np.sort(rng.choice(np.arange(1, n - 1), size=k, replace=False))
            row[posn] = -200
        ids[b, :n] = row
        mask[b, :n] = True
        lab = rng.integers(0, MERGE_VOCAB, size=n)
        lab[rng.random(n) < 0.4] = -100
        labels[b, :n] = lab
    return torch.from_numpy(ids), torch.from_numpy(mask), torch.from_numpy(labels)


# ---------------------------------------------------------------------------------------------- preprocessing
def preprocess_cases():
    return {
        "c2_1024_gray_noise": dict(seed=1, size=(1024, 1024), kind="gray_noise"),
        "rgb_500x300_noise": dict(seed=2, size=(500, 300), kind="rgb_noise"),
        "tall_800x1200_mix": dict(seed=3, size=(800, 1200), kind="rgb_mix"),
        "wide_3000x1000_noise": dict(seed=4, size=(3000, 1000), kind="gray_noise"),
        "exact_384_noise": dict(seed=5, size=(384, 384), kind="rgb_noise"),
        "small_130x100_mix": dict(seed=6, size=(130, 100), kind="rgb_mix", store=True),
        "mimic_2544x3056_mix": dict(seed=7, size=(2544, 3056), kind="gray_mix"),
        "canvas_exact_768x384": dict(seed=8, size=(768, 384), kind="rgb_noise"),
    }


def preprocess_image(case) -> np.ndarray:
    """uint8 [H,W,3] (rgb_*) or [H,W] (gray_*; replicated to RGB by the consumer)."""
    rng = np.random.default_rng(3000 + case["seed"])
    W, H = case["size"]
    kind = case["kind"]
    ch = () if kind.startswith("gray") else (3,)
    noise = rng.integers(0, 256, size=(H, W) + ch, dtype=np.uint8)
    if kind.endswith("noise"):
        return noise
    yy, xx = np.mgrid[0:H, 0:W]
    smooth = (127.5 + 127.5 * np.sin(xx / 37.0) * np.cos(yy / 53.0))
    if ch:
        smooth = np.stack([smooth, np.roll(smooth, 11, axis=0), np.roll(smooth, 23, axis=1)], axis=2)
    mix = 0.8 * smooth + 0.2 * noise
    return np.clip(np.rint(mix), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------- encoder
SMALL_VISION = dict(hidden_size=144, intermediate_size=272, num_hidden_layers=4, num_attention_heads=2)
SMALL_PROJ = 256
SMALL_SEED = 5
FULL_SEED = 7
FULL_SAMPLE_ROWS = [0, 1, 13, 26, 27, 100, 101, 200, 255, 256, 300, 364, 365, 400, 450, 500,
                    511, 512, 550, 600, 639, 640, 650, 700, 701, 710, 715, 720, 725, 726, 727, 728]


def normalize_lut_f32() -> np.ndarray:
    u = np.arange(256, dtype=np.uint8)
    x = (u.astype(np.float64) * (1 / 255)).astype(np.float32)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)


def encoder_pixels(n: int, seed: int) -> torch.Tensor:
    """[n,3,384,384] fp32: uint8 noise (grayscale replicated to RGB) through the preprocessing LUT (C1)."""
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 256, size=(n, 1, 384, 384), dtype=np.uint8)
    u = np.repeat(u, 3, axis=1)
    return torch.from_numpy(normalize_lut_f32()[u])


# ---------------------------------------------------------------------------------------------- training (gradients)
GRAD_SEED = 9
GRAD_IMAGES = ["rgb_500x300_noise", "grad_1536_gray_mix"]     # 2x1 grid (crop, no pool) and 4x4 grid (bilinear pool)
GRAD_IDS = [[11, 12, -200, 13, 14], [21, -200, 22, 23, 24, 25]]


def grad_image(name) -> np.ndarray:
    if name == "grad_1536_gray_mix":
        return preprocess_image(dict(seed=21, size=(1536, 1536), kind="gray_mix"))
    return preprocess_image(preprocess_cases()[name])


def grad_loss_weights(shape) -> torch.Tensor:
    """dL/d(inputs_embeds) of the synthetic loss  L = sum(inputs_embeds * R)"""
    g = torch.Generator(device="cpu").manual_seed(4242)
    return torch.randn(shape, generator=g)
