"""DataLoader-side integration (radvlm_b200.data): the collator keeps the reference's batch keys / flattening order
(train.py:1269-1281) and ships raw uint8 pixels; tiling on the device is covered by the -m gpu test below."""
import numpy as np
import pytest
import torch

import golden_inputs as gi
from radvlm_b200 import data as rdata


def _instances():
    a = torch.from_numpy(gi.preprocess_image(gi.preprocess_cases()["rgb_500x300_noise"]))
    b = torch.from_numpy(gi.preprocess_image(gi.preprocess_cases()["small_130x100_mix"]))
    c = torch.from_numpy(gi.preprocess_image(gi.preprocess_cases()["exact_384_noise"]))
    return [{"image": [(a, (500, 300), "image")]}, {"image": [(b, (130, 100), "image"), (c, (384, 384), "image")]}]


def test_collate_images_matches_reference_flattening():
    batch = rdata.collate_images(_instances(), {})
    assert batch["image_sizes"] == [(500, 300), (130, 100), (384, 384)]
    assert batch["modalities"] == ["image"] * 3
    assert [tuple(t.shape) for t in batch["images"]] == [(300, 500, 3), (100, 130, 3), (384, 384, 3)]
    assert all(t.dtype == torch.uint8 for t in batch["images"])
    assert rdata.collate_images([{"input_ids": 1}], {"x": 1}) == {"x": 1}      # text-only batch: untouched


def test_raw_image_validation_and_bytes():
    with pytest.raises(TypeError):
        rdata._as_uint8_hwc(np.zeros((4, 4, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        rdata._as_uint8_hwc(torch.zeros(3, 4, 4, dtype=torch.uint8))
    assert tuple(rdata._as_uint8_hwc(np.zeros((5, 7), dtype=np.uint8)).shape) == (5, 7)
    b = rdata.raw_bytes_saved([(1024, 1024)], [10])
    assert b == {"reference_fp32_tiles": 10 * 3 * 384 * 384 * 4, "raw_uint8": 1024 * 1024 * 3}
    with pytest.raises(NotImplementedError):
        rdata.process_image_raw("x.png", None, "pad")


@pytest.mark.gpu
def test_tile_batch_on_device_equals_reference_preprocessing():
    from oracle import resample_oracle as ro
    batch = rdata.collate_images(_instances(), {})
    batch = rdata.tile_batch_on_device(batch, gi.PINPOINTS, dtype=torch.float32)
    assert batch["image_sizes"] == [(500, 300), (130, 100), (384, 384)]
    for tiles, name in zip(batch["images"], ["rgb_500x300_noise", "small_130x100_mix", "exact_384_noise"]):
        ref = ro.process_anyres_image(gi.preprocess_image(gi.preprocess_cases()[name]), gi.PINPOINTS)
        assert np.array_equal(tiles.cpu().numpy(), ref), name          # bit-exact with the (pinned) Pillow restatement
    bad = rdata.collate_images(_instances(), {})
    bad["image_sizes"][0] = (300, 500)
    with pytest.raises(ValueError):
        rdata.tile_batch_on_device(bad, gi.PINPOINTS)
