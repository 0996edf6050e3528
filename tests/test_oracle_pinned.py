"""Container-only: the oracle restatements against the LIVE reference (and PIL) — skipped where /root/reference
is absent (the GPU box), where the committed golden vectors (tests/test_oracle_golden.py) take over."""
import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import planner_oracle as po
from oracle import resample_oracle as ro
from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_loader import import_reference
    return import_reference()


def test_pil_bicubic_bit_exact_direct():
    from PIL import Image
    rng = np.random.default_rng(3)
    for (w, h), (ow, oh) in [((1024, 1024), (1152, 1152)), ((1024, 1024), (384, 384)), ((800, 1200), (1024, 1536)),
                             ((333, 517), (384, 384)), ((3000, 1000), (2304, 768)), ((37, 29), (384, 384)),
                             ((500, 300), (768, 461)), ((640, 480), (640, 384)), ((640, 480), (500, 480))]:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((ow, oh)))
        assert np.array_equal(ro.pil_resize_bicubic(img, (ow, oh)), want), ((w, h), (ow, oh))


def test_planner_oracle_random_vs_reference(ref):
    rng = np.random.default_rng(5)
    for _ in range(1500):
        W, H = int(rng.integers(16, 5000)), int(rng.integers(16, 5000))
        assert tuple(po.select_best_resolution((W, H), gi.PINPOINTS)) == tuple(ref.mm_utils.select_best_resolution((W, H), gi.PINPOINTS))
        gw, gh = ref.mm_utils.get_anyres_image_grid_shape((W, H), gi.PINPOINTS, 384)
        assert po.get_anyres_image_grid_shape((W, H), gi.PINPOINTS, 384) == (gw, gh)
        idx = torch.arange(gh * 27 * gw * 27, dtype=torch.float32).view(1, gh * 27, gw * 27)
        u = ref.llava_arch.unpad_image(idx, (W, H))
        r0, r1, c0, c1 = po.unpad_window((W, H), gh * 27, gw * 27)
        assert (r1 - r0, c1 - c0) == tuple(u.shape[1:]) and int(u[0, 0, 0]) == r0 * gw * 27 + c0


def test_process_anyres_image_oracle_vs_reference(ref):
    from PIL import Image
    proc = ref.siglip_encoder.SigLipImageProcessor()
    rng = np.random.default_rng(9)
    for (w, h) in [(1024, 1024), (641, 377), (384, 384), (200, 900)]:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        want = ref.mm_utils.process_anyres_image(Image.fromarray(img), proc, gi.PINPOINTS).numpy()
        assert np.array_equal(ro.process_anyres_image(img, gi.PINPOINTS), want), (w, h)


def test_reference_add_faster_video_grid_raises_nameerror():
    """What `add_faster_video` does in the reference (mirrored by radvlm_b200.mm_arch._video_entry): with
    mm_newline_position == "grid" prepare_inputs_labels_for_multimodal reads `all_faster_video_features`, which only the
    commented-out encode_multimodals call would define (llava_arch.py:281, 317) -> NameError; "frame" ignores the flag."""
    from oracle.ref_loader import build_reference_host
    host, _ = build_reference_host(vocab=gi.MERGE_VOCAB, hidden_size=gi.MERGE_HIDDEN, seed=0,
                                   vision_kwargs=dict(hidden_size=16, intermediate_size=16, num_hidden_layers=2,
                                                      num_attention_heads=1))
    case = gi.video_cases()["bilinear_grid"]
    feats = gi.merge_features(case)
    host.encode_images = lambda images, _f=feats: _f
    host.config.mm_spatial_pool_mode, host.config.mm_spatial_pool_stride = "bilinear", 2
    host.config.mm_patch_merge_type, host.config.add_faster_video = "spatial_unpad", True
    images = [torch.zeros(n, 3, 2, 2) for n in case["tiles"]]
    ids, mask, labels = gi.merge_ids(case)
    pos = torch.arange(ids.shape[1])[None].expand(ids.shape[0], -1).contiguous()
    host.config.mm_newline_position = "grid"
    with pytest.raises(NameError, match="all_faster_video_features"):
        host.prepare_inputs_labels_for_multimodal(ids, pos, mask, None, labels, images, modalities=["video"],
                                                  image_sizes=case["sizes"])
    host.config.mm_newline_position = "frame"
    out = host.prepare_inputs_labels_for_multimodal(ids, pos, mask, None, labels, images, modalities=["video"],
                                                    image_sizes=case["sizes"])
    host.config.add_faster_video = False
    ref = host.prepare_inputs_labels_for_multimodal(ids, pos, mask, None, labels, images, modalities=["video"],
                                                    image_sizes=case["sizes"])
    assert torch.equal(out[4], ref[4])
