"""Host planner (C++ behind the C ABI) vs the reference's golden vectors and the pure-Python oracle.
Bit-exact: grid selection, resize geometry, unpad window, pooled size, token counts, splice layout."""
import json
import os

import numpy as np
import pytest

import golden_inputs as gi
from oracle import planner_oracle as po
from radvlm_b200 import _lib, planner


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "planner_golden.json")) as f:
        return json.load(f)


def test_planner_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    pins = g["pinpoints"]
    assert len(g["cases"]) >= 350
    for c in g["cases"]:
        W, H = c["size"]
        assert list(planner.select_best_resolution((W, H), pins)) == c["best"]
        assert list(planner.get_anyres_image_grid_shape((W, H), pins, 384)) == c["grid"]
        p = planner.plan_image((W, H), pins)
        assert [p.crop_r0, p.crop_c0, p.crop_h, p.crop_w] == c["unpad"], (W, H)
        if "resize" in c:
            assert [p.resized_w, p.resized_h, p.paste_x, p.paste_y] == c["resize"], (W, H)


def test_oracle_planner_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    pins = g["pinpoints"]
    for c in g["cases"]:
        W, H = c["size"]
        p = po.plan_image((W, H), pins)
        assert [p["best_w"], p["best_h"]] == c["best"]
        assert [p["grid_w"], p["grid_h"]] == c["grid"]
        assert [p["crop_r0"], p["crop_c0"], p["crop_h"], p["crop_w"]] == c["unpad"]
        if "resize" in c:
            assert [p["resized_w"], p["resized_h"], p["paste_x"], p["paste_y"]] == c["resize"]


def test_known_token_counts():
    # SURVEY.md section 8(a) A8 [probe] values from the reference
    want = {(1024, 1024): 7371, (800, 1200): 8721, (3000, 1000): 7215, (2544, 3056): 7241, (1536, 1536): 7371}
    for sz, n in want.items():
        assert planner.plan_image(sz, gi.PINPOINTS).n_tokens == n
        assert po.plan_image(sz, gi.PINPOINTS)["n_tokens"] == n
    p = planner.plan_image((3000, 1000), gi.PINPOINTS)
    assert (p.pool, p.out_h, p.out_w) == (1, 46, 140)
    p = planner.plan_image((2544, 3056), gi.PINPOINTS)
    assert (p.n_tiles, p.pool, p.out_h, p.out_w) == (31, 1, 88, 73)


def test_python_float_floordiv_trap():
    # 135 // 1.6666666666666667 == 80.0 in Python although 135 / 1.666... rounds to 81.0
    assert int(135 // 1.6666666666666667) == 80
    rng = np.random.default_rng(7)
    for _ in range(3000):
        W, H = int(rng.integers(32, 6000)), int(rng.integers(32, 6000))
        for mx in (9, 4, 0):
            a = planner.plan_image((W, H), gi.PINPOINTS, max_num_patches=mx).as_dict()
            b = po.plan_image((W, H), gi.PINPOINTS, max_num_patches=mx)
            assert a == b, ((W, H), mx)


def test_grid_pinpoints_string_form():
    s = "(1x1),...,(6x6)"
    assert planner.parse_grid_pinpoints(s, 384) == gi.PINPOINTS
    assert planner.get_anyres_image_grid_shape((1024, 1024), s, 384) == (3, 3)
    assert planner.get_anyres_image_grid_shape((3000, 1000), str(gi.PINPOINTS), 384) == (6, 2)


def _random_batch(rng, B, L, n_img_max=2):
    ids = rng.integers(1, 1000, size=(B, L)).astype(np.int64)
    mask = np.zeros((B, L), dtype=np.uint8)
    for b in range(B):
        n = int(rng.integers(2, L + 1))
        if rng.random() < 0.5:
            mask[b, :n] = 1
        else:
            mask[b, L - n:] = 1          # left padded raw batch
        if rng.random() < 0.2:
            mask[b, rng.integers(0, L)] ^= 1  # a hole in the mask
        k = int(rng.integers(0, n_img_max + 1))
        idx = np.nonzero(mask[b])[0]
        if k and len(idx) >= k:
            ids[b, rng.choice(idx, size=k, replace=False)] = -200
    return ids, mask


def _rows_from_plan(plan, B):
    rows = [[("pad",)] * plan.max_len for _ in range(B)]
    covered = 0
    for s in range(plan.n_segments):
        seg = plan.segments[s]
        b, p0 = divmod(seg.dst_row, plan.max_len)
        for i in range(seg.length):
            if seg.kind == _lib.SEG_TEXT:
                src = int(plan.text_src[seg.src_off + i])
                rows[b][p0 + i] = ("text", src)
            elif seg.kind == _lib.SEG_IMAGE:
                rows[b][p0 + i] = ("image", seg.image, seg.src_off + i)
        covered += seg.length
    return rows, covered


@pytest.mark.parametrize("left_pad", [False, True])
@pytest.mark.parametrize("max_length", [None, 700])
def test_splice_plan_matches_oracle(left_pad, max_length):
    rng = np.random.default_rng(11 + int(left_pad) + (max_length or 0))
    for trial in range(40):
        B, L = int(rng.integers(1, 7)), int(rng.integers(4, 40))
        ids, mask = _random_batch(rng, B, L)
        use_mask = mask if trial % 3 else None
        n_need = 0
        for b in range(B):
            m = mask[b].astype(bool) if use_mask is not None else np.ones(L, bool)
            k = int(((ids[b] == -200) & m).sum())
            n_need += max(k, 1)
        tokens = [int(rng.integers(1, 400)) for _ in range(n_need)]
        plan = planner.plan_splice(ids, use_mask, tokens, max_length, left_pad)
        want = po.splice_layout(ids.tolist(), None if use_mask is None else use_mask.astype(bool).tolist(), tokens,
                                max_length, left_pad)
        assert plan.max_len == want["max_len"]
        assert plan.lengths.tolist() == want["lengths"]
        rows, covered = _rows_from_plan(plan, B)
        assert covered == B * plan.max_len  # segments tile the whole padded output exactly once
        for b in range(B):
            for p, src in enumerate(want["rows"][b]):
                got = rows[b][p]
                if src[0] == "text":
                    assert got == ("text", src[1] * L + src[2])
                else:
                    assert got == src


def test_splice_index_error_semantics():
    ids = np.array([[5, -200, 6, -200, 7]], dtype=np.int64)
    # two placeholders, one image: the second reuses the previous image (llava_arch.py:478-481)
    plan = planner.plan_splice(ids, None, [3], None, False)
    assert plan.max_len == 3 + 2 * 3
    want = po.splice_layout(ids.tolist(), None, [3], None, False)
    assert want["max_len"] == plan.max_len
    # text-only sample with an exhausted image list raises IndexError like the reference's list indexing
    ids2 = np.array([[5, 6, 7]], dtype=np.int64)
    with pytest.raises(IndexError):
        planner.plan_splice(ids2, None, [], None, False)
    with pytest.raises(IndexError):
        po.splice_layout(ids2.tolist(), None, [], None, False)
    # text-only sample consumes one image slot and contributes no visual tokens
    ids3 = np.array([[5, 6, 7], [1, -200, 2]], dtype=np.int64)
    plan = planner.plan_splice(ids3, None, [9, 4], None, False)
    assert plan.lengths.tolist() == [3, 2 + 4]


def test_spatial_merge_types_token_counts_match_oracle():
    """Host logic of the spatial merge types other than RadVLM's spatial_unpad (llava_arch.py:373-404: plain "spatial",
    'maxpool2x2', 'nobase', the fixed 2 x 2 grid of the non-anyres aspect modes): token counts equal the oracle's."""
    import torch
    import golden_inputs as gi
    from oracle import encoder_oracle as eo
    from radvlm_b200 import _lib, mm_arch, synthetic
    host = synthetic.build_host(hidden_size=8, vocab=16, dtype=torch.float32, device="cpu",
                                vision_cfg=synthetic.siglip_config(hidden_size=144, intermediate_size=272,
                                                                   num_hidden_layers=1, num_attention_heads=2))
    newline = torch.zeros(8)
    for name, case in gi.merge_cases().items():
        aspect = case.get("aspect", "anyres_max_9")
        mt = case.get("merge_type", "spatial_unpad")
        host.config.image_aspect_ratio, host.config.mm_patch_merge_type = aspect, mt
        table, tokens = mm_arch._merge_table(host, case["tiles"], case["sizes"], False)
        for i, (n, size) in enumerate(zip(case["tiles"], case["sizes"])):
            want = eo.merge_image(torch.zeros(n, 729, 8), size, newline, gi.PINPOINTS,
                                  max_num_patches=9 if aspect == "anyres_max_9" else None, merge_type=mt,
                                  anyres="anyres" in aspect).shape[0]
            assert tokens[i] == want, (name, i, tokens[i], want)
            if n > 1:
                assert bool(table[i].reserved & _lib.ANYRES_NO_BASE) == ("nobase" in mt)
                assert (table[i].pool == _lib.POOL_MAX) == ("maxpool2x2" in mt)


def test_video_merge_table_token_counts_match_oracle():
    """Host logic of the video / get_2dPool branch (mm_arch._video_entry) on the CPU: descriptor fields and token
    counts for every pooling mode x newline placement x merge type equal the oracle's merged sequence length."""
    import itertools
    import torch
    from oracle import encoder_oracle as eo
    from radvlm_b200 import _lib, mm_arch, synthetic
    host = synthetic.build_host(hidden_size=8, vocab=16, dtype=torch.float32, device="cpu",
                                vision_cfg=synthetic.siglip_config(hidden_size=144, intermediate_size=272,
                                                                   num_hidden_layers=1, num_attention_heads=2))
    newline = torch.zeros(8)
    feat = torch.zeros(3, 729, 8)
    for pool, pos, mt in itertools.product(["bilinear", "average", "max"], ["grid", "frame", "one_token", "no_token"],
                                           ["spatial_unpad", "flat"]):
        host.config.mm_spatial_pool_mode, host.config.mm_newline_position, host.config.mm_patch_merge_type = pool, pos, mt
        table, tokens = mm_arch._merge_table(host, [3, 1], [(384, 384), (384, 384)], False, {0})
        want = eo.merge_video(feat, newline, pool, pos, mt).shape[0]
        assert tokens[0] == want, (pool, pos, mt, tokens[0], want)
        m = table[0]
        assert m.mode == _lib.MERGE_VIDEO and m.grid_w == 3 and m.tile_base == 0
        assert m.out_h == m.out_w == (14 if pool == "bilinear" else 13)
        assert table[1].tile_base == 3 and table[1].mode != _lib.MERGE_VIDEO   # the image after the video
    # add_faster_video: the reference dies with a NameError for mm_newline_position == "grid" (llava_arch.py:317 reads
    # `all_faster_video_features`, defined only by the commented-out call at :281) and ignores the flag everywhere else
    import pytest
    host.config.add_faster_video = True
    host.config.mm_spatial_pool_mode, host.config.mm_patch_merge_type = "bilinear", "spatial_unpad"
    host.config.mm_newline_position = "grid"
    with pytest.raises(NameError, match="all_faster_video_features"):
        mm_arch._merge_table(host, [3], [(384, 384)], False, {0})
    host.config.mm_newline_position = "frame"
    _, tokens = mm_arch._merge_table(host, [3], [(384, 384)], False, {0})
    assert tokens[0] == eo.merge_video(feat, newline, "bilinear", "frame", "spatial_unpad").shape[0]
    host.config.add_faster_video = False
    host.config.mm_spatial_pool_mode = "median"
    with pytest.raises(ValueError):
        mm_arch._merge_table(host, [3], [(384, 384)], False, {0})
