"""The CPU oracle (oracle/*.py) against the golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  This is what pins the oracle; the CUDA path is then compared with the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import encoder_oracle as eo
from oracle import planner_oracle as po
from oracle import resample_oracle as ro


# ------------------------------------------------------------------------------------------ preprocessing
def _pre_meta(golden_dir):
    with open(os.path.join(golden_dir, "preprocess_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(gi.preprocess_cases()))
def test_resample_oracle_bit_exact_vs_reference(golden_dir, name):
    case = gi.preprocess_cases()[name]
    if case["size"][0] * case["size"][1] > 3000 * 1000 + 1:
        pytest.skip("largest case is checked in test_resample_oracle_large (slow numpy loop)")
    meta = _pre_meta(golden_dir)[name]
    out = ro.process_anyres_image(gi.preprocess_image(case), gi.PINPOINTS)
    assert list(out.shape) == meta["shape"] and out.dtype == np.float32
    assert hashlib.sha256(out.tobytes()).hexdigest() == meta["sha256"]
    z = np.load(os.path.join(golden_dir, "preprocess_golden.npz"))
    if name + "/u8" in z:
        assert np.array_equal(gi.normalize_lut_f32()[z[name + "/u8"]], out)
    else:
        assert np.array_equal(np.stack([out[0, :, 7, :], out[-1, :, 200, :]]), z[name + "/sample"])


def test_resample_oracle_large(golden_dir):
    name = "mimic_2544x3056_mix"
    meta = _pre_meta(golden_dir)[name]
    out = ro.process_anyres_image(gi.preprocess_image(gi.preprocess_cases()[name]), gi.PINPOINTS)
    assert hashlib.sha256(out.tobytes()).hexdigest() == meta["sha256"]


def test_normalize_lut_values():
    lut = ro.normalize_lut()
    assert lut[0] == -1.0 and lut[255] == 1.0 and np.all(np.diff(lut) > 0)
    assert np.array_equal(lut, gi.normalize_lut_f32())


def test_siglip_image_processor_oracle_vs_reference(golden_dir):
    """SigLipImageProcessor.preprocess (siglip_encoder.py:47-67) restated in the oracle == the real reference (sha256)."""
    import hashlib
    import json
    meta = json.load(open(os.path.join(golden_dir, "processor_golden.json")))
    cases = gi.preprocess_cases()
    for name, m in meta.items():
        out = ro.siglip_image_processor(gi.preprocess_image(cases[name]))
        assert list(out.shape) == m["shape"]
        assert hashlib.sha256(out.tobytes()).hexdigest() == m["sha256"], name


# ------------------------------------------------------------------------------------------ merge + splice
@pytest.mark.parametrize("name", sorted(gi.merge_cases()))
def test_merge_splice_oracle_vs_reference(golden_dir, name):
    case = gi.merge_cases()[name]
    z = np.load(os.path.join(golden_dir, "merge_splice_golden.npz"))
    feats = gi.merge_features(case)
    newline = gi.merge_newline()
    aspect = case.get("aspect", "anyres_max_9")
    mx = 9 if aspect == "anyres_max_9" else None
    per_image, base = [], 0
    for n, size in zip(case["tiles"], case["sizes"]):
        per_image.append(eo.merge_image(feats[base:base + n], size, newline, gi.PINPOINTS, max_num_patches=mx,
                                        merge_type=case.get("merge_type", "spatial_unpad"), anyres="anyres" in aspect))
        base += n
    ids, mask, labels = gi.merge_ids(case)
    emb, lab, am, pos = eo.prepare_inputs_labels(gi.merge_embed_table(), per_image, ids, mask, labels,
                                                 case.get("max_length", 32768), case.get("padding_side") == "left")
    assert tuple(emb.shape) == z[name + "/embeds"].shape
    assert np.array_equal(lab.numpy(), z[name + "/labels"])
    assert np.array_equal(am.numpy().astype(np.uint8), z[name + "/mask"])
    assert np.array_equal(pos.numpy(), z[name + "/pos"])
    assert np.array_equal(emb.numpy(), z[name + "/embeds"])  # same torch ops as the reference -> bit-exact


@pytest.mark.parametrize("name", sorted(gi.video_cases()))
def test_video_merge_oracle_vs_reference(golden_dir, name):
    """Video / get_2dPool branch (llava_arch.py:171-190, 222-250, 286-349) against the real reference's outputs."""
    case = gi.video_cases()[name]
    z = np.load(os.path.join(golden_dir, "video_golden.npz"))
    feats = gi.merge_features(case)
    newline = gi.merge_newline()
    per_image, base = [], 0
    for i, (n, size) in enumerate(zip(case["tiles"], case["sizes"])):
        f = feats[base:base + n]
        if case["modalities"][i] == "video":
            per_image.append(eo.merge_video(f, newline, case["pool"], case["newline"], case.get("merge_type", "spatial_unpad")))
        else:
            per_image.append(eo.merge_image(f, size, newline, gi.PINPOINTS, max_num_patches=9))
        base += n
    ids, mask, labels = gi.merge_ids(case)
    emb, lab, am, pos = eo.prepare_inputs_labels(gi.merge_embed_table(), per_image, ids, mask, labels, 32768, False)
    assert tuple(emb.shape) == z[name + "/embeds"].shape
    assert np.array_equal(lab.numpy(), z[name + "/labels"])
    assert np.array_equal(am.numpy().astype(np.uint8), z[name + "/mask"])
    assert np.array_equal(pos.numpy(), z[name + "/pos"])
    assert np.array_equal(emb.numpy(), z[name + "/embeds"])  # same torch ops as the reference -> bit-exact


def test_merge_source_map_closed_form():
    """The closed-form per-token source map (what the CUDA gather implements) == the tensor-op merge."""
    newline = gi.merge_newline()
    for size, tiles in [((1024, 1024), 10), ((800, 1200), 13), ((500, 300), 3), ((384, 384), 2)]:
        g = torch.Generator().manual_seed(5)
        feats = torch.randn(tiles, 729, 8, generator=g)
        merged = eo.merge_image(feats, size, newline, gi.PINPOINTS)
        plan = po.plan_image(size, gi.PINPOINTS)
        assert merged.shape[0] == plan["n_tokens"]
        for t in list(range(0, plan["n_tokens"], 97)) + [plan["n_tokens"] - 1, 729, 728]:
            src = po.merge_source(plan, t)
            want = newline if src[0] == "newline" else feats[src[1], src[2]]
            assert torch.equal(merged[t], want), (size, t, src)


# ------------------------------------------------------------------------------------------ tower + projector
def _host_state(vision_cfg, proj_hidden, seed):
    from radvlm_b200 import synthetic
    host = synthetic.build_host(hidden_size=proj_hidden, vocab=64, seed=seed, dtype=torch.float32, device="cpu",
                                vision_cfg=vision_cfg)
    names = sorted(n for n, _ in host.named_parameters())
    sha = np.frombuffer(hashlib.sha256("\n".join(names).encode()).digest(), dtype=np.uint8)
    return host, sha


def test_encoder_oracle_small_vs_reference(golden_dir):
    from radvlm_b200 import synthetic
    z = np.load(os.path.join(golden_dir, "encoder_golden.npz"))
    v = dict(gi.SMALL_VISION)
    v["num_hidden_layers"] -= 1  # load_model drops the last layer (siglip_encoder.py:570)
    host, sha = _host_state(synthetic.siglip_config(**v), gi.SMALL_PROJ, gi.SMALL_SEED)
    assert np.array_equal(sha, z["small/param_names_sha"]), "synthetic module tree must mirror the reference's names"
    x = gi.encoder_pixels(2, seed=11)
    tsd = host.model.vision_tower.vision_tower.state_dict()
    psd = host.model.mm_projector.state_dict()
    tower = eo.tower_forward(tsd, x, num_heads=2)
    feat = eo.projector_forward(psd, tower)
    np.testing.assert_allclose(tower.numpy(), z["small/tower"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(feat.numpy(), z["small/features"], rtol=0, atol=2e-5)


def test_encoder_oracle_full_vs_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "encoder_golden.npz"))
    host, sha = _host_state(None, 3584, gi.FULL_SEED)
    assert np.array_equal(sha, z["full/param_names_sha"])
    x = gi.encoder_pixels(1, seed=12)
    tower = eo.tower_forward(host.model.vision_tower.vision_tower.state_dict(), x)
    feat = eo.projector_forward(host.model.mm_projector.state_dict(), tower)
    rows = gi.FULL_SAMPLE_ROWS
    np.testing.assert_allclose(tower[0, rows].numpy(), z["full/tower_rows"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(feat[0, rows].numpy(), z["full/features_rows"], rtol=0, atol=1e-4)
    st = np.array([tower.mean().item(), tower.std().item(), tower.abs().max().item()])
    np.testing.assert_allclose(st, z["full/tower_stats"], rtol=1e-4, atol=1e-5)
