"""Generate the committed golden vectors by running the REAL reference (rfahrn/RadVLM) in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.json / *.npz

Needs ``/root/reference`` (read-only) — it cannot run on the GPU box; the fixtures it writes can.
Everything is seeded; inputs are regenerated from the recorded recipes by ``tests/golden_inputs.py`` so the
fixtures only store reference OUTPUTS (plus small inputs where convenient).

Reference entry points exercised (unmodified code):
  mm_utils.select_best_resolution / get_anyres_image_grid_shape / resize_and_pad_image / process_anyres_image
  llava_arch.unpad_image / LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal / encode_images
  siglip_encoder.SigLipVisionTower.forward, multimodal_projector mlp2x_gelu
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_loader import build_reference_host, import_reference  # noqa: E402
import golden_inputs as gi  # noqa: E402


def planner_golden(ref):
    from PIL import Image
    mm, arch = ref.mm_utils, ref.llava_arch
    pins = gi.PINPOINTS
    out = {"pinpoints": pins, "cases": []}
    for (W, H) in gi.planner_sizes():
        best = mm.select_best_resolution((W, H), pins)
        gw, gh = mm.get_anyres_image_grid_shape((W, H), pins, 384)
        case = {"size": [W, H], "best": list(best), "grid": [gw, gh]}
        # unpad window from the real unpad_image on an index tensor
        ch, cw = gh * 27, gw * 27
        idx = torch.arange(ch * cw, dtype=torch.float32).view(1, ch, cw)
        u = arch.unpad_image(idx, (W, H))
        first = int(u[0, 0, 0].item())
        case["unpad"] = [first // cw, first % cw, int(u.shape[1]), int(u.shape[2])]
        if W * H <= 1600 * 1600:
            img = Image.new("RGB", (W, H))
            padded = mm.resize_and_pad_image(img, best)
            assert padded.size == tuple(best)
            # recover the resized size / paste offset with a white image
            white = Image.new("RGB", (W, H), (255, 255, 255))
            arr = np.asarray(mm.resize_and_pad_image(white, best))[:, :, 0]
            ys, xs = np.nonzero(arr)
            case["resize"] = [int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1), int(xs.min()), int(ys.min())]
        out["cases"].append(case)
    return out


def merge_splice_golden():
    """Real prepare_inputs_labels_for_multimodal with encode_images stubbed by deterministic features."""
    C = gi.MERGE_HIDDEN
    host, ref = build_reference_host(vocab=gi.MERGE_VOCAB, hidden_size=C, seed=0,
                                     vision_kwargs=dict(hidden_size=16, intermediate_size=16, num_hidden_layers=2,
                                                        num_attention_heads=1))
    with torch.no_grad():
        host.model.embed_tokens.weight.copy_(gi.merge_embed_table())
        host.model.image_newline.copy_(gi.merge_newline())
    results = {}
    for name, case in gi.merge_cases().items():
        host.config.tokenizer_padding_side = case.get("padding_side", "right")
        host.config.tokenizer_model_max_length = case.get("max_length", 32768)
        host.config.image_aspect_ratio = case.get("aspect", "anyres_max_9")
        host.config.mm_patch_merge_type = case.get("merge_type", "spatial_unpad")
        feats = gi.merge_features(case)
        host.encode_images = lambda images, _f=feats: _f  # stub: tower/projector are not under test here
        images = [torch.zeros(n, 3, 2, 2) for n in case["tiles"]]
        ids, mask, labels = gi.merge_ids(case)
        position_ids = torch.arange(ids.shape[1], dtype=torch.long)[None].expand(ids.shape[0], -1).contiguous()
        out = host.prepare_inputs_labels_for_multimodal(
            ids, position_ids, mask, None, labels, images, modalities=["image"] * ids.shape[0], image_sizes=case["sizes"])
        _, pos, am, _, emb, lab = out
        results[name + "/embeds"] = emb.detach().numpy().astype(np.float32)
        results[name + "/labels"] = lab.numpy() if lab is not None else np.zeros(0, dtype=np.int64)
        results[name + "/mask"] = am.numpy().astype(np.uint8) if am is not None else np.zeros(0, dtype=np.uint8)
        results[name + "/pos"] = pos.numpy() if pos is not None else np.zeros(0, dtype=np.int64)
        print("merge/splice", name, tuple(emb.shape))
    return results


def video_golden():
    """Real prepare_inputs_labels_for_multimodal on video samples (get_2dPool, add_token_per_grid / _frame,
    one_token / no_token; llava_arch.py:171-190, 222-250, 286-349), encode_images stubbed as above."""
    C = gi.MERGE_HIDDEN
    host, ref = build_reference_host(vocab=gi.MERGE_VOCAB, hidden_size=C, seed=0,
                                     vision_kwargs=dict(hidden_size=16, intermediate_size=16, num_hidden_layers=2,
                                                        num_attention_heads=1))
    with torch.no_grad():
        host.model.embed_tokens.weight.copy_(gi.merge_embed_table())
        host.model.image_newline.copy_(gi.merge_newline())
    results = {}
    for name, case in gi.video_cases().items():
        host.config.tokenizer_padding_side = "right"
        host.config.tokenizer_model_max_length = 32768
        host.config.image_aspect_ratio = "anyres_max_9"
        host.config.mm_spatial_pool_mode = case["pool"]
        host.config.mm_spatial_pool_stride = 2
        host.config.mm_newline_position = case["newline"]
        host.config.mm_patch_merge_type = case.get("merge_type", "spatial_unpad")
        host.config.add_faster_video = False
        feats = gi.merge_features(case)
        host.encode_images = lambda images, _f=feats: _f
        images = [torch.zeros(n, 3, 2, 2) for n in case["tiles"]]
        ids, mask, labels = gi.merge_ids(case)
        position_ids = torch.arange(ids.shape[1], dtype=torch.long)[None].expand(ids.shape[0], -1).contiguous()
        out = host.prepare_inputs_labels_for_multimodal(
            ids, position_ids, mask, None, labels, images, modalities=case["modalities"], image_sizes=case["sizes"])
        _, pos, am, _, emb, lab = out
        results[name + "/embeds"] = emb.detach().numpy().astype(np.float32)
        results[name + "/labels"] = lab.numpy()
        results[name + "/mask"] = am.numpy().astype(np.uint8)
        results[name + "/pos"] = pos.numpy()
        print("video", name, tuple(emb.shape))
    return results


def processor_golden(ref):
    """Real SigLipImageProcessor.preprocess (siglip_encoder.py:47-67) on three of the preprocessing images."""
    from PIL import Image
    import importlib
    enc = importlib.import_module("llava.model.multimodal_encoder.siglip_encoder")
    proc = enc.SigLipImageProcessor()
    meta = {}
    cases = gi.preprocess_cases()
    for name in ["rgb_500x300_noise", "c2_1024_gray_noise", "small_130x100_mix"]:
        img = gi.preprocess_image(cases[name])
        pil = Image.fromarray(img).convert("RGB")
        out = proc.preprocess(pil, return_tensors="pt")["pixel_values"][0].numpy().astype(np.float32)
        meta[name] = {"shape": list(out.shape), "sha256": hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()}
        print("processor", name, out.shape)
    return meta


def preprocess_golden(ref):
    from PIL import Image
    proc = ref.siglip_encoder.SigLipImageProcessor()
    meta, arrays = {}, {}
    for name, case in gi.preprocess_cases().items():
        img = gi.preprocess_image(case)
        pil = Image.fromarray(img if img.ndim == 3 else np.repeat(img[:, :, None], 3, axis=2))
        out = ref.mm_utils.process_anyres_image(pil, proc, gi.PINPOINTS).numpy()
        assert out.dtype == np.float32
        meta[name] = {"shape": list(out.shape), "sha256": hashlib.sha256(out.tobytes()).hexdigest()}
        if case.get("store", False):  # outputs are exactly LUT[u8]: store the u8 indices (lossless, compressible)
            lut = gi.normalize_lut_f32()
            u8 = np.searchsorted(lut, out).astype(np.uint8)
            assert np.array_equal(lut[u8], out)
            arrays[name + "/u8"] = u8
        else:  # sampled rows: tile 0 row 7, last tile row 200
            arrays[name + "/sample"] = np.stack([out[0, :, 7, :], out[-1, :, 200, :]])
        print("preprocess", name, out.shape, meta[name]["sha256"][:12])
    return meta, arrays


def encoder_golden():
    from radvlm_b200.synthetic import seeded_init_
    res = {}
    # (a) reduced tower: every op of the path at small width (3 executed layers)
    host, ref = build_reference_host(vocab=64, hidden_size=gi.SMALL_PROJ, seed=0, vision_kwargs=gi.SMALL_VISION)
    seeded_init_(host, gi.SMALL_SEED)
    names = sorted(n for n, _ in host.named_parameters())
    res["small/param_names_sha"] = np.frombuffer(hashlib.sha256("\n".join(names).encode()).digest(), dtype=np.uint8)
    x = gi.encoder_pixels(2, seed=11)
    with torch.no_grad():
        tower_out = host.get_vision_tower()(x)
        feat = host.encode_images(x)
    res["small/tower"] = tower_out.numpy()
    res["small/features"] = feat.numpy()
    print("encoder small", tuple(tower_out.shape), tuple(feat.shape))
    # (b) full-size SigLIP-so400m/14-384 (26 executed layers) + 1152->3584->3584 projector, one tile (config 1)
    host, ref = build_reference_host(vocab=64, hidden_size=3584, seed=0)
    seeded_init_(host, gi.FULL_SEED)
    names = sorted(n for n, _ in host.named_parameters())
    res["full/param_names_sha"] = np.frombuffer(hashlib.sha256("\n".join(names).encode()).digest(), dtype=np.uint8)
    x = gi.encoder_pixels(1, seed=12)
    with torch.no_grad():
        tower_out = host.get_vision_tower()(x)
        feat = host.encode_images(x)
    rows = gi.FULL_SAMPLE_ROWS
    res["full/tower_rows"] = tower_out[0, rows].numpy()
    res["full/features_rows"] = feat[0, rows].numpy()
    res["full/tower_stats"] = np.array([tower_out.mean().item(), tower_out.std().item(), tower_out.abs().max().item()])
    res["full/features_stats"] = np.array([feat.mean().item(), feat.std().item(), feat.abs().max().item()])
    print("encoder full", tuple(tower_out.shape), tuple(feat.shape), res["full/tower_stats"], res["full/features_stats"])
    return res


def grad_golden(ref):
    """Gradients of the REAL reference path (fp32, CPU): process_anyres_image -> prepare_inputs_labels_for_multimodal
    (tower, projector, unpad / bilinear pool / newline, splice) -> L = sum(inputs_embeds * R) -> autograd."""
    from PIL import Image
    from radvlm_b200.synthetic import seeded_init_
    host, _ = build_reference_host(vocab=64, hidden_size=gi.SMALL_PROJ, seed=0, vision_kwargs=gi.SMALL_VISION)
    seeded_init_(host, gi.GRAD_SEED)
    host.requires_grad_(True)
    host.train()
    proc = host.get_vision_tower().image_processor
    tiles, sizes = [], []
    for name in gi.GRAD_IMAGES:
        arr = gi.grad_image(name)
        if arr.ndim == 2:
            arr = np.repeat(arr[:, :, None], 3, axis=2)
        img = Image.fromarray(arr)
        tiles.append(ref.mm_utils.process_anyres_image(img, proc, host.config.image_grid_pinpoints))
        sizes.append(img.size)
    L = max(len(r) for r in gi.GRAD_IDS)
    ids = torch.full((len(gi.GRAD_IDS), L), 0, dtype=torch.long)
    mask = torch.zeros(len(gi.GRAD_IDS), L, dtype=torch.bool)
    for b, r in enumerate(gi.GRAD_IDS):
        ids[b, :len(r)] = torch.tensor(r)
        mask[b, :len(r)] = True
    labels = torch.where(ids < 0, torch.full_like(ids, -100), ids)
    pos = torch.arange(L)[None].expand(len(gi.GRAD_IDS), -1).contiguous()
    out = host.prepare_inputs_labels_for_multimodal(ids, pos, mask, None, labels, tiles, ["image"] * len(tiles), sizes)
    emb = out[4]
    R = gi.grad_loss_weights(emb.shape)
    (emb * R).sum().backward()
    rows = np.linspace(0, emb.shape[1] - 1, 64).astype(np.int64)
    res = {"embeds_rows": emb.detach()[:, rows].numpy().astype(np.float32), "rows": rows,
           "embeds_shape": np.array(emb.shape), "tile_counts": np.array([t.shape[0] for t in tiles])}
    n = 0
    for name, p in host.named_parameters():
        if p.grad is None:
            continue
        res["grad/" + name] = p.grad.detach().numpy().astype(np.float32)
        n += 1
    print("grad golden: embeds", tuple(emb.shape), "tiles", [t.shape[0] for t in tiles], n, "parameter gradients")
    return res


def main():
    parts = set(sys.argv[1:]) or {"planner", "merge", "video", "preprocess", "processor", "encoder", "grad"}
    ref = import_reference()
    if "planner" in parts:
        with open(os.path.join(HERE, "planner_golden.json"), "w") as f:
            json.dump(planner_golden(ref), f)
    if "merge" in parts:
        np.savez_compressed(os.path.join(HERE, "merge_splice_golden.npz"), **merge_splice_golden())
    if "video" in parts:
        np.savez_compressed(os.path.join(HERE, "video_golden.npz"), **video_golden())
    if "processor" in parts:
        with open(os.path.join(HERE, "processor_golden.json"), "w") as f:
            json.dump(processor_golden(ref), f, indent=1)
    if "preprocess" in parts:
        meta, arrays = preprocess_golden(ref)
        with open(os.path.join(HERE, "preprocess_golden.json"), "w") as f:
            json.dump(meta, f, indent=1)
        np.savez_compressed(os.path.join(HERE, "preprocess_golden.npz"), **arrays)
    if "encoder" in parts:
        np.savez_compressed(os.path.join(HERE, "encoder_golden.npz"), **encoder_golden())
    if "grad" in parts:
        np.savez_compressed(os.path.join(HERE, "grad_golden.npz"), **grad_golden(ref))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
