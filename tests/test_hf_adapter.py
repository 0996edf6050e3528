"""HF-format adapter (radvlm_b200.hf_adapter): wiring on CPU against a tiny transformers LlavaOnevision model, and on the
GPU the packed features against transformers' own get_image_features / pack_image_features (an independent
implementation of the same algorithm)."""
import pytest
import torch

import golden_inputs as gi


def _tiny_hf(dtype=torch.float32, device="cpu"):
    transformers = pytest.importorskip("transformers")
    from transformers import LlavaOnevisionConfig, LlavaOnevisionModel
    vision = dict(model_type="siglip_vision_model", hidden_size=144, intermediate_size=272, num_hidden_layers=3,
                  num_attention_heads=2, image_size=384, patch_size=14, hidden_act="gelu_pytorch_tanh",
                  layer_norm_eps=1e-6, vision_use_head=False)
    text = dict(model_type="qwen2", hidden_size=256, intermediate_size=128, num_hidden_layers=1, num_attention_heads=4,
                num_key_value_heads=4, vocab_size=320)
    cfg = LlavaOnevisionConfig(vision_config=vision, text_config=text, image_grid_pinpoints=gi.PINPOINTS,
                               vision_feature_layer=-1, vision_feature_select_strategy="full",
                               vision_aspect_ratio="anyres_max_9", projector_hidden_act="gelu")
    torch.manual_seed(0)
    model = LlavaOnevisionModel(cfg).to(device=device, dtype=dtype).eval()
    with torch.no_grad():
        model.image_newline.normal_(0, 0.02)
    return model


def test_adapter_wiring_cpu():
    from radvlm_b200.hf_adapter import B200OnevisionFeatures
    hf = _tiny_hf()
    ad = B200OnevisionFeatures(hf)
    # the projector view shares the HF Parameters (nothing copied) under the names PackedWeights reads
    sd = ad._projector.state_dict()
    assert sd["0.weight"].data_ptr() == hf.multi_modal_projector.linear_1.weight.data_ptr()
    assert sd["2.bias"].data_ptr() == hf.multi_modal_projector.linear_2.bias.data_ptr()
    assert "vision_model.encoder.layers.2.self_attn.q_proj.weight" in hf.vision_tower.state_dict()
    # HF (H, W) sizes -> tile counts identical to transformers' image_size_to_num_patches
    from transformers.models.llava_onevision.modeling_llava_onevision import image_size_to_num_patches
    sizes = [(300, 500), (1024, 1024), (384, 384), (3056, 2544)]
    want = [image_size_to_num_patches(s, gi.PINPOINTS, 384) for s in sizes]
    assert ad.image_num_patches(sizes) == want
    assert ad.image_num_patches(sizes[:2], batch_num_images=[2]) == [1, 1]
    with pytest.raises(NotImplementedError):
        ad.get_image_features(torch.zeros(1, 3, 384, 384), [(384, 384)], vision_feature_layer=-2)
    with pytest.raises(TypeError):
        B200OnevisionFeatures(torch.nn.Linear(2, 2))


@pytest.mark.gpu
def test_adapter_matches_transformers_on_gpu():
    from radvlm_b200 import mm_utils
    from radvlm_b200.hf_adapter import B200OnevisionFeatures
    hf = _tiny_hf(dtype=torch.float32, device="cuda")
    ad = B200OnevisionFeatures(hf)
    names = ["rgb_500x300_noise", "exact_384_noise"]
    imgs = [torch.from_numpy(gi.preprocess_image(gi.preprocess_cases()[n])) for n in names]
    big = torch.from_numpy(gi.grad_image("grad_1536_gray_mix"))              # 4x4 grid: exercises the bilinear pooling
    tiles, sizes_wh, splits, _ = mm_utils.preprocess_anyres_batch(imgs + [big], gi.PINPOINTS, dtype=torch.float32)
    sizes_hw = [(h, w) for (w, h) in sizes_wh]
    got, lens = ad.get_image_features(tiles, sizes_hw)
    with torch.no_grad():
        ref = hf.get_image_features(tiles, torch.tensor(sizes_hw), vision_feature_layer=-1,
                                    vision_feature_select_strategy="full").pooler_output
    ref_cat = torch.cat(list(ref), dim=0) if isinstance(ref, (list, tuple)) else ref
    assert lens == [int(r.shape[0]) for r in ref] if isinstance(ref, (list, tuple)) else sum(lens) == ref_cat.shape[0]
    assert got.shape == ref_cat.shape
    g, r = got.double().flatten().cpu(), ref_cat.double().flatten().cpu()
    cos = float(torch.dot(g, r) / (g.norm() * r.norm()))
    relmax = float((g - r).abs().max() / r.abs().max())
    assert cos >= 0.999 and relmax <= 2e-2, "HF parity: cos=%.6f relmax=%.3e" % (cos, relmax)
