"""HF-format entry point: the same sm_100a kernels behind
``transformers.LlavaOnevisionModel.get_image_features`` / ``pack_image_features``
(transformers/models/llava_onevision/modeling_llava_onevision.py), which is what RadVLM's own evaluation loads
(models_loading_inference.py:107-112 via convert_llava_onevision_weights_to_hf.py:49-59).

The HF checkpoint stores the SAME tensors under HF names: ``vision_tower.vision_model.*`` (SigLIP, the dropped 27th layer
already gone), ``multi_modal_projector.linear_1 / linear_2`` (= mm_projector.0 / .2) and ``image_newline``; sizes are
``(H, W)`` instead of the reference's PIL ``(W, H)``.  Nothing is copied: the adapter wraps the HF modules' Parameters.

    feats = B200OnevisionFeatures(hf_model)
    image_features, feature_lens = feats.get_image_features(pixel_values, image_sizes)      # HF semantics
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import mm_arch, planner
from .encoder import B200VisionEncoder


class B200OnevisionFeatures:
    def __init__(self, hf_model):
        core = getattr(hf_model, "model", hf_model)              # ...ForConditionalGeneration -> LlavaOnevisionModel
        for attr in ("vision_tower", "multi_modal_projector", "image_newline", "config"):
            if not hasattr(core, attr):
                raise TypeError("expected a transformers LlavaOnevision model, %s has no %r" % (type(hf_model).__name__, attr))
        self.core = core
        self.config = core.config
        vcfg = core.config.vision_config
        act = getattr(core.config, "projector_hidden_act", "gelu")
        if act != "gelu" or getattr(vcfg, "hidden_act", "gelu_pytorch_tanh") != "gelu_pytorch_tanh":
            raise NotImplementedError("radvlm_b200 implements SigLIP (gelu_pytorch_tanh) + an erf-GELU mlp2x projector")
        proj = core.multi_modal_projector
        # nn.Sequential(Linear, GELU, Linear) view over the HF projector's own Parameters (builder.py:41-48 naming)
        self._projector = nn.Sequential(proj.linear_1, nn.GELU(), proj.linear_2)
        self.encoder = B200VisionEncoder(core.vision_tower, self._projector, num_heads=vcfg.num_attention_heads,
                                         image_size=vcfg.image_size, ln_eps=vcfg.layer_norm_eps)
        self.patches_per_side = vcfg.image_size // vcfg.patch_size

    # -- shim with the attribute names the reference-side merge code reads (llava_arch.py:350-413)
    def _shim(self, vision_aspect_ratio: str):
        vcfg = self.config.vision_config
        tower = SimpleNamespace(num_patches_per_side=self.patches_per_side, image_size=vcfg.image_size)
        model = SimpleNamespace(image_newline=self.core.image_newline)
        cfg = SimpleNamespace(mm_patch_merge_type="spatial_unpad", image_aspect_ratio=vision_aspect_ratio,
                              image_grid_pinpoints=[list(p) for p in self.config.image_grid_pinpoints])
        return SimpleNamespace(config=cfg, get_vision_tower=lambda: tower, get_model=lambda: model)

    def image_num_patches(self, image_sizes_hw: Sequence, batch_num_images: Optional[Sequence[int]] = None) -> List[int]:
        """HF ``image_size_to_num_patches`` per image ((H, W) sizes); multi-image samples get one tile per image."""
        vcfg = self.config.vision_config
        if batch_num_images is None:
            need = [True] * len(image_sizes_hw)
        else:
            need = [int(n) == 1 for n in batch_num_images for _ in range(int(n))]
        out = []
        for (h, w), patch in zip(image_sizes_hw, need):
            out.append(planner.plan_image((int(w), int(h)), self.config.image_grid_pinpoints, vcfg.image_size,
                                          self.patches_per_side, 0).n_tiles if patch else 1)
        return out

    def get_image_features(self, pixel_values: torch.Tensor, image_sizes, vision_feature_layer: int = -1,
                           vision_feature_select_strategy: str = "full", vision_aspect_ratio: str = "anyres_max_9",
                           batch_num_images=None) -> Tuple[torch.Tensor, List[int]]:
        """Returns ``(image_features [sum(feature_lens), embed_dim], feature_lens)`` exactly as
        ``LlavaOnevisionModel.get_image_features(...).pooler_output`` / ``pack_image_features`` do."""
        n_layers = len(self.core.vision_tower.vision_model.encoder.layers)
        if vision_feature_layer not in (-1, n_layers):
            raise NotImplementedError("radvlm_b200 returns hidden_states[-1] (RadVLM's setting); got layer %r" % (vision_feature_layer,))
        if vision_feature_select_strategy != "full":
            raise NotImplementedError("SigLIP has no CLS token: vision_feature_select_strategy must be 'full'")
        sizes_hw = [(int(s[0]), int(s[1])) for s in (image_sizes.tolist() if isinstance(image_sizes, torch.Tensor) else image_sizes)]
        counts = self.image_num_patches(sizes_hw, None if batch_num_images is None else
                                        [int(v) for v in (batch_num_images.tolist() if isinstance(batch_num_images, torch.Tensor) else batch_num_images)])
        if pixel_values.dim() == 5:
            pixel_values = torch.cat([pv[:n] for pv, n in zip(pixel_values, counts)], dim=0)
        elif pixel_values.dim() != 4:
            raise ValueError(f"pixel_values of shape {pixel_values.shape}, expect to be of 4 or 5 dimensions")
        feats = self.encoder.encode_images(pixel_values)                      # [sum(counts), T, embed_dim]
        return self.pack_image_features(feats, sizes_hw, counts, vision_aspect_ratio)

    def pack_image_features(self, features: torch.Tensor, image_sizes_hw: Sequence, tile_counts: Sequence[int],
                            vision_aspect_ratio: str = "anyres_max_9") -> Tuple[torch.Tensor, List[int]]:
        sizes_wh = [(int(w), int(h)) for (h, w) in image_sizes_hw]          # the merge code speaks PIL (W, H)
        shim = self._shim(vision_aspect_ratio)
        _, tokens = mm_arch._merge_table(shim, list(tile_counts), sizes_wh, False)
        packed = mm_arch.merge_images(shim, features, list(tile_counts), sizes_wh)
        return packed, [int(t) for t in tokens]
