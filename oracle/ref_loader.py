"""TEST INFRASTRUCTURE ONLY — loader for the *real* reference (rfahrn/RadVLM, finetuning/llava).

Only usable in the build container where ``/root/reference`` exists; it never travels to the GPU
box.  It is used by ``tests/golden/make_golden.py`` to generate the committed golden vectors and by
the ``not gpu`` tests that cross-check the restated oracle (``oracle/*.py``) against the reference.

The recipe follows SURVEY.md §8(c):
  * put ``/root/reference/finetuning`` on ``sys.path``;
  * stub ``llava.model.multimodal_resampler.qformer`` (its imports were removed in transformers 5.x;
    Q-Former is out of scope);
  * never instantiate ``LlavaQwenForCausalLM`` — host the reference mixins
    (``LlavaMetaModel`` / ``LlavaMetaForCausalLM``, llava_arch.py:34-124,162-555) on a minimal module;
  * emulate ``SigLipVisionTower.load_model`` offline (siglip_encoder.py:563-574): build
    ``SigLipVisionModel(SigLipVisionConfig())``, drop the last layer, replace the head by Identity.

Nothing under the product package may import this module.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("RADVLM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "finetuning", "llava"))


def import_reference():
    """Import the reference ``llava`` package (unmodified) and return the modules we need."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected at %s)" % REFERENCE_ROOT)
    ft = os.path.join(REFERENCE_ROOT, "finetuning")
    if ft not in sys.path:
        sys.path.insert(0, ft)
    name = "llava.model.multimodal_resampler.qformer"
    if name not in sys.modules:
        stub = types.ModuleType(name)

        class Qformer:  # pragma: no cover - placeholder, never instantiated
            def __init__(self, *a, **k):
                raise RuntimeError("Q-Former is out of scope (stub)")

        stub.Qformer = Qformer
        sys.modules[name] = stub
    import llava  # noqa: F401
    from llava import mm_utils
    from llava.model import llava_arch
    from llava.model.multimodal_encoder import siglip_encoder
    from llava.model.multimodal_projector import builder as projector_builder

    return SimpleNamespace(mm_utils=mm_utils, llava_arch=llava_arch, siglip_encoder=siglip_encoder,
                           projector_builder=projector_builder)


def radvlm_config(hidden_size=3584, vision_kwargs=None):
    """The config attributes the hot path reads, RadVLM values."""
    return SimpleNamespace(
        mm_vision_tower="google/siglip-so400m-patch14-384",
        mm_projector_type="mlp2x_gelu",
        mm_hidden_size=(vision_kwargs or {}).get("hidden_size", 1152),
        hidden_size=hidden_size,
        mm_patch_merge_type="spatial_unpad",
        image_aspect_ratio="anyres_max_9",
        image_grid_pinpoints=[[384 * i, 384 * j] for i in range(1, 7) for j in range(1, 7)],
        mm_vision_select_layer=-2,
        mm_vision_select_feature="patch",
        mm_newline_position="grid",
        tokenizer_padding_side="right",
        tokenizer_model_max_length=32768,
        mm_use_im_start_end=False,
        use_pos_skipping=False,
        mm_tunable_parts="mm_vision_tower,mm_mlp_adapter,mm_language_model",
        vision_tower_pretrained=None,
        delay_load=True,
        use_mm_proj=True,
    )


def build_reference_host(vocab=1024, hidden_size=3584, seed=0, vision_kwargs=None, dtype=None):
    """Build ``Host(nn.Module, LlavaMetaForCausalLM)`` around the unmodified reference mixins.

    vision_kwargs: optional overrides for ``SigLipVisionConfig`` (reduced-size fixtures).
    Returns (host, ref_modules).
    """
    import torch
    import torch.nn as nn

    ref = import_reference()
    LlavaMetaModel = ref.llava_arch.LlavaMetaModel
    LlavaMetaForCausalLM = ref.llava_arch.LlavaMetaForCausalLM
    sig = ref.siglip_encoder

    cfg = radvlm_config(hidden_size=hidden_size, vision_kwargs=vision_kwargs)

    class _Base(nn.Module):
        def __init__(self, config):
            super().__init__()
            self.config = config
            self.embed_tokens = nn.Embedding(vocab, config.hidden_size)

        @property
        def dtype(self):
            return self.embed_tokens.weight.dtype

    class HostModel(LlavaMetaModel, _Base):
        pass

    class Host(nn.Module, LlavaMetaForCausalLM):
        def __init__(self, config):
            super().__init__()
            self.config = config
            self.model = HostModel(config)

        def get_model(self):
            return self.model

        @property
        def device(self):
            return self.model.embed_tokens.weight.device

    torch.manual_seed(seed)
    # SigLipVisionTower.__init__ calls from_pretrained (network) when mm_tunable_parts names the tower
    # (siglip_encoder.py:557-562); construct with it blanked, restore afterwards.
    tunable = cfg.mm_tunable_parts
    cfg.mm_tunable_parts = ""
    host = Host(cfg)
    cfg.mm_tunable_parts = tunable
    tower = host.get_vision_tower()
    # offline emulation of SigLipVisionTower.load_model (siglip_encoder.py:563-574)
    vcfg = sig.SigLipVisionConfig(**(vision_kwargs or {}))
    tower.config = vcfg
    vm = sig.SigLipVisionModel(vcfg)
    del vm.vision_model.encoder.layers[-1:]
    vm.vision_model.head = nn.Identity()
    vm.requires_grad_(False)
    tower.vision_tower = vm
    tower.is_loaded = True
    # image_newline is torch.empty at construction (llava_arch.py:46)
    with torch.no_grad():
        host.model.image_newline.copy_(torch.randn(cfg.hidden_size) * 0.02)
    if dtype is not None:
        host.to(dtype)
    host.eval()
    return host, ref
