"""Random-init parameter containers with the reference's module tree and state-dict names.

There is no network for checkpoints, so tests / bench / smoke use random-init weights of the real
architecture (BASELINE.json configs).  These classes hold PARAMETERS ONLY — they deliberately have no
``forward``: the only implementation of the math in the product package is the sm_100a library.

Names follow SURVEY.md section 5 (``vision_model.embeddings.patch_embedding.weight`` ...,
``mm_projector.0.weight``, ``image_newline``, ``embed_tokens.weight``), i.e. the reference's
siglip_encoder.py:148-305,408-456 / builder.py:41-48 / llava_arch.py:34-46 module structure.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn

from .mm_arch import B200LlavaMetaForCausalLM


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))


class _Mlp(nn.Module):
    def __init__(self, d, i):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(d, i), nn.Linear(i, d)


class _Layer(nn.Module):
    def __init__(self, d, i, eps):
        super().__init__()
        self.self_attn = _Attn(d)
        self.layer_norm1 = nn.LayerNorm(d, eps=eps)
        self.mlp = _Mlp(d, i)
        self.layer_norm2 = nn.LayerNorm(d, eps=eps)


class _Embeddings(nn.Module):
    def __init__(self, d, c, ps, n_pos):
        super().__init__()
        self.patch_embedding = nn.Conv2d(c, d, kernel_size=ps, stride=ps, padding="valid")
        self.position_embedding = nn.Embedding(n_pos, d)


class _Encoder(nn.Module):
    def __init__(self, d, i, n, eps):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(d, i, eps) for _ in range(n)])


class _VisionTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        n_pos = (cfg.image_size // cfg.patch_size) ** 2
        self.embeddings = _Embeddings(cfg.hidden_size, cfg.num_channels, cfg.patch_size, n_pos)
        self.encoder = _Encoder(cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.layer_norm_eps)
        self.post_layernorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)  # executed-but-discarded in the reference


class SyntheticSiglipVisionModel(nn.Module):
    """Parameter container == reference ``SigLipVisionModel`` after ``load_model`` dropped the last layer."""

    def __init__(self, cfg):
        super().__init__()
        self.vision_model = _VisionTransformer(cfg)


def siglip_config(hidden_size=1152, intermediate_size=4304, num_hidden_layers=26, num_attention_heads=16,
                  num_channels=3, image_size=384, patch_size=14, layer_norm_eps=1e-6):
    """Executed configuration: 26 layers (27 in SigLipVisionConfig minus the one load_model deletes)."""
    return SimpleNamespace(**locals())


class SyntheticVisionTower(nn.Module):
    """Stand-in for ``SigLipVisionTower`` (siglip_encoder.py:538-620): same properties, parameters only."""

    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.vision_tower = SyntheticSiglipVisionModel(cfg)
        self.is_loaded = True

    @property
    def image_processor(self):
        """siglip_encoder.py:546: the tower carries its SigLipImageProcessor (here: the GPU mirror of it)."""
        from .mm_utils import SigLipImageProcessor
        s = self.config.image_size
        return SigLipImageProcessor(size=(s, s), crop_size={"height": s, "width": s})

    def load_model(self, device_map=None):
        """siglip_encoder.py:559-574 loads and truncates the checkpoint; the synthetic tower is always 'loaded'."""
        self.is_loaded = True

    @property
    def num_patches_per_side(self):
        return self.config.image_size // self.config.patch_size

    @property
    def num_patches(self):
        return self.num_patches_per_side ** 2

    @property
    def image_size(self):
        return self.config.image_size

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def dtype(self):
        return next(self.vision_tower.parameters()).dtype

    @property
    def device(self):
        return next(self.vision_tower.parameters()).device


class _HostModel(nn.Module):
    def __init__(self, cfg, vcfg, vocab):
        super().__init__()
        self.config = cfg
        self.embed_tokens = nn.Embedding(vocab, cfg.hidden_size)
        self.vision_tower = SyntheticVisionTower(vcfg)
        self.mm_projector = nn.Sequential(nn.Linear(vcfg.hidden_size, cfg.hidden_size), nn.GELU(),
                                          nn.Linear(cfg.hidden_size, cfg.hidden_size))
        self.image_newline = nn.Parameter(torch.empty(cfg.hidden_size))

    def get_vision_tower(self):
        return self.vision_tower


class SyntheticLlavaHost(B200LlavaMetaForCausalLM, nn.Module):
    """Minimal causal-LM host (no decoder): ``.model`` / ``get_model()`` / ``config`` like LlavaQwenForCausalLM."""

    def __init__(self, cfg, vcfg, vocab):
        super().__init__()
        self.config = cfg
        self.model = _HostModel(cfg, vcfg, vocab)

    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.model.get_vision_tower()

    @property
    def device(self):
        return self.model.embed_tokens.weight.device


def radvlm_config(hidden_size=3584):
    return SimpleNamespace(
        mm_vision_tower="google/siglip-so400m-patch14-384", mm_projector_type="mlp2x_gelu",
        mm_hidden_size=1152, hidden_size=hidden_size, mm_patch_merge_type="spatial_unpad",
        image_aspect_ratio="anyres_max_9",
        image_grid_pinpoints=[[384 * i, 384 * j] for i in range(1, 7) for j in range(1, 7)],
        mm_newline_position="grid", tokenizer_padding_side="right", tokenizer_model_max_length=32768,
        mm_use_im_start_end=False, use_pos_skipping=False)


@torch.no_grad()
def seeded_init_(module: nn.Module, seed: int = 0):
    """Deterministic init in parameter-name order (independent of construction order / torch RNG state).

    Matrices ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default; LayerNorm affine is perturbed
    (gamma 1 +- 0.1, beta +- 0.1) so that it is exercised; embeddings / newline ~ N(0, 0.02^2)-ish.
    """
    g = torch.Generator(device="cpu")
    for idx, (name, p) in enumerate(sorted(module.named_parameters(), key=lambda kv: kv[0])):
        g.manual_seed(seed * 1000003 + idx)
        shape = tuple(p.shape)
        if "layer_norm" in name or "layernorm" in name:
            if name.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                v = 0.1 * torch.randn(shape, generator=g)
        elif "position_embedding" in name:
            v = 0.5 * torch.randn(shape, generator=g)
        elif "embed_tokens" in name or "image_newline" in name:
            v = 0.02 * torch.randn(shape, generator=g)
        elif p.dim() >= 2:
            fan_in = int(torch.tensor(shape[1:]).prod())
            v = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        p.copy_(v.to(p.dtype))
    return module


def build_host(hidden_size=3584, vocab=4096, seed=0, dtype=torch.bfloat16, device="cuda", vision_cfg=None):
    """Random-init RadVLM-shaped host: SigLIP-so400m/14-384 (26 executed layers) + mlp2x_gelu + embed table."""
    cfg = radvlm_config(hidden_size)
    vcfg = vision_cfg or siglip_config()
    with torch.device("cpu"):
        host = SyntheticLlavaHost(cfg, vcfg, vocab)
    seeded_init_(host, seed)
    host.requires_grad_(False)
    host.eval()
    return host.to(device=device, dtype=dtype)
