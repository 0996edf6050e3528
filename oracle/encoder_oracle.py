"""TEST INFRASTRUCTURE ONLY (oracle) — plain PyTorch fp32 CPU restatement of the floating-point part of the
reference path.  Never imported by the product package.

  SigLipVisionEmbeddings / Attention / MLP / EncoderLayer / Tower    siglip_encoder.py:148-305,576-589
  mm_projector (mlp2x_gelu: Linear, GELU(erf), Linear)                builder.py:41-48 ; llava_arch.py:192-196
  spatial_unpad + anyres_max merge + image_newline                    llava_arch.py:350-412
  splice / truncate / pad / stack                                     llava_arch.py:428-555

Pinned against the real reference modules in the container (``tests/test_oracle_pinned.py``) and against
``tests/golden/encoder_golden.npz`` (reference outputs on reduced- and full-size towers) everywhere.
All functions take a flat state dict with the reference's parameter names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import planner_oracle as po


def tower_forward(sd: Dict[str, torch.Tensor], pixel_values: torch.Tensor, num_heads: int = 16,
                  eps: float = 1e-6, prefix: str = "vision_model.") -> torch.Tensor:
    """hidden_states[-1] of the truncated tower (pre post_layernorm), fp32.  siglip_encoder.py:576-589."""
    g = lambda n: sd[prefix + n].float()
    x = pixel_values.float()
    pw = g("embeddings.patch_embedding.weight")
    ps = pw.shape[-1]
    x = F.conv2d(x, pw, g("embeddings.patch_embedding.bias"), stride=ps)          # :170 ("valid")
    x = x.flatten(2).transpose(1, 2)                                              # :171
    x = x + g("embeddings.position_embedding.weight")[None]                       # :173
    n_layers = 0
    while (prefix + "encoder.layers.%d.layer_norm1.weight" % n_layers) in sd:
        n_layers += 1
    B, T, D = x.shape
    hd = D // num_heads
    for i in range(n_layers):
        lg = lambda n: g("encoder.layers.%d.%s" % (i, n))
        r = x
        h = F.layer_norm(x, (D,), lg("layer_norm1.weight"), lg("layer_norm1.bias"), eps)  # :287
        q = F.linear(h, lg("self_attn.q_proj.weight"), lg("self_attn.q_proj.bias"))
        k = F.linear(h, lg("self_attn.k_proj.weight"), lg("self_attn.k_proj.bias"))
        v = F.linear(h, lg("self_attn.v_proj.weight"), lg("self_attn.v_proj.bias"))
        q = q.view(B, T, num_heads, hd).transpose(1, 2)
        k = k.view(B, T, num_heads, hd).transpose(1, 2)
        v = v.view(B, T, num_heads, hd).transpose(1, 2)
        w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)                     # :216
        w = F.softmax(w, dim=-1, dtype=torch.float32)                             # :227
        a = torch.matmul(w, v).transpose(1, 2).contiguous().reshape(B, T, D)      # :229-235
        a = F.linear(a, lg("self_attn.out_proj.weight"), lg("self_attn.out_proj.bias"))
        x = r + a                                                                 # :293
        r = x
        h = F.layer_norm(x, (D,), lg("layer_norm2.weight"), lg("layer_norm2.bias"), eps)  # :296
        h = F.linear(h, lg("mlp.fc1.weight"), lg("mlp.fc1.bias"))
        h = F.gelu(h, approximate="tanh")                                         # gelu_pytorch_tanh (:83,247)
        h = F.linear(h, lg("mlp.fc2.weight"), lg("mlp.fc2.bias"))
        x = r + h                                                                 # :298
    return x


def projector_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """mlp2x_gelu (builder.py:44-48): Linear -> GELU (exact erf) -> Linear."""
    h = F.linear(x.float(), sd[prefix + "0.weight"].float(), sd[prefix + "0.bias"].float())
    h = F.gelu(h)
    return F.linear(h, sd[prefix + "2.weight"].float(), sd[prefix + "2.bias"].float())


def encode_images(tower_sd, proj_sd, pixel_values, num_heads=16, eps=1e-6) -> torch.Tensor:
    return projector_forward(proj_sd, tower_forward(tower_sd, pixel_values, num_heads, eps))


def merge_image(feat: torch.Tensor, image_size, newline: torch.Tensor, possible_resolutions=None,
                tile_size: int = 384, unit: int = 27, max_num_patches: Optional[int] = 9,
                merge_type: str = "spatial_unpad", anyres: bool = True) -> torch.Tensor:
    """llava_arch.py:350-412 for one image: [tiles, T, C] -> [N, C].  merge_type is mm_patch_merge_type ("spatial*");
    anyres = image_aspect_ratio is "anyres" / "anyres_max_N" (otherwise the fixed 2 x 2 grid of :373-374);
    max_num_patches = N of a matched "anyres_max_N" (None / 0: the un-pooled 'unpad' branch :393-397)."""
    if feat.shape[0] == 1:                                                         # :407-412
        return torch.cat((feat[0], newline[None]), dim=0) if "unpad" in merge_type else feat[0]
    base, rest = feat[0], feat[1:]
    if anyres:
        if possible_resolutions is None:
            possible_resolutions = po.default_pinpoints(tile_size)
        gw, gh = po.get_anyres_image_grid_shape(image_size, possible_resolutions, tile_size)   # :366
    else:
        gw, gh = 2, 2                                                              # :374
    x = rest.view(gh, gw, unit, unit, -1)                                          # :372
    if "maxpool2x2" in merge_type:                                                 # :376-380
        x = x.permute(4, 0, 2, 1, 3).contiguous().flatten(1, 2).flatten(2, 3)
        x = F.max_pool2d(x, 2)
        x = x.flatten(1, 2).transpose(0, 1)
    elif "unpad" in merge_type:                                                    # :381-397
        x = x.permute(4, 0, 2, 1, 3).contiguous().flatten(1, 2).flatten(2, 3)     # :383-384
        r0, r1, c0, c1 = po.unpad_window(image_size, x.shape[1], x.shape[2])      # unpad_image :127-159
        x = x[:, r0:r1, c0:c1]
        c, h, w = x.shape
        if max_num_patches and anyres:
            times = math.sqrt(h * w / (max_num_patches * unit ** 2))               # :387
            if times > 1.1:
                x = F.interpolate(x[None], [int(h // times), int(w // times)], mode="bilinear")[0]  # :390
        x = torch.cat((x, newline[:, None, None].expand(*x.shape[:-1], 1)), dim=-1)    # :391
        x = x.flatten(1, 2).transpose(0, 1)                                        # :392
    else:                                                                          # :398-400
        x = x.permute(0, 2, 1, 3, 4).contiguous().flatten(0, 3)
    if "nobase" in merge_type:                                                     # :401-402
        return x
    return torch.cat((base, x), dim=0)                                             # :404


def get_2d_pool(feat: torch.Tensor, mode: str, stride: int = 2, unit: int = 27) -> torch.Tensor:
    """llava_arch.py:171-190 (get_2dPool): [frames, unit*unit, C] -> [frames, h'*w', C]."""
    frames, _, c = feat.shape
    x = feat.view(frames, unit, unit, -1).permute(0, 3, 1, 2).contiguous()
    if mode == "average":
        x = F.avg_pool2d(x, stride)
    elif mode == "max":
        x = F.max_pool2d(x, stride)
    elif mode == "bilinear":
        h, w = x.shape[2:]
        x = F.interpolate(x, size=[math.ceil(h / stride), math.ceil(w / stride)], mode="bilinear")
    else:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")
    return x.permute(0, 2, 3, 1).reshape(frames, -1, c)


def merge_video(feat: torch.Tensor, newline: torch.Tensor, pool_mode: str = "bilinear", newline_position: str = "grid",
                merge_type: str = "spatial_unpad", unit: int = 27) -> torch.Tensor:
    """llava_arch.py:286-290 + 299-349 for one video sample: [frames, T, C] -> [N, C]."""
    x = get_2d_pool(feat, pool_mode, 2, unit)                                      # :286-288 (default stride 2)
    if merge_type == "flat":
        return x.flatten(0, 1)                                                     # :299-300
    frames, p2, c = x.shape
    if newline_position == "grid":                                                 # :222-244 add_token_per_grid
        h = int(math.sqrt(p2))
        y = x.view(frames, 1, h, h, -1).permute(4, 0, 2, 1, 3).contiguous().flatten(1, 2).flatten(2, 3)
        y = torch.cat((y, newline[:, None, None].expand(*y.shape[:-1], 1)), dim=-1)
        return y.flatten(1, 2).transpose(0, 1)
    if newline_position == "frame":                                                # :246-250 add_token_per_frame
        y = x.permute(2, 0, 1).contiguous()
        y = torch.cat((y, newline[:, None, None].expand(*y.shape[:-1], 1)), dim=-1)
        return y.permute(1, 2, 0).contiguous().flatten(0, 1)
    if newline_position == "one_token":                                            # :337-345
        y = x.flatten(0, 1)
        if "unpad" in merge_type:
            y = torch.cat((y, newline[None]), dim=0)
        return y
    if newline_position == "no_token":                                             # :346-347
        return x.flatten(0, 1)
    raise ValueError(f"Unexpected mm_newline_position: {newline_position}")


def prepare_inputs_labels(embed_table: torch.Tensor, image_features: List[torch.Tensor], input_ids: torch.Tensor,
                          attention_mask: Optional[torch.Tensor], labels: Optional[torch.Tensor],
                          max_length: Optional[int] = None, left_pad: bool = False,
                          n_modalities: Optional[int] = None):
    """llava_arch.py:428-555 given the per-image merged features.
    Returns (inputs_embeds [B,max_len,H], labels [B,max_len], attention_mask bool, position_ids)."""
    ids = input_ids.tolist()
    mask = None if attention_mask is None else attention_mask.bool().tolist()
    lay = po.splice_layout(ids, mask, [int(f.shape[0]) for f in image_features], max_length, left_pad, n_modalities)
    B, max_len, H = len(lay["rows"]), lay["max_len"], embed_table.shape[1]
    out = torch.zeros(B, max_len, H, dtype=embed_table.dtype)
    out_labels = torch.full((B, max_len), po.IGNORE_INDEX, dtype=torch.long)
    out_mask = torch.zeros(B, max_len, dtype=torch.bool)
    out_pos = torch.zeros(B, max_len, dtype=torch.long)
    for b, row in enumerate(lay["rows"]):
        pos = 0
        for p, src in enumerate(row):
            if src[0] == "pad":
                continue
            if src[0] == "text":
                out[b, p] = embed_table[ids[src[1]][src[2]]]
                if labels is not None:
                    out_labels[b, p] = labels[src[1], src[2]]
            else:
                out[b, p] = image_features[src[1]][src[2]].to(embed_table.dtype)
            out_mask[b, p] = True
            out_pos[b, p] = pos
            pos += 1
    return out, out_labels, out_mask, out_pos
