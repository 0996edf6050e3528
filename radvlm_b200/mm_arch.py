"""Drop-in replacements for ``LlavaMetaForCausalLM.encode_images`` and
``LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal`` (finetuning/llava/model/llava_arch.py:192-196,
251-555) backed by the sm_100a library.

Same signatures, argument meaning, return tuple and error behaviour as the reference for the RadVLM
configuration (SigLIP tower, ``mlp2x_gelu`` projector, ``mm_patch_merge_type="spatial_unpad"``,
``image_aspect_ratio="anyres_max_9"``); see ``INTEGRATION.md`` for how to attach it to the reference's
``LlavaQwenForCausalLM`` without touching its callers (llava_qwen.py:83,131).

Differences by design: one host sync per batch (the ids are read once to plan the splice) instead of two
per sample; visual tokens are written exactly once, directly into the padded ``[B, max_len, H]`` tensor.
"""
from __future__ import annotations

import ctypes as C
import math
import random
import re
from typing import List, Optional

import numpy as np
import torch

from . import _lib, planner
from .encoder import B200VisionEncoder
from .planner import IGNORE_INDEX, IMAGE_TOKEN_INDEX

_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16}


def _encoder_for(self) -> B200VisionEncoder:
    enc = getattr(self, "_radvlm_b200_encoder", None)
    tower = self.get_vision_tower()
    proj = self.get_model().mm_projector
    if enc is None or enc.tower_module is not tower.vision_tower or enc.projector_module is not proj:
        cfg = getattr(tower, "config", None)
        enc = B200VisionEncoder(
            tower.vision_tower, proj,
            num_heads=getattr(cfg, "num_attention_heads", 16),
            image_size=getattr(cfg, "image_size", 384),
            ln_eps=getattr(cfg, "layer_norm_eps", 1e-6))
        object.__setattr__(self, "_radvlm_b200_encoder", enc)
    return enc


def encode_images(self, images: torch.Tensor) -> torch.Tensor:
    """llava_arch.py:192-196: tower -> mm_projector.  [n,3,S,S] -> [n, 729, hidden_size], dtype = images.dtype.
    With autograd enabled and trainable tower / projector parameters the result carries a grad_fn whose backward
    runs the sm_100a backward kernels (radvlm_projector_backward, radvlm_siglip_tower_backward)."""
    enc = _encoder_for(self)
    if torch.is_grad_enabled() and any(p.requires_grad for p in enc._source_tensors()):
        return enc.encode_images_train(images)
    return enc.encode_images(images)


_POOL_MODES = {"bilinear": _lib.POOL_BILINEAR, "average": _lib.POOL_AVERAGE, "max": _lib.POOL_MAX}


def _video_entry(self, m, frames: int, S: int, merge_type: str):
    """Merge descriptor of one video sample: get_2dPool with the default stride 2 (llava_arch.py:171-190, 286-288),
    then the newline placement of llava_arch.py:310-349 (add_token_per_grid / add_token_per_frame :222-250)."""
    cfg = self.config
    if (getattr(cfg, "add_faster_video", False) and merge_type.startswith("spatial")
            and getattr(cfg, "mm_newline_position", "one_token") == "grid"):
        # The reference cannot run this configuration either: llava_arch.py:317 reads `all_faster_video_features`, which
        # only the commented-out encode_multimodals call (:281) would have defined, so it dies with this NameError
        # after add_token_per_grid.  Every other newline position ignores add_faster_video (:327-349), as does this path.
        raise NameError("name 'all_faster_video_features' is not defined")
    mode = getattr(cfg, "mm_spatial_pool_mode", None)
    if mode not in _POOL_MODES:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")
    side = math.ceil(S / 2) if mode == "bilinear" else S // 2
    m.mode, m.grid_w, m.pool, m.out_h, m.out_w = _lib.MERGE_VIDEO, frames, _POOL_MODES[mode], side, side
    m.crop_r0, m.crop_c0, m.crop_h, m.crop_w = 0, 0, S, S
    per_frame = side * side
    if merge_type == "flat":
        m.reserved, m.n_tokens = _lib.NEWLINE_NONE, frames * per_frame
    elif merge_type.startswith("spatial"):
        position = getattr(cfg, "mm_newline_position", "one_token")
        if position == "grid":
            m.reserved, m.n_tokens = _lib.NEWLINE_GRID, frames * side * (side + 1)
        elif position == "frame":
            m.reserved, m.n_tokens = _lib.NEWLINE_FRAME, frames * (per_frame + 1)
        elif position == "one_token":
            if "unpad" in merge_type:
                m.reserved, m.n_tokens = _lib.NEWLINE_ONE, frames * per_frame + 1
            else:
                m.reserved, m.n_tokens = _lib.NEWLINE_NONE, frames * per_frame
        elif position == "no_token":
            m.reserved, m.n_tokens = _lib.NEWLINE_NONE, frames * per_frame
        else:
            raise ValueError(f"Unexpected mm_newline_position: {position}")
    else:
        raise ValueError("Unexpected mm_patch_merge_type: %s" % merge_type)


def _merge_table(self, tile_counts: List[int], image_sizes, flat_batch: bool, video_idx=()):
    """Per-image merge descriptors + visual-token counts (llava_arch.py:294-413); entries whose index is in
    `video_idx` are video samples (llava_arch.py:269-272, 283-290, 310-349)."""
    cfg = self.config
    tower = self.get_vision_tower()
    S = tower.num_patches_per_side
    T = S * S
    merge_type = getattr(cfg, "mm_patch_merge_type", "flat")
    aspect = getattr(cfg, "image_aspect_ratio", "square")
    table = (_lib.MergeImage * max(len(tile_counts), 1))()
    tokens = []
    base = 0
    for i, tiles in enumerate(tile_counts):
        m = table[i]
        m.tile_base = base
        if i in video_idx and not flat_batch:
            _video_entry(self, m, tiles, S, merge_type)
        elif flat_batch or merge_type == "flat":
            m.mode, m.n_tokens = _lib.MERGE_FLAT, tiles * T
        elif merge_type.startswith("spatial"):
            if tiles > 1:
                max_num_patches = 0
                mt = re.match(r"anyres_max_(\d+)", aspect) if "anyres_max" in aspect else None
                if mt:
                    max_num_patches = int(mt.group(1))
                ts = getattr(tower, "image_size", None)
                if aspect == "anyres" or "anyres_max" in aspect:          # llava_arch.py:362-372
                    if ts is None:
                        raise ValueError("vision_tower_image_size is not found in the vision tower.")
                    pinpoints = cfg.image_grid_pinpoints
                else:                                                     # :373-374  view(2, 2, h, w, -1)
                    ts = ts or 384
                    pinpoints = [[2 * ts, 2 * ts]]
                unpad = "unpad" in merge_type and "maxpool2x2" not in merge_type
                # the pooled branch (:381-392) needs 'unpad', an anyres_max aspect and a matched number
                pool_patches = max_num_patches if (unpad and "anyres_max" in aspect and mt) else 0
                try:
                    if unpad:
                        plan = planner.plan_image(image_sizes[i], pinpoints, ts, S, pool_patches)
                    else:   # no unpad window: only the grid shape depends on the image size
                        plan = planner.plan_image(image_sizes[i], pinpoints, ts, S, 0)
                except Exception as e:
                    if not (aspect == "anyres" or "anyres_max" in aspect):
                        raise
                    # reference: grid-shape failure falls back to a 2x2 grid (llava_arch.py:367-371); the unpad
                    # window is still computed from image_sizes[i] and raises if that is unusable.
                    print("Error: %s" % e)
                    plan = planner.plan_image(image_sizes[i], [[2 * ts, 2 * ts]], ts, S, pool_patches)
                if plan.grid_w * plan.grid_h != tiles - 1:
                    raise RuntimeError("shape '[%d, %d, %d, %d, -1]' is invalid for input of %d tiles (image %d)"
                                       % (plan.grid_h, plan.grid_w, S, S, tiles - 1, i))
                m.mode = _lib.MERGE_ANYRES
                m.grid_w = plan.grid_w
                flags = _lib.ANYRES_NO_BASE if "nobase" in merge_type else 0
                full_h, full_w = S * plan.grid_h, S * plan.grid_w
                if "maxpool2x2" in merge_type:                            # :376-380
                    flags |= _lib.ANYRES_NO_NEWLINE
                    m.crop_r0, m.crop_c0, m.crop_h, m.crop_w = 0, 0, full_h, full_w
                    m.pool, m.out_h, m.out_w = _lib.POOL_MAX, full_h // 2, full_w // 2
                    grid_tokens = m.out_h * m.out_w
                elif unpad:                                               # :381-397
                    m.crop_r0, m.crop_c0, m.crop_h, m.crop_w = plan.crop_r0, plan.crop_c0, plan.crop_h, plan.crop_w
                    m.pool, m.out_h, m.out_w = plan.pool, plan.out_h, plan.out_w
                    grid_tokens = plan.out_h * (plan.out_w + 1)
                else:                                                     # :398-400  plain "spatial": tile-row-major scan
                    flags |= _lib.ANYRES_NO_NEWLINE
                    m.crop_r0, m.crop_c0, m.crop_h, m.crop_w = 0, 0, full_h, full_w
                    m.pool, m.out_h, m.out_w = _lib.POOL_NONE, full_h, full_w
                    grid_tokens = full_h * full_w
                m.reserved = flags
                m.n_tokens = grid_tokens + (0 if flags & _lib.ANYRES_NO_BASE else T)
            else:
                if "unpad" in merge_type:
                    m.mode, m.n_tokens = _lib.MERGE_SINGLE, T + 1
                else:
                    m.mode, m.n_tokens = _lib.MERGE_FLAT, T
        else:
            raise ValueError("Unexpected mm_patch_merge_type: %s" % merge_type)
        tokens.append(int(m.n_tokens))
        base += tiles
    return table, tokens


class _NullCtx:
    """stand-in autograd context for calling a Function's forward directly (inference)"""

    def mark_non_differentiable(self, *a):
        pass


class _MergeSpliceFn(torch.autograd.Function):
    """radvlm_merge_splice under autograd: gradients flow to the visual features, image_newline and embed_tokens
    (llava_arch.py:350-531 are all differentiable gathers / lerps in the reference)."""

    @staticmethod
    def forward(ctx, features, newline, embed, m):
        lib = _lib.load()
        dev = embed.device
        H = embed.shape[1]
        gather = m.get("gather")
        with torch.cuda.device(dev):
            if gather is not None:   # inputs_embeds lives in slice [rank] of a peer-memory gather slot (dist.PeerGather)
                if m["total_rows"] > gather.rows or H != gather.hidden or embed.dtype != gather.dtype:
                    raise ValueError("PeerGather(rows=%d, hidden=%d, %s) cannot hold %d rows of %d x %s" % (
                        gather.rows, gather.hidden, gather.dtype, m["total_rows"], H, embed.dtype))
                m["gather_slot"] = gather.next_slot()
                gather.wait(m["gather_slot"])   # the previous use of this slot has been fully exchanged
                out = gather.local_rows(m["gather_slot"], m["total_rows"]).view(m["B"], m["max_len"], H)
            else:
                out = torch.empty(m["B"], m["max_len"], H, dtype=embed.dtype, device=dev)
            out_labels = torch.empty(m["B"], m["max_len"], dtype=torch.int64, device=dev)
            out_mask = torch.empty(m["B"], m["max_len"], dtype=torch.uint8, device=dev)
            out_pos = torch.empty(m["B"], m["max_len"], dtype=torch.int64, device=dev)
            if m["total_rows"] > 0:
                tables = m["tables"]
                _lib.check(lib.radvlm_merge_splice(
                    features.data_ptr(), newline.data_ptr(), embed.data_ptr(), _DT[embed.dtype], H, m["T"], m["S"],
                    m["ids_dev"].data_ptr(), None if m["labels_dev"] is None else m["labels_dev"].data_ptr(),
                    tables.data_ptr() + m["off_txt"], tables.data_ptr(), m["n_segments"],
                    tables.data_ptr() + m["off_img"], m["n_images"], m["total_rows"],
                    out.data_ptr(), out_labels.data_ptr(), out_mask.data_ptr(), out_pos.data_ptr(), IGNORE_INDEX,
                    torch.cuda.current_stream(dev).cuda_stream))
                if gather is not None:
                    # fused merge + all-gather: the same gather kernel writes every row into slice [rank] of all other
                    # ranks' buffers over NVLink, on the gather's side stream (radvlm_merge_splice_scatter)
                    def launch(dests, n, max_ctas, stream):
                        _lib.check(lib.radvlm_merge_splice_scatter(
                            features.data_ptr(), newline.data_ptr(), embed.data_ptr(), _DT[embed.dtype], H, m["T"], m["S"],
                            m["ids_dev"].data_ptr(), None, tables.data_ptr() + m["off_txt"], tables.data_ptr(),
                            m["n_segments"], tables.data_ptr() + m["off_img"], m["n_images"], m["total_rows"],
                            dests, n, max_ctas, None, None, None, IGNORE_INDEX, stream))
                    for t in (features, newline, embed, tables, m["ids_dev"]):
                        t.record_stream(gather.stream)
                    gather.exchange(m["gather_slot"], m["total_rows"], launch)
                    if getattr(gather, "copy_out_when_grad", False) and not isinstance(ctx, _NullCtx):
                        # under autograd the result may be kept (saved activations) far beyond the slot's reuse
                        out = out.clone()
        ctx.m = m
        # max pooling (video 'max', 'maxpool2x2') routes gradients to the arg-max of every window: keep the features
        ctx.saved_features = features.detach() if m.get("needs_features") and not isinstance(ctx, _NullCtx) else None
        ctx.feat_shape, ctx.feat_dtype = tuple(features.shape), features.dtype
        ctx.newline_dtype, ctx.embed_shape, ctx.embed_dtype = newline.dtype, tuple(embed.shape), embed.dtype
        ctx.mark_non_differentiable(out_labels, out_mask, out_pos)
        return out, out_labels, out_mask, out_pos

    @staticmethod
    def backward(ctx, d_out, *_unused):
        m = ctx.m
        lib = _lib.load()
        dev = d_out.device
        H = ctx.embed_shape[1]
        need_f, need_n, need_e = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_out = d_out.to(ctx.embed_dtype).contiguous()
        with torch.cuda.device(dev):
            n_rows = ctx.feat_shape[0] * ctx.feat_shape[1] if len(ctx.feat_shape) == 3 else ctx.feat_shape[0]
            d_feat = torch.zeros(max(n_rows, 1), H, dtype=torch.float32, device=dev)
            d_newline = torch.zeros(H, dtype=torch.float32, device=dev)
            # zeros, not empty: text tokens cut off by tokenizer_model_max_length are listed in text_src but covered by
            # no segment, so the kernel never writes their rows; the reference gives them zero gradient
            d_text = torch.zeros(max(m["n_text"], 1), H, dtype=ctx.embed_dtype, device=dev) if need_e else None
            if m["total_rows"] > 0:
                tables = m["tables"]
                _lib.check(lib.radvlm_merge_splice_backward(
                    d_out.data_ptr(), _DT[ctx.embed_dtype], H, m["T"], m["S"], tables.data_ptr(), m["n_segments"],
                    tables.data_ptr() + m["off_img"], m["n_images"], m["total_rows"], d_feat.data_ptr(),
                    d_newline.data_ptr(), None if d_text is None else d_text.data_ptr(),
                    ctx.saved_features.data_ptr() if ctx.saved_features is not None else None,
                    torch.cuda.current_stream(dev).cuda_stream))
            g_feat = d_feat.to(ctx.feat_dtype).reshape(ctx.feat_shape) if need_f else None
            g_newline = d_newline.to(ctx.newline_dtype) if need_n else None
            g_embed = None
            if need_e:
                g_embed = torch.zeros(ctx.embed_shape, dtype=ctx.embed_dtype, device=dev)
                if m["n_text"] > 0:
                    src = tables[m["off_txt"]:m["off_txt"] + 4 * m["n_text"]].view(torch.int32).long()
                    tok = m["ids_dev"].reshape(-1)[src]
                    g_embed.index_add_(0, tok, d_text[:m["n_text"]])
        return g_feat, g_newline, g_embed, None


def _to_host_async(t: torch.Tensor, dtype: torch.dtype):
    """Start copying a (small) integer / mask tensor to the host without blocking: -> (host tensor, event or None).
    uint8 means "as a mask": non-zero -> 1.  A synchronous ``.to("cpu")`` issued after the encode kernels would block
    the host until the whole tower has run and leave the GPU idle while the host then plans the splice (measured: host
    and device in lock-step, ~1.3 ms of idle stream per 16-image step)."""
    t = t.detach()
    if dtype == torch.uint8:
        t = t if t.dtype == torch.bool else t.ne(0)
    if t.device.type != "cuda":
        return t.to(dtype).contiguous(), None
    with torch.cuda.device(t.device):
        src = t.to(dtype).contiguous()
        host = torch.empty(src.shape, dtype=dtype, pin_memory=True)
        host.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(t.device))
    return host, ev


def _host_result(fetch) -> np.ndarray:
    host, ev = fetch
    if ev is not None:
        ev.synchronize()
    return host.numpy()


def _concat_tiles(images_list) -> torch.Tensor:
    """``torch.cat(images, 0)`` of llava_arch.py:272 without the copy when the per-image tile blocks already lie back to
    back in one allocation (what ``mm_utils.preprocess_anyres_batch(...)[0].split(splits)`` and the DataLoader collate of
    ``radvlm_b200.data`` hand over): the blocks are then one strided view.  Anything else is concatenated as usual."""
    first = images_list[0]
    if len(images_list) == 1:
        return first
    tile_shape, stride = tuple(first.shape[1:]), first.stride()
    if first.is_contiguous() and first.shape[0] > 0:
        base, off, ok = first.untyped_storage().data_ptr(), first.storage_offset(), True
        per_tile = first[0].numel()
        for im in images_list:
            if (tuple(im.shape[1:]) != tile_shape or im.dtype != first.dtype or im.device != first.device
                    or not im.is_contiguous() or im.untyped_storage().data_ptr() != base or im.storage_offset() != off
                    or im.requires_grad):
                ok = False
                break
            off += im.shape[0] * per_tile
        if ok:
            n = sum(int(im.shape[0]) for im in images_list)
            return first.as_strided((n,) + tile_shape, stride, first.storage_offset())
    return torch.cat([im for im in images_list], dim=0)


def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels,
                                         images, modalities=["image"], image_sizes=None):
    """llava_arch.py:251-555.  Returns (None, position_ids, attention_mask, past_key_values, inputs_embeds, labels)."""
    vision_tower = self.get_vision_tower()
    if vision_tower is None or images is None or input_ids.shape[1] == 1:
        return input_ids, position_ids, attention_mask, past_key_values, None, labels
    if isinstance(modalities, str):
        modalities = [modalities]
    video_idx = {i for i, m in enumerate(modalities) if m == "video"}   # llava_arch.py:269-272
    lib = _lib.load()

    # ---- encode all tiles of all images in one call (llava_arch.py:261-279)
    if type(images) is list or images.ndim == 5:
        if type(images) is list:
            images = [x.unsqueeze(0) if x.ndim == 3 else x for x in images]
        images_list = [im if im.ndim == 4 else im.unsqueeze(0) for im in images]
        concat_images = _concat_tiles(images_list)
        tile_counts = [int(im.shape[0]) for im in images_list]
        flat_batch = False
    else:
        concat_images = images
        tile_counts = [1] * int(images.shape[0])
        flat_batch = True
    # The splice plan needs input_ids / attention_mask on the host.  Their D2H copies are enqueued BEFORE the encode
    # kernels and waited for AFTER (see _to_host_async): the host plans and enqueues the merge kernel while the GPU runs
    # the tower, and the next call's kernels are queued before this call's have finished — the stream never runs dry.
    ids_fetch = _to_host_async(input_ids, torch.int64)
    mask_fetch = None if attention_mask is None else _to_host_async(attention_mask, torch.uint8)
    features = self.encode_images(concat_images)            # [tiles, T, H]  (llava_arch.py:279)
    embed = self.get_model().embed_tokens.weight
    dev = embed.device
    if dev.type != "cuda":
        raise RuntimeError("radvlm_b200: embed_tokens lives on %s; the merge/splice kernel has no CPU fallback" % dev)
    features = features.to(device=dev, dtype=embed.dtype).contiguous()
    if getattr(self.config, "tune_mm_mlp_adapter", False) and getattr(self.config, "mm_use_im_start_end", False):
        raise NotImplementedError
    merge_table, image_tokens = _merge_table(self, tile_counts, image_sizes, flat_batch, video_idx)

    # ---- splice plan on the host: one D2H of the ids instead of 2 syncs per sample (llava_arch.py:428-493)
    _labels, _position_ids, _attention_mask = labels, position_ids, attention_mask
    B, L = input_ids.shape
    B_eff = min(B, len(modalities))  # the reference's zip(new_input_embeds, modalities) truncates (llava_arch.py:499-500)
    ids_host = _host_result(ids_fetch)
    mask_host = None if mask_fetch is None else _host_result(mask_fetch)
    max_length = getattr(self.config, "tokenizer_model_max_length", None)
    left_pad = getattr(self.config, "tokenizer_padding_side", "right") == "left"
    plan = planner.plan_splice(ids_host, mask_host, image_tokens, max_length, left_pad)
    if B_eff < B:
        plan = planner.plan_splice(ids_host[:B_eff], None if mask_host is None else mask_host[:B_eff],
                                   image_tokens, max_length, left_pad)
    max_len = plan.max_len
    newline = getattr(self.get_model(), "image_newline", None)
    if newline is None:
        newline = torch.zeros(embed.shape[1], dtype=embed.dtype, device=dev)
    newline = newline.to(device=dev, dtype=embed.dtype).contiguous()
    H = embed.shape[1]
    total_rows = B_eff * max_len

    # ---- tables -> device (one pinned staging buffer)
    seg_bytes = plan.n_segments * C.sizeof(_lib.SpliceSegment)
    img_bytes = len(tile_counts) * C.sizeof(_lib.MergeImage)
    txt_bytes = plan.n_text * 4
    off_img = (seg_bytes + 15) // 16 * 16
    off_txt = off_img + (img_bytes + 15) // 16 * 16
    tot = off_txt + (txt_bytes + 15) // 16 * 16 + 16
    host = torch.zeros(tot, dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    hv[:seg_bytes] = np.frombuffer(bytes(plan.segments)[:seg_bytes], dtype=np.uint8)
    hv[off_img:off_img + img_bytes] = np.frombuffer(bytes(merge_table)[:img_bytes], dtype=np.uint8)
    if txt_bytes:
        hv[off_txt:off_txt + txt_bytes] = plan.text_src.view(np.uint8)
    with torch.cuda.device(dev):
        meta = dict(
            tables=host.to(dev, non_blocking=True), off_img=off_img, off_txt=off_txt, n_segments=plan.n_segments,
            n_images=len(tile_counts), n_text=plan.n_text, total_rows=total_rows, B=B_eff, max_len=max_len,
            T=vision_tower.num_patches_per_side ** 2, S=vision_tower.num_patches_per_side,
            ids_dev=input_ids.detach().to(dev, torch.int64).contiguous(),
            labels_dev=None if labels is None else labels.detach().to(dev, torch.int64).contiguous(),
            gather=getattr(self, "radvlm_b200_gather", None),
            needs_features=any(int(e.pool) == _lib.POOL_MAX for e in merge_table[:len(tile_counts)]))
    if torch.is_grad_enabled() and (features.requires_grad or newline.requires_grad or embed.requires_grad):
        out, out_labels, out_mask, out_pos = _MergeSpliceFn.apply(features, newline, embed, meta)
    else:
        with torch.no_grad():
            out, out_labels, out_mask, out_pos = _MergeSpliceFn.forward(_NullCtx(), features.detach(), newline.detach(),
                                                                        embed.detach(), meta)

    # ---- return contract (llava_arch.py:533-555)
    new_labels = None if _labels is None else out_labels.to(_labels.dtype)
    if _attention_mask is None:
        new_mask = None
    else:
        new_mask = out_mask.bool().to(dtype=_attention_mask.dtype)
    new_pos = None if _position_ids is None else out_pos.to(_position_ids.dtype)
    if getattr(self.config, "use_pos_skipping", False) and self.training:
        new_pos = torch.arange(out.size(1), device=out.device).unsqueeze(0).to(out.device)
        split_position = random.randint(0, out.size(1))
        left_add = random.randint(0, self.config.pos_skipping_range)
        right_add = random.randint(left_add, self.config.pos_skipping_range)
        new_pos[:, :split_position] += left_add
        new_pos[:, split_position:] += right_add
    return None, new_pos, new_mask, past_key_values, out, new_labels


def merge_images(self, features: torch.Tensor, tile_counts: List[int], image_sizes) -> torch.Tensor:
    """Per-image spatial merge only (llava_arch.py:350-412): features [tiles, T, H] -> merged visual tokens
    of all images concatenated, [sum N_i, H].  Used by the data-parallel encode (radvlm_b200.dist), where the
    token blocks are all-gathered before the splice."""
    lib = _lib.load()
    tower = self.get_vision_tower()
    table, tokens = _merge_table(self, tile_counts, image_sizes, False)
    total = int(sum(tokens))
    dev = features.device
    features = features.contiguous()
    H = features.shape[-1]
    newline = getattr(self.get_model(), "image_newline", None)
    if newline is None:
        newline = torch.zeros(H, dtype=features.dtype, device=dev)
    newline = newline.detach().to(device=dev, dtype=features.dtype).contiguous()
    segs = (_lib.SpliceSegment * max(len(tokens), 1))()
    row = 0
    for i, n in enumerate(tokens):
        segs[i].dst_row, segs[i].length, segs[i].kind, segs[i].image = row, n, _lib.SEG_IMAGE, i
        row += n
    seg_bytes = len(tokens) * C.sizeof(_lib.SpliceSegment)
    img_bytes = len(tokens) * C.sizeof(_lib.MergeImage)
    off_img = (seg_bytes + 15) // 16 * 16
    host = torch.zeros(off_img + img_bytes + 16, dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    hv[:seg_bytes] = np.frombuffer(bytes(segs)[:seg_bytes], dtype=np.uint8)
    hv[off_img:off_img + img_bytes] = np.frombuffer(bytes(table)[:img_bytes], dtype=np.uint8)
    out = torch.empty(total, H, dtype=features.dtype, device=dev)
    if total == 0:
        return out
    with torch.cuda.device(dev):
        tables = host.to(dev, non_blocking=True)
        _lib.check(lib.radvlm_merge_splice(
            features.data_ptr(), newline.data_ptr(), None, _DT[features.dtype], H,
            tower.num_patches_per_side ** 2, tower.num_patches_per_side, None, None, None,
            tables.data_ptr(), len(tokens), tables.data_ptr() + off_img, len(tokens), total,
            out.data_ptr(), None, None, None, IGNORE_INDEX, torch.cuda.current_stream(dev).cuda_stream))
    return out


class B200LlavaMetaForCausalLM:
    """Mixin with the reference's method names; put it BEFORE ``LlavaMetaForCausalLM`` in the MRO."""

    encode_images = encode_images
    prepare_inputs_labels_for_multimodal = prepare_inputs_labels_for_multimodal


def attach(model):
    """Monkey-patch one model instance (any class built on LlavaMetaForCausalLM) to use the B200 path."""
    import types

    model.encode_images = types.MethodType(encode_images, model)
    model.prepare_inputs_labels_for_multimodal = types.MethodType(prepare_inputs_labels_for_multimodal, model)
    return model
