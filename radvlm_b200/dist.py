"""Data-parallel encode: shard images across ranks, all-gather the variable-length visual-token blocks.

SURVEY.md section 8(e): tiles are independent through the tower and projector, the spatial merge needs all tiles
of ONE image, so the unit of sharding is the image.  The reference has no collective on this path (its DP lives
in DeepSpeed); this module is the ``BASELINE.json`` configs[2] plumbing:

  * ``shard_images_lpt``      — greedy longest-processing-time assignment by tile count (images have 2..37 tiles)
  * ``gather_visual_tokens``  — all-gather-v of the merged ``[N_i, H]`` blocks.  Token counts come from the
                                planner on every rank (``image_sizes`` are global), so no size exchange is needed.
  * ``encode_images_sharded`` — preprocess + encode + merge this rank's images, then gather.

Collectives go through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo on CPU for the host-logic tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_images_lpt(tile_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Deterministic greedy LPT: heaviest image first onto the least-loaded rank (ties -> lowest rank).
    Returns, per rank, the global image indices it owns, in ascending order."""
    order = sorted(range(len(tile_counts)), key=lambda i: (-int(tile_counts[i]), i))
    load = [0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += int(tile_counts[i])
    return [sorted(o) for o in owned]


def gather_visual_tokens(local_tokens: torch.Tensor, owned: Sequence[Sequence[int]], token_counts: Sequence[int],
                         group=None) -> List[torch.Tensor]:
    """All-gather-v.  ``local_tokens``: this rank's merged tokens, images concatenated in the order of
    ``owned[rank]``, shape [sum(token_counts[i] for i in owned[rank]), H].  Returns the per-image token blocks in
    GLOBAL image order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per_rank = [sum(int(token_counts[i]) for i in owned[r]) for r in range(world)]
    assert local_tokens.shape[0] == per_rank[rank], (local_tokens.shape, per_rank, rank)
    H = local_tokens.shape[1]
    if world == 1:
        gathered = [local_tokens]
    else:
        pad_to = max(per_rank)
        send = local_tokens
        if send.shape[0] < pad_to:
            send = torch.cat([send, send.new_zeros(pad_to - send.shape[0], H)], dim=0)
        send = send.contiguous()
        buf = send.new_empty(world, pad_to, H)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(buf, send, group=group)
        else:
            parts = [buf[r] for r in range(world)]
            dist.all_gather(parts, send, group=group)
        gathered = [buf[r, : per_rank[r]] for r in range(world)]
    out: List[Optional[torch.Tensor]] = [None] * len(token_counts)
    for r in range(world):
        off = 0
        for i in owned[r]:
            n = int(token_counts[i])
            out[i] = gathered[r][off: off + n]
            off += n
    return out  # type: ignore[return-value]


def encode_images_sharded(images: Sequence, image_sizes: Sequence, tile_counts: Sequence[int],
                          token_counts: Sequence[int], encode_and_merge: Callable[[List[int]], torch.Tensor],
                          group=None) -> List[torch.Tensor]:
    """Shard -> encode this rank's images -> all-gather-v.

    ``encode_and_merge(indices)`` runs the single-GPU path (preprocess, tower, projector, merge) on the images with
    the given global indices and returns their merged tokens concatenated ``[sum N_i, H]``
    (see ``radvlm_b200.mm_arch.merge_images``)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    owned = shard_images_lpt(tile_counts, world)
    local = encode_and_merge(list(owned[rank]))
    return gather_visual_tokens(local, owned, token_counts, group=group)
