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


# ---------------------------------------------------------------------------------------------------------------
# Training mode (BASELINE.json configs[4]): replicas of the tower + projector, gradients all-reduced in buckets
# ---------------------------------------------------------------------------------------------------------------
def plan_gradient_buckets(numels: Sequence[int], bucket_elems: int) -> List[List[int]]:
    """Greedy buckets in the given order (the caller passes tensors in the order the backward finishes them:
    top layers first), each at most ``bucket_elems`` elements unless a single tensor is larger."""
    buckets: List[List[int]] = []
    cur, cur_n = [], 0
    for i, n in enumerate(numels):
        if cur and cur_n + int(n) > bucket_elems:
            buckets.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += int(n)
    if cur:
        buckets.append(cur)
    return buckets


class GradientAllReducer:
    """Bucketed, asynchronous sum (then scale) of gradient tensors over the data-parallel group.

    ``submit(tensors)`` may be called several times while the backward is still running (one call per finished
    layer range); every bucket is flattened, all-reduced with ``async_op=True`` (NCCL runs it on its own stream, over
    NVLink / NVSwitch) and written back in ``finish()``, which also applies the 1 / world_size average."""

    def __init__(self, group=None, bucket_bytes: int = 64 << 20, average: bool = True):
        self.group, self.bucket_bytes, self.average = group, bucket_bytes, average
        self._pending = []

    def submit(self, tensors: Sequence[torch.Tensor]) -> None:
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1 or not tensors:
            return
        tensors = list(tensors)
        per = max(1, self.bucket_bytes // max(tensors[0].element_size(), 1))
        for bucket in plan_gradient_buckets([t.numel() for t in tensors], per):
            members = [tensors[i] for i in bucket]
            flat = torch.cat([t.reshape(-1) for t in members]) if len(members) > 1 else members[0].reshape(-1)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, flat, members))

    def finish(self) -> None:
        if not self._pending:
            return
        scale = 1.0 / dist.get_world_size(self.group) if self.average else 1.0
        for work, flat, members in self._pending:
            work.wait()
            off = 0
            for t in members:
                n = t.numel()
                src = flat[off:off + n].view_as(t)
                if src.data_ptr() != t.data_ptr():
                    t.copy_(src)
                if scale != 1.0:
                    t.mul_(scale)
                off += n
        self._pending = []


def allreduce_gradients(tensors: Sequence[torch.Tensor], group=None, bucket_bytes: int = 64 << 20,
                        average: bool = True) -> None:
    """One-shot form of :class:`GradientAllReducer` (blocking)."""
    r = GradientAllReducer(group, bucket_bytes, average)
    r.submit(tensors)
    r.finish()
