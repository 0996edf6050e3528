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

    def submit_flat(self, flat: torch.Tensor) -> None:
        """All-reduce a contiguous 1-D gradient slice IN PLACE, in ``bucket_bytes`` pieces (no flatten, no copy back):
        the encoder keeps all gradient accumulators in one flat buffer, ordered so that what a layer range of the
        backward finishes is one slice.  NCCL averages inside the collective (ReduceOp.AVG); other backends sum and
        ``finish()`` scales."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1 or flat.numel() == 0:
            return
        assert flat.dim() == 1 and flat.is_contiguous()
        per = max(1, self.bucket_bytes // flat.element_size())
        avg_in_op = self.average and dist.get_backend(self.group) == "nccl"
        op = dist.ReduceOp.AVG if avg_in_op else dist.ReduceOp.SUM
        for off in range(0, flat.numel(), per):
            piece = flat[off:off + per]
            work = dist.all_reduce(piece, op=op, group=self.group, async_op=True)
            self._pending.append((work, piece, None if avg_in_op or not self.average else "scale"))

    def finish(self) -> None:
        if not self._pending:
            return
        scale = 1.0 / dist.get_world_size(self.group) if self.average else 1.0
        for work, flat, members in self._pending:
            work.wait()
            if members is None or isinstance(members, str):   # in-place slices (submit_flat)
                if members == "scale" and scale != 1.0:
                    flat.mul_(scale)
                continue
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


# ---------------------------------------------------------------------------------------------------------------
# Fused merge + all-gather over peer memory (radvlm_merge_splice_scatter): no collective call moves the tokens
# ---------------------------------------------------------------------------------------------------------------
class _DevicePtr:
    """Minimal __cuda_array_interface__ wrapper so torch can view memory this library allocated."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerGather:
    """Gathered ``[world, rows, H]`` buffers living in peer-mapped device memory, ``slots`` of them.

    Every rank allocates one region (``radvlm_peer_alloc``), the 64-byte cudaIpc handles are exchanged once through
    ``torch.distributed`` and opened on every rank (``radvlm_peer_open``; one process per GPU on one node, NVLink /
    NVSwitch).  ``prepare_inputs_labels_for_multimodal`` writes the local ``inputs_embeds`` straight into slice
    ``[rank]`` of its own buffer; ``exchange()`` then moves that slice into slice ``[rank]`` of all OTHER ranks'
    buffers on a side stream, in one of two ways:

      * ``mode="ce"`` (default): one ``radvlm_peer_copy`` (``cudaMemcpyAsync``) per peer — DMA engines over NVLink,
        no SM is taken from the next step's persistent GEMM / attention kernels;
      * ``mode="kernel"``: ``radvlm_merge_splice_scatter`` — the gather kernel itself recomputes every row and stores
        it to all peers (fused merge + all-gather; ``max_ctas`` CTAs).

    Two flag exchanges (``radvlm_peer_signal_wait``) bracket the transfer: before it, every rank publishes that it has
    finished READING the previous contents of the slot (consumer release — so a slot is never overwritten under a
    reader, whatever the drift between ranks) and waits for all peers' releases; after it, every rank publishes that
    its rows have landed and waits for all peers'.  After ``wait(slot)`` the whole ``[world, rows, H]`` tensor of
    that slot is valid.

    Invariant the caller keeps: every read of a slot's tensors (``inputs_embeds`` returned while the gather is
    attached, ``gathered(slot)``) is queued on the current stream before the ``slots``-th following call that takes a
    slot.  Under autograd ``inputs_embeds`` is therefore COPIED out of the slot (``copy_out_when_grad``): saved
    activations may live until a backward pass that runs arbitrarily late.

    A peer that does not arrive within ``timeout_s`` (default 600 s, ``RADVLM_B200_PEER_TIMEOUT_S``; 0 = wait for ever)
    does not kill the CUDA context: the wait kernel records the missing rank in a pinned status word and
    ``check()`` (called by ``wait`` / ``drain`` / ``close``) raises ``RuntimeError`` on the host.
    """

    def __init__(self, rows: int, hidden: int, dtype: torch.dtype, device, group=None, slots: int = 2, max_ctas: int = 32,
                 mode: str = "ce", timeout_s: Optional[float] = None, copy_out_when_grad: bool = True):
        import ctypes as C
        import os
        from . import _lib
        self._C, self._lib = C, _lib
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > _lib.MAX_PEERS:
            raise ValueError("PeerGather supports at most %d ranks (one node)" % _lib.MAX_PEERS)
        if mode not in ("ce", "kernel"):
            raise ValueError("PeerGather mode must be 'ce' or 'kernel'")
        self.mode, self.copy_out_when_grad = mode, copy_out_when_grad
        self.timeout_s = float(os.environ.get("RADVLM_B200_PEER_TIMEOUT_S", 600.0)) if timeout_s is None else float(timeout_s)
        self.rows, self.hidden, self.dtype, self.device = int(rows), int(hidden), dtype, torch.device(device)
        self.slots, self.max_ctas = slots, max_ctas
        self.esize = torch.empty(0, dtype=dtype).element_size()
        self.slice_bytes = (self.rows * self.hidden * self.esize + 255) // 256 * 256
        self.slot_bytes = self.world * self.slice_bytes
        self.flag_off = slots * self.slot_bytes
        total = self.flag_off + slots * 256   # per slot: uint64 arrived[8] at +0, uint64 released[8] at +64
        with torch.cuda.device(self.device):
            base = C.c_void_p()
            handle = (C.c_uint8 * 64)()
            _lib.check(self.lib.radvlm_peer_alloc(total, C.byref(base), handle))
            self.base = int(base.value)
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            allh = torch.empty(self.world, 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu().numpy()
            self.peer_base = []
            for r in range(self.world):
                if r == self.rank:
                    self.peer_base.append(self.base)
                    continue
                h = (C.c_uint8 * 64)(*[int(v) for v in allh[r]])
                p = C.c_void_p()
                _lib.check(self.lib.radvlm_peer_open(h, C.byref(p)))
                self.peer_base.append(int(p.value))
            self._bytes = torch.as_tensor(_DevicePtr(self.base, total), device=self.device)
            # per slot: DEVICE arrays of every rank's flag arrays (uint64[world]) for radvlm_peer_signal_wait
            self._arrive_ptrs = [torch.tensor([b + self.flag_off + s * 256 for b in self.peer_base], dtype=torch.int64,
                                              device=self.device) for s in range(slots)]
            self._release_ptrs = [torch.tensor([b + self.flag_off + s * 256 + 64 for b in self.peer_base],
                                               dtype=torch.int64, device=self.device) for s in range(slots)]
            self._status = torch.zeros(1, dtype=torch.int32).pin_memory()   # written by the wait kernel on a timeout
            self.stream = torch.cuda.Stream(self.device)
            self._events = [None] * slots
        self._step = [0] * slots
        self._turn = 0
        dist.barrier(group)

    # -- views
    def gathered(self, slot: int) -> torch.Tensor:
        """[world, rows, H] view of slot ``slot`` in this rank's buffer."""
        off = slot * self.slot_bytes
        flat = self._bytes[off: off + self.slot_bytes].view(self.world, self.slice_bytes)
        return flat[:, : self.rows * self.hidden * self.esize].view(self.dtype).view(self.world, self.rows, self.hidden)

    def local_rows(self, slot: int, n_rows: int) -> torch.Tensor:
        """This rank's own slice (the tensor the merge kernel fills as ``inputs_embeds``)."""
        return self.gathered(slot)[self.rank, :n_rows]

    def peek_slot(self) -> int:
        """The slot the next ``prepare_inputs_labels_for_multimodal`` call will fill."""
        return self._turn % self.slots

    def next_slot(self) -> int:
        s = self._turn % self.slots
        self._turn += 1
        return s

    def remote_dests(self, slot: int):
        """ctypes array of the slice [rank] of slot ``slot`` in every OTHER rank's buffer."""
        C = self._C
        ptrs = [b + slot * self.slot_bytes + self.rank * self.slice_bytes
                for r, b in enumerate(self.peer_base) if r != self.rank]
        return (C.c_void_p * max(len(ptrs), 1))(*ptrs), len(ptrs)

    def _signal_wait(self, ptrs: torch.Tensor, local: int, value: int) -> None:
        self._lib.check(self.lib.radvlm_peer_signal_wait(
            ptrs.data_ptr(), local, self.world, self.rank, value, self.timeout_s, self._status.data_ptr(),
            self.stream.cuda_stream))

    def exchange(self, slot: int, n_rows: int, launch_kernel) -> None:
        """Move this rank's slice of ``slot`` to all peers on the side stream, after everything queued so far on the
        current stream: release barrier, transfer (copy engines, or ``launch_kernel(dests, n_dests, max_ctas,
        cuda_stream)`` = the scatter form of the merge kernel), arrival barrier."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        dests, n = self.remote_dests(slot)
        flags_local = self.base + self.flag_off + slot * 256
        with torch.cuda.stream(self.stream):
            self._step[slot] += 1
            gen = self._step[slot]
            if gen > 1 and n > 0:   # every rank has finished reading generation gen-1 of this slot
                self._signal_wait(self._release_ptrs[slot], flags_local + 64, gen - 1)
            if n > 0:
                if self.mode == "kernel":
                    launch_kernel(dests, n, self.max_ctas, self.stream.cuda_stream)
                else:
                    src = self.base + slot * self.slot_bytes + self.rank * self.slice_bytes
                    nbytes = int(n_rows) * self.hidden * self.esize
                    for d in range(n):
                        self._lib.check(self.lib.radvlm_peer_copy(dests[d], src, nbytes, self.stream.cuda_stream))
            self._signal_wait(self._arrive_ptrs[slot], flags_local, gen)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self._events[slot] = ev

    def scatter(self, slot: int, launch) -> None:   # round-1 name: the kernel form over all rows of the slot
        self.exchange(slot, self.rows, launch)

    def check(self) -> None:
        """Raise if a flag exchange timed out (the status word is only looked at, never waited for)."""
        st = int(self._status[0])
        if st != 0:
            self._status[0] = 0
            raise RuntimeError("radvlm_b200.PeerGather: rank %d did not reach the exchange within %.0f s (rank %d gave up "
                               "waiting; the gathered buffer of that step is incomplete)" % (st - 1, self.timeout_s, self.rank))

    def wait(self, slot: int) -> torch.Tensor:
        """Make the current stream wait until every rank's rows of ``slot`` have landed; returns the gathered view."""
        self.check()
        if self._events[slot] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._events[slot])
        return self.gathered(slot)

    def drain(self) -> None:
        for s in range(self.slots):
            self.wait(s)

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        self.check()
        dist.barrier(self.group)
        for r, b in enumerate(self.peer_base):
            if r != self.rank:
                self.lib.radvlm_peer_close(b)
        dist.barrier(self.group)
        self._bytes = None
        self.lib.radvlm_peer_free(self.base)
