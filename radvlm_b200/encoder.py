"""Host side of the B200 encode path: weight packing and the ``encode_images`` runner.

Mirrors ``SigLipVisionTower.forward`` (siglip_encoder.py:576-589) followed by ``mm_projector``
(llava_arch.py:192-196).  PyTorch is used only for device memory, streams and parameter storage; all
compute is done by the sm_100a kernels behind the C ABI (``libradvlm_b200.so``).

Weights are consumed from the reference's own Parameters (state-dict names).  The
bf16 matrices the tensor-core kernels read are packed *copies* (or aliases of bf16 Parameters) that are refreshed
in place: see ``B200VisionEncoder.packed`` for when (storage change -> rebuild; version bump -> refresh; any
trainable Parameter -> refresh on every call, because DeepSpeed-style ``p.data`` updates bypass ``_version``).
"""
from __future__ import annotations

import ctypes as C
import os
import math
import time
import warnings
from collections import OrderedDict
from typing import Dict, Mapping, Optional

import torch

from . import _lib

_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class PackedWeights:
    """bf16 / fp32 device copies of tower + projector weights and the C structs pointing at them.

    Every destination buffer is allocated once; ``refresh()`` re-copies the CURRENT contents of the source
    Parameters into the same buffers (same device pointers: the C structs, cached TMA descriptors and captured CUDA
    graphs stay valid).  A source that already is a contiguous tensor of the wanted dtype on the device is not copied
    at all: the struct points at the Parameter's own storage (``alias``), so optimizer updates are seen for free."""

    def __init__(self, tower_sd: Mapping[str, torch.Tensor], proj_sd: Mapping[str, torch.Tensor], device,
                 num_heads: int = 16, image_size: int = 384, ln_eps: float = 1e-6,
                 num_layers: Optional[int] = None):
        dev = torch.device(device)
        keep = []       # every device tensor referenced by the structs
        copies = []     # (destination view, source tensor): what refresh() re-copies
        folds = []      # LayerNorm folded into the QKV / fc1 GEMMs: (W sources, bias sources, gamma, beta, wf, sf, bf)
        self.n_alias = 0
        # RADVLM_B200_LN=kernel keeps the stand-alone LayerNorm kernels (A/B switch; also the safer choice for a
        # residual stream whose per-token mean dwarfs its spread: the fold multiplies bf16(x), not bf16(LN(x)))
        self.ln_fold = os.environ.get("RADVLM_B200_LN", "fold") != "kernel"

        def bind(srcs, dtype, pad_cols: int = 0):
            """rows of `srcs` stacked -> one contiguous [sum rows, cols (+ zero padding)] device tensor of `dtype`"""
            srcs = [t.detach() for t in srcs]
            if len(srcs) == 1 and pad_cols == 0 and srcs[0].device == dev and srcs[0].dtype == dtype \
                    and srcs[0].is_contiguous():
                keep.append(srcs[0])     # alias: the Parameter's own storage
                self.n_alias += 1
                return srcs[0]
            flat = [t.reshape(t.shape[0], -1) if t.dim() > 1 else t for t in srcs]
            rows = sum(t.shape[0] for t in flat)
            if flat[0].dim() == 1:
                dst = torch.empty(rows, dtype=dtype, device=dev)
            else:
                cols = flat[0].shape[1]
                dst = torch.zeros(rows, cols + pad_cols, dtype=dtype, device=dev)
            off = 0
            for t in flat:
                view = dst[off:off + t.shape[0]] if t.dim() == 1 else dst[off:off + t.shape[0], :t.shape[1]]
                copies.append((view, t))
                off += t.shape[0]
            keep.append(dst)
            return dst

        w = lambda *ts, pad=0: bind(ts, torch.bfloat16, pad)   # matrices -> bf16 [out, in]
        v = lambda *ts: bind(ts, torch.float32)                # vectors / tables -> fp32

        pre = "vision_model."
        pw = tower_sd[pre + "embeddings.patch_embedding.weight"]  # [hidden, C, ps, ps]
        hidden, channels, ps, _ = pw.shape
        k = channels * ps * ps
        k_pad = (k + 63) // 64 * 64
        patch_w = w(pw, pad=k_pad - k)
        pos = v(tower_sd[pre + "embeddings.position_embedding.weight"])
        if num_layers is None:
            num_layers = 0
            while (pre + "encoder.layers.%d.layer_norm1.weight" % num_layers) in tower_sd:
                num_layers += 1
        layers = (_lib.VitLayerWeights * num_layers)()
        inter = None
        for i in range(num_layers):
            lp = pre + "encoder.layers.%d." % i
            g = lambda n: tower_sd[lp + n]
            L = layers[i]
            L.ln1_gamma = v(g("layer_norm1.weight")).data_ptr()
            L.ln1_beta = v(g("layer_norm1.bias")).data_ptr()
            L.qkv_w = w(g("self_attn.q_proj.weight"), g("self_attn.k_proj.weight"), g("self_attn.v_proj.weight")).data_ptr()
            L.qkv_b = v(g("self_attn.q_proj.bias"), g("self_attn.k_proj.bias"), g("self_attn.v_proj.bias")).data_ptr()
            L.out_w = w(g("self_attn.out_proj.weight")).data_ptr()
            L.out_b = v(g("self_attn.out_proj.bias")).data_ptr()
            L.ln2_gamma = v(g("layer_norm2.weight")).data_ptr()
            L.ln2_beta = v(g("layer_norm2.bias")).data_ptr()
            L.fc1_w = w(g("mlp.fc1.weight")).data_ptr()
            L.fc1_b = v(g("mlp.fc1.bias")).data_ptr()
            L.fc2_w = w(g("mlp.fc2.weight")).data_ptr()
            L.fc2_b = v(g("mlp.fc2.bias")).data_ptr()
            inter = g("mlp.fc1.weight").shape[0]
            if self.ln_fold:
                for tag, ws_, bs_, ln in (
                        ("qkv", [g("self_attn.%s_proj.weight" % n) for n in "qkv"],
                         [g("self_attn.%s_proj.bias" % n) for n in "qkv"], "layer_norm1"),
                        ("fc1", [g("mlp.fc1.weight")], [g("mlp.fc1.bias")], "layer_norm2")):
                    rows = sum(t.shape[0] for t in ws_)
                    wf = torch.empty(rows, hidden, dtype=torch.bfloat16, device=dev)
                    sf = torch.empty(rows, dtype=torch.float32, device=dev)
                    bf = torch.empty(rows, dtype=torch.float32, device=dev)
                    keep.extend((wf, sf, bf))
                    folds.append(([t.detach() for t in ws_], [t.detach() for t in bs_], g(ln + ".weight").detach(),
                                  g(ln + ".bias").detach(), wf, sf, bf))
                    setattr(L, tag + "_wf", wf.data_ptr())
                    setattr(L, tag + "_sf", sf.data_ptr())
                    setattr(L, tag + "_bf", bf.data_ptr())
        tw = _lib.SiglipWeights()
        tw.hidden, tw.intermediate, tw.heads, tw.num_layers = hidden, int(inter or 0), num_heads, num_layers
        tw.image_size, tw.patch_size, tw.channels, tw.patch_k_pad = image_size, ps, channels, k_pad
        tw.ln_eps = ln_eps
        tw.patch_w = patch_w.data_ptr()
        tw.patch_b = v(tower_sd[pre + "embeddings.patch_embedding.bias"]).data_ptr()
        tw.pos_embed = pos.data_ptr()
        tw.layers = C.cast(layers, C.POINTER(_lib.VitLayerWeights))
        pj = _lib.ProjectorWeights()
        w1, w2 = proj_sd["0.weight"], proj_sd["2.weight"]
        pj.in_dim, pj.hidden = w1.shape[1], w1.shape[0]
        assert tuple(w2.shape) == (pj.hidden, pj.hidden), "mlp2x_gelu projector expected (builder.py:41-48)"
        pj.w1 = w(w1).data_ptr()
        pj.b1 = v(proj_sd["0.bias"]).data_ptr()
        pj.w2 = w(w2).data_ptr()
        pj.b2 = v(proj_sd["2.bias"]).data_ptr()
        self.tower, self.projector = tw, pj
        self._layers, self._keep, self._copies = layers, keep, copies
        self._folds, self._folds_stale = folds, True
        self.device = dev
        self.tokens_per_tile = (image_size // ps) ** 2
        self.patches_per_side = image_size // ps
        self.hidden, self.proj_hidden = hidden, pj.hidden
        self.refresh()

    @torch.no_grad()
    def refresh(self) -> None:
        """Re-copy the current values of every non-aliased source into its packed buffer (dtype conversion included)."""
        self._folds_stale = True
        if not self._copies:
            return
        same = [(d, s) for d, s in self._copies if s.device == d.device]
        if same:
            torch._foreach_copy_([d for d, _ in same], [s.reshape(d.shape) for d, s in same])
        for d, s in self._copies:
            if s.device != d.device:
                d.copy_(s.reshape(d.shape))


    @torch.no_grad()
    def ensure_folds(self) -> None:
        """(gamma o W, its row sums, b + W beta) of every folded LayerNorm, recomputed from the CURRENT source values when
        they may have changed (inference calls only: the training forward keeps the stand-alone LayerNorm kernels).
        LN(x) W^T + b = rstd (x (gamma o W)^T - mean s) + (b + W beta); siglip_encoder.py:264,266,287,296."""
        if not self._folds_stale:
            return
        dev = self.device
        for ws_, bs_, gamma, beta, wf, sf, bf in self._folds:
            W = torch.cat([t.to(dev, torch.float32) for t in ws_], 0)
            wf.copy_(W * gamma.to(dev, torch.float32)[None, :])
            sf.copy_(wf.float().sum(1))          # of the bf16 values the tensor core multiplies
            bf.copy_(torch.cat([t.to(dev, torch.float32) for t in bs_], 0) + W @ beta.to(dev, torch.float32))
        self._folds_stale = False


def _layout_key(tensors) -> tuple:
    """What the packed buffers / aliases depend on structurally: storage address, shape, dtype, device."""
    return tuple((t.data_ptr(), tuple(t.shape), t.dtype, t.device) for t in tensors)


def _version_key(tensors) -> tuple:
    return tuple(t._version for t in tensors)


def _check_materialised(named) -> None:
    """DeepSpeed ZeRO-3 replaces the storage of partitioned Parameters by an empty placeholder outside of its own
    module hooks (which this path does not go through: it reads the Parameters directly)."""
    for name, p in named:
        if p.numel() == 0 or hasattr(p, "ds_tensor") and getattr(p, "ds_status", None) is not None \
                and "NOT_AVAILABLE" in str(p.ds_status):
            raise RuntimeError(
                "radvlm_b200: parameter %r is partitioned (DeepSpeed ZeRO-3) and not gathered.  Exclude the vision tower "
                "and mm_projector from partitioning, or call encode_images / prepare_inputs_labels_for_multimodal inside "
                "deepspeed.zero.GatheredParameters(list(tower.parameters()) + list(projector.parameters())) "
                "(INTEGRATION.md, 'DeepSpeed')." % name)


class B200VisionEncoder:
    """``encode_images`` for a (vision tower, projector) pair, executed by the sm_100a library."""

    def __init__(self, tower_module: torch.nn.Module, projector_module: torch.nn.Module, num_heads: int = 16,
                 image_size: int = 384, ln_eps: float = 1e-6, max_tiles_per_call: int = 80):
        self.tower_module = tower_module          # SigLipVisionModel (state-dict prefix "vision_model.")
        self.projector_module = projector_module  # nn.Sequential(Linear, GELU, Linear)
        self.num_heads, self.image_size, self.ln_eps = num_heads, image_size, ln_eps
        self.max_tiles_per_call = int(os.environ.get("RADVLM_B200_MAX_TILES", max_tiles_per_call))   # tuning override
        # training mode, data parallel: gradients are all-reduced (averaged) inside the backward, overlapped with it.
        # False (default) leaves them local (e.g. when DistributedDataParallel / DeepSpeed owns the reduction);
        # None = the default process group; or pass a process group.
        self.grad_allreduce_group = False
        self.grad_bucket_bytes = 64 << 20
        self.backward_layers_per_range = 4
        self._packed: Optional[PackedWeights] = None
        self._layout = None
        self._versions = None
        self._stale = False
        self._frozen = False
        self.n_repacks = 0     # diagnostics (tests): full rebuilds / in-place refreshes of the packed weights
        self.n_refreshes = 0
        self._ws: Dict[torch.device, torch.Tensor] = {}
        # Opt-in (RADVLM_B200_GRAPH=1 or `enc.graph_mode = True`): inference calls are replayed from CUDA graphs keyed on
        # everything the C call depends on (see _launch_encode).  Off by default: a replay saves ~0.3 ms per call
        # (0.4 % of a 16-image step, 3 % of a one-image call) but a capture blocks the host for 15-100 ms
        # (profiles/r02s_*, r02t_*), which a latency-bound caller with drifting buffer addresses would notice.
        # `capture()` is the explicit form for fixed-shape serving loops.
        self.graph_mode = os.environ.get("RADVLM_B200_GRAPH", "0") == "1"
        self.graph_capacity = 16
        self.graph_min_sightings = 2          # eager launches of a key before it is captured
        self.graph_free_captures = 2
        self.graph_replays_per_capture = 48
        self.graph_capture_seconds = 0.0      # host time spent capturing (diagnostics)
        self._graphs: "OrderedDict[tuple, torch.cuda.CUDAGraph]" = OrderedDict()
        self._graph_seen: "OrderedDict[tuple, int]" = OrderedDict()
        self._graph_streams: Dict[torch.device, torch.cuda.Stream] = {}
        self._pk_serial = 0
        self.n_graph_replays = self.n_graph_captures = self.n_eager_launches = 0
        _lib.load()  # fail loudly at construction time if the extension is missing

    # ---- weights -------------------------------------------------------------------------------
    def _source_named(self):
        return list(self.tower_module.named_parameters()) + list(self.projector_module.named_parameters())

    def _source_tensors(self):
        return [p for _, p in self._source_named()]

    def freeze(self, flag: bool = True):
        """Skip every per-call check (pure inference: the weights never change)."""
        self._frozen = flag
        return self

    def invalidate(self):
        """Force the packed copies to be refreshed on the next call.  Needed only when weights are changed through
        ``p.data`` (no ``_version`` bump) while no Parameter requires grad, e.g. a manual ``p.data.copy_`` at inference."""
        self._stale = True
        return self

    def packed(self, device, training: bool = False) -> PackedWeights:
        """The packed weights, up to date with the source Parameters.

        * a different storage / shape / dtype of any source (checkpoint load, ``p.data = ...``, ``.to()``) rebuilds;
        * an in-place update seen by autograd's version counter (torch.optim) refreshes the copies in place;
        * updates through ``p.data`` do NOT bump ``_version`` (DeepSpeed's optimizers, ZeRO gathers): while any source
          Parameter requires grad (the model is being trained, also during its eval passes) the copies are therefore
          refreshed on EVERY call (~0.4 ms for the full tower: one multi-tensor copy into the existing buffers)."""
        dev = torch.device(device)
        if self._packed is not None and self._frozen and self._packed.device == dev:
            return self._packed
        named = self._source_named()
        srcs = [p for _, p in named]
        _check_materialised(named)
        layout = (dev, _layout_key(srcs))
        if self._packed is None or layout != self._layout:
            n_layers = len(self.tower_module.vision_model.encoder.layers)
            self._packed = PackedWeights(self.tower_module.state_dict(), self.projector_module.state_dict(),
                                         device, num_heads=self.num_heads, image_size=self.image_size,
                                         ln_eps=self.ln_eps, num_layers=n_layers)
            self._layout, self._versions, self._stale = layout, _version_key(srcs), False
            self.n_repacks += 1
            self._pk_serial += 1       # captured graphs hold the old buffers' addresses: their keys no longer match
            self._graphs.clear()
            self._graph_seen.clear()
            return self._packed
        versions = _version_key(srcs)
        if training or self._stale or versions != self._versions or any(p.requires_grad for p in srcs):
            self._packed.refresh()
            self._versions, self._stale = versions, False
            self.n_refreshes += 1
        return self._packed

    def _workspace(self, device, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(device)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws[device] = ws
        return ws

    # ---- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_images(self, images: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """images [n, 3, S, S] -> features [n, T, projector_hidden] (dtype = images.dtype, siglip_encoder.py:586)."""
        if images.dim() != 4:
            raise ValueError("encode_images expects [n, C, S, S], got %s" % (tuple(images.shape),))
        if not torch.cuda.is_available():
            raise RuntimeError("radvlm_b200: no CUDA device; the encode path has no CPU fallback")
        dev = next(self.projector_module.parameters()).device
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        out_dtype = out_dtype or images.dtype
        if images.dtype not in _DT:
            images = images.float()
        images = images.to(dev).contiguous()
        pk = self.packed(dev)
        pk.ensure_folds()
        lib = _lib.load()
        n = images.shape[0]
        T, Hp = pk.tokens_per_tile, pk.proj_hidden
        kernel_out = out_dtype if out_dtype in _DT else torch.float32
        out = torch.empty(n, T, Hp, dtype=kernel_out, device=dev)
        with torch.cuda.device(dev):
            stream = _stream_ptr(dev)
            step = max(1, self.max_tiles_per_call)
            for s in range(0, n, step):
                m = min(step, n - s)
                need = lib.radvlm_encode_workspace_bytes(C.byref(pk.tower), C.byref(pk.projector), m)
                if need == 0:
                    raise _lib.RadvlmError(_lib.ERR_BAD_ARGUMENT, _lib.last_error())
                ws = self._workspace(dev, need)
                self._launch_encode(lib, pk, dev, images[s:s + m].data_ptr(), _DT[images.dtype], m,
                                    out[s:s + m].data_ptr(), _DT[kernel_out], ws)
        return out if out.dtype == out_dtype else out.to(out_dtype)

    def _launch_encode(self, lib, pk: PackedWeights, dev, in_ptr: int, in_dt: int, m: int, out_ptr: int, out_dt: int,
                       ws: torch.Tensor) -> None:
        """One ``radvlm_encode_images`` call (~5 launches per layer) on the current stream — replayed from a CUDA graph when
        this exact call has been seen before.

        The C entry point allocates nothing and never synchronises, and everything it does is a function of (packed
        weights, input / output / workspace addresses, tile count, dtypes): that tuple is the graph key.  In a steady
        loop the caching allocator hands the same addresses back every step, so once a key has come back a few times the
        call costs ONE graph launch: the kernels then follow each other over graph edges instead of stream order
        (measured on B200: 8.07 -> 7.55 ms for a 10-tile image, profiles/r02q_bench_final.json).  A key seen for the
        first time is launched eagerly; weight refreshes keep the addresses (replays read the new values), a rebuild of
        the packed weights clears the cache.  Never used while the per-launch profiler is on or while the caller itself
        is capturing."""
        args = (C.byref(pk.tower), C.byref(pk.projector), in_ptr, in_dt, m, out_ptr, out_dt, ws.data_ptr(), ws.numel())
        cur = torch.cuda.current_stream(dev)
        if not self.graph_mode or _lib.PROFILING or torch.cuda.is_current_stream_capturing():
            self.n_eager_launches += 1
            _lib.check(lib.radvlm_encode_images(*args, cur.cuda_stream))
            return
        key = (self._pk_serial, dev.index, in_ptr, in_dt, m, out_ptr, out_dt, ws.data_ptr(), ws.numel())
        g = self._graphs.get(key)
        if g is None:
            # A capture (record + instantiate + first upload) costs the host about as much as 30-40 replays save, so it has
            # to be earned: the key must have come back `graph_min_sightings` times (a steady loop, not a one-off set of
            # addresses), and beyond the first `graph_free_captures` every capture must be covered by
            # `graph_replays_per_capture` replays already served — a caller whose addresses never repeat pays for two
            # captures at most and is then launched eagerly for good.
            seen = self._graph_seen.get(key, 0) + 1
            self._graph_seen[key] = seen
            self._graph_seen.move_to_end(key)
            if len(self._graph_seen) > 8 * self.graph_capacity:
                self._graph_seen.popitem(last=False)
            allowed = self.graph_free_captures + self.n_graph_replays // self.graph_replays_per_capture
            if seen <= self.graph_min_sightings or self.n_graph_captures >= allowed:
                self.n_eager_launches += 1
                _lib.check(lib.radvlm_encode_images(*args, cur.cuda_stream))
                return
            t_cap = time.perf_counter()
            side = self._graph_streams.get(dev)
            if side is None:
                side = self._graph_streams[dev] = torch.cuda.Stream(dev)
            g = torch.cuda.CUDAGraph()
            status, began = _lib.OK, False
            try:
                with torch.cuda.stream(side):
                    # thread_local: only this thread's calls are checked (an NCCL watchdog or a pin-memory thread may
                    # touch the CUDA API meanwhile); nothing executes during capture, so `side` needs no ordering
                    g.capture_begin(capture_error_mode="thread_local")
                    began = True
                    try:
                        status = lib.radvlm_encode_images(*args, side.cuda_stream)
                    finally:
                        g.capture_end()
                _lib.check(status)
            except Exception as e:   # keep working without graphs rather than fail the call
                self.graph_mode = False
                self._graphs.clear()
                warnings.warn("radvlm_b200: CUDA-graph capture of encode_images failed (%s: %s); launching eagerly from "
                              "now on" % (type(e).__name__, e))
                if began and status != _lib.OK:
                    raise
                self.n_eager_launches += 1
                _lib.check(lib.radvlm_encode_images(*args, cur.cuda_stream))
                return
            self.n_graph_captures += 1
            self.graph_capture_seconds += time.perf_counter() - t_cap
            self._graphs[key] = g
            if len(self._graphs) > self.graph_capacity:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        self.n_graph_replays += 1
        g.replay()

    @torch.no_grad()
    def tower_forward(self, images: torch.Tensor) -> torch.Tensor:
        """SigLipVisionTower.forward only: [n,3,S,S] -> fp32 hidden_states[-1] [n, T, hidden]."""
        dev = next(self.projector_module.parameters()).device
        if images.dtype not in _DT:
            images = images.float()
        images = images.to(dev).contiguous()
        pk = self.packed(dev)
        pk.ensure_folds()
        lib = _lib.load()
        n = images.shape[0]
        out = torch.empty(n, pk.tokens_per_tile, pk.hidden, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            need = lib.radvlm_encode_workspace_bytes(C.byref(pk.tower), C.byref(pk.projector), n)
            ws = self._workspace(dev, need)
            _lib.check(lib.radvlm_siglip_tower_forward(C.byref(pk.tower), images.data_ptr(), _DT[images.dtype], n,
                                                       out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)))
        return out


    def capture(self, n_tiles: int, in_dtype: torch.dtype = torch.bfloat16, out_dtype: Optional[torch.dtype] = None,
                channels: int = 3) -> "GraphedEncode":
        """``encode_images`` for a fixed tile count as ONE CUDA graph (see ``GraphedEncode``)."""
        return GraphedEncode(self, n_tiles, in_dtype, out_dtype, channels)


class GraphedEncode:
    """One ``encode_images`` call of ``n_tiles`` tiles captured in a CUDA graph and replayed.

    The C entry point allocates nothing, never synchronises and takes its stream as an argument, so the whole call
    (im2col, ~135 GEMM / attention launches, the projector) records into a graph as is; the kernel parameters — TMA
    descriptors and tile schedules included — are frozen in the graph nodes, a replay costs one launch on the host.
    Serving-side use (model_worker.py:124-127: the same tile count per request shape): small calls, where the host-side
    launch path (~5 us per kernel) is comparable with the kernels themselves.

    * ``images`` (static input) / ``features`` (static output, overwritten by the next replay) belong to the object;
      ``__call__(x)`` copies ``x`` in, replays and returns ``features``.
    * The graph reads the packed weights at their addresses: in-place refreshes (``PackedWeights.refresh``, aliases of
      bf16 Parameters) are seen by later replays; a REBUILD of the packed weights (new storage: checkpoint load,
      ``.to()``) invalidates the graph and ``__call__`` raises.  Refreshes are not part of the graph: call
      ``encoder.packed(device)`` before a replay if weights may have changed through ``p.data``.
    * Inference only (no autograd), per-launch profiling must be off while capturing."""

    def __init__(self, enc: B200VisionEncoder, n_tiles: int, in_dtype: torch.dtype = torch.bfloat16,
                 out_dtype: Optional[torch.dtype] = None, channels: int = 3):
        if n_tiles < 1 or n_tiles > enc.max_tiles_per_call:
            raise ValueError("capture: 1 <= n_tiles <= max_tiles_per_call (%d), got %d" % (enc.max_tiles_per_call, n_tiles))
        dev = next(enc.projector_module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("radvlm_b200: CUDA-graph capture needs the modules on a CUDA device")
        self.enc, self.device = enc, dev
        S = enc.image_size
        self.images = torch.zeros(n_tiles, channels, S, S, dtype=in_dtype, device=dev)
        # eager warm-up: packs / refreshes the weights, sizes the workspace, sets the kernels' shared-memory attributes
        enc.encode_images(self.images, out_dtype)
        torch.cuda.synchronize(dev)
        self._pk = enc._packed
        self._ws = enc._ws[dev]       # the graph holds its address: keep it alive even if the encoder grows a new one
        self.graph = torch.cuda.CUDAGraph()
        frozen = enc._frozen
        enc._frozen = True            # no weight refresh (a multi-tensor copy) inside the captured region
        try:
            with torch.cuda.graph(self.graph):
                self.features = enc.encode_images(self.images, out_dtype)
        finally:
            enc._frozen = frozen
        if enc._ws[dev] is not self._ws:
            raise RuntimeError("radvlm_b200: the workspace was re-allocated during capture")

    def replay(self) -> torch.Tensor:
        if self.enc._packed is not self._pk:
            raise RuntimeError("radvlm_b200: the packed weights were rebuilt (new parameter storage) after this graph "
                               "was captured; capture again")
        self.graph.replay()
        return self.features

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        if tuple(images.shape) != tuple(self.images.shape):
            raise ValueError("graph captured for %s, got %s" % (tuple(self.images.shape), tuple(images.shape)))
        self.images.copy_(images, non_blocking=True)
        return self.replay()


# =================================================================================================
# Training mode: encode_images under autograd (BASELINE config 5; mm_tunable_parts = vision tower + projector,
# train.py:1642-1665).  The forward keeps what radvlm_siglip_tower_forward_train saves; the backward runs
# radvlm_projector_backward + radvlm_siglip_tower_backward and hands the fp32 gradient buffers back to autograd as
# the gradients of the module Parameters (so accumulation, DDP hooks and optimizers see ordinary .grad tensors).
# =================================================================================================
class _GradBuffers:
    """fp32 accumulators with the packed-weight shapes + the C structs pointing at them (NULL = frozen)."""

    def __init__(self, enc: "B200VisionEncoder", pk: PackedWeights, dev):
        tw = pk.tower
        D, I, NL = tw.hidden, tw.intermediate, tw.num_layers
        tsd = dict(enc.tower_module.named_parameters())
        psd = dict(enc.projector_module.named_parameters())
        pre = "vision_model."
        self.bufs: Dict[str, torch.Tensor] = {}
        # ONE flat fp32 allocation carved into 256-byte aligned views, in the order [embeddings | layer 0 .. NL-1 |
        # projector]: the gradients of a layer range (+ the embeddings when it starts at layer 0) are one contiguous
        # slice, which the data-parallel all-reduce reduces IN PLACE (no flatten / copy back).  The object is cached on
        # the encoder and re-zeroed per backward (one memset) instead of re-allocated.
        cap = NL * (3 * D * D + D * D + 2 * D * I + 9 * D + I + 64 * 12) + D * tw.patch_k_pad + D + pk.tokens_per_tile * D \
            + pk.proj_hidden * (D + pk.proj_hidden + 2) + 64 * 8
        flat = torch.zeros(cap, dtype=torch.float32, device=dev)
        self.flat = flat
        self.spans: Dict[str, tuple] = {}    # "emb" / "l<i>" / "proj" -> (start, end) element offsets in flat
        cursor = [0]

        def alloc(key, shape, *src_names, table=tsd):
            if not any(table[n].requires_grad for n in src_names):
                return None
            n = 1
            for d in shape:
                n *= int(d)
            t = flat[cursor[0]:cursor[0] + n].view(shape)
            cursor[0] += (n + 63) // 64 * 64
            self.bufs[key] = t
            return t.data_ptr()

        self.tower = _lib.SiglipGrads()
        c0 = cursor[0]
        self.tower.patch_w = alloc("patch_w", (D, tw.patch_k_pad), pre + "embeddings.patch_embedding.weight")
        self.tower.patch_b = alloc("patch_b", (D,), pre + "embeddings.patch_embedding.bias")
        self.tower.pos_embed = alloc("pos", (pk.tokens_per_tile, D), pre + "embeddings.position_embedding.weight")
        self.spans["emb"] = (c0, cursor[0])
        self.layers = (_lib.VitLayerGrads * NL)()
        for i in range(NL):
            lp = pre + "encoder.layers.%d." % i
            L = self.layers[i]
            c0 = cursor[0]
            L.ln1_gamma = alloc("l%d.ln1_g" % i, (D,), lp + "layer_norm1.weight")
            L.ln1_beta = alloc("l%d.ln1_b" % i, (D,), lp + "layer_norm1.bias")
            L.qkv_w = alloc("l%d.qkv_w" % i, (3 * D, D), lp + "self_attn.q_proj.weight", lp + "self_attn.k_proj.weight",
                            lp + "self_attn.v_proj.weight")
            L.qkv_b = alloc("l%d.qkv_b" % i, (3 * D,), lp + "self_attn.q_proj.bias", lp + "self_attn.k_proj.bias",
                            lp + "self_attn.v_proj.bias")
            L.out_w = alloc("l%d.out_w" % i, (D, D), lp + "self_attn.out_proj.weight")
            L.out_b = alloc("l%d.out_b" % i, (D,), lp + "self_attn.out_proj.bias")
            L.ln2_gamma = alloc("l%d.ln2_g" % i, (D,), lp + "layer_norm2.weight")
            L.ln2_beta = alloc("l%d.ln2_b" % i, (D,), lp + "layer_norm2.bias")
            L.fc1_w = alloc("l%d.fc1_w" % i, (I, D), lp + "mlp.fc1.weight")
            L.fc1_b = alloc("l%d.fc1_b" % i, (I,), lp + "mlp.fc1.bias")
            L.fc2_w = alloc("l%d.fc2_w" % i, (D, I), lp + "mlp.fc2.weight")
            L.fc2_b = alloc("l%d.fc2_b" % i, (D,), lp + "mlp.fc2.bias")
            self.spans["l%d" % i] = (c0, cursor[0])
        self.tower.layers = C.cast(self.layers, C.POINTER(_lib.VitLayerGrads))
        self.tower_trainable = bool(self.bufs)
        P = pk.proj_hidden
        self.proj = _lib.ProjectorGrads()
        c0 = cursor[0]
        self.proj.w1 = alloc("p.w1", (P, D), "0.weight", table=psd)
        self.proj.b1 = alloc("p.b1", (P,), "0.bias", table=psd)
        self.proj.w2 = alloc("p.w2", (P, P), "2.weight", table=psd)
        self.proj.b2 = alloc("p.b2", (P,), "2.bias", table=psd)
        self.spans["proj"] = (c0, cursor[0])
        self._D, self._NL, self._pre = D, NL, pre
        self._locs: Dict[tuple, object] = {}

    def layer_slice(self, lo: int, hi: int, include_embeddings: bool) -> torch.Tensor:
        """The contiguous slice of the flat buffer holding the gradients of layers [lo, hi) (+ the embeddings)."""
        a = self.spans["emb"][0] if include_embeddings else self.spans["l%d" % lo][0]
        b = self.spans["l%d" % (hi - 1)][1] if hi > lo else self.spans["emb"][1]
        return self.flat[a:b]

    def projector_slice(self) -> torch.Tensor:
        a, b = self.spans["proj"]
        return self.flat[a:b]

    def _locate(self, name: str, is_tower: bool):
        """(buffer key, q/k/v index or None) of one module Parameter, or None when the path does not use it."""
        D, b = self._D, self.bufs
        if not is_tower:
            key = {"0.weight": "p.w1", "0.bias": "p.b1", "2.weight": "p.w2", "2.bias": "p.b2"}.get(name)
            return (key, None) if key in b else None
        emb = {"embeddings.patch_embedding.weight": "patch_w", "embeddings.patch_embedding.bias": "patch_b",
               "embeddings.position_embedding.weight": "pos"}
        rel = name[len(self._pre):] if name.startswith(self._pre) else name
        if rel in emb:
            return (emb[rel], None) if emb[rel] in b else None
        if rel.startswith("encoder.layers."):
            i, leaf = rel[len("encoder.layers."):].split(".", 1)
            if int(i) >= self._NL:
                return None
            k = "l%d." % int(i)
            simple = {"layer_norm1.weight": "ln1_g", "layer_norm1.bias": "ln1_b", "layer_norm2.weight": "ln2_g",
                      "layer_norm2.bias": "ln2_b", "self_attn.out_proj.weight": "out_w",
                      "self_attn.out_proj.bias": "out_b", "mlp.fc1.weight": "fc1_w", "mlp.fc1.bias": "fc1_b",
                      "mlp.fc2.weight": "fc2_w", "mlp.fc2.bias": "fc2_b"}
            if leaf in simple:
                return (k + simple[leaf], None) if k + simple[leaf] in b else None
            for j, proj in enumerate(("q_proj", "k_proj", "v_proj")):
                if leaf == "self_attn.%s.weight" % proj:
                    return (k + "qkv_w", j) if k + "qkv_w" in b else None
                if leaf == "self_attn.%s.bias" % proj:
                    return (k + "qkv_b", j) if k + "qkv_b" in b else None
        return None   # post_layernorm, head, the dropped 27th layer

    def export(self, param_names, params, needs):
        """Gradients of the module Parameters (their shapes / dtypes; None when frozen or off the path, like the
        reference's autograd, so AdamW skips those entirely).  The whole flat buffer is converted / copied ONCE per
        dtype and the results are views of that copy: one kernel instead of one cast per Parameter (420 of them), and
        nothing aliases the cached accumulators, which the next backward zeroes while .grad may still be summing
        micro-batches."""
        D = self._D
        conv: Dict[torch.dtype, torch.Tensor] = {}
        outs = []
        for (name, is_tower), p, need in zip(param_names, params, needs):
            loc = self._locs.get((name, is_tower), False)
            if loc is False:
                loc = self._locs[(name, is_tower)] = self._locate(name, is_tower)
            if not need or not p.requires_grad or loc is None:
                outs.append(None)
                continue
            key, j = loc
            src = conv.get(p.dtype)
            if src is None:
                src = conv[p.dtype] = self.flat.clone() if p.dtype == torch.float32 else self.flat.to(p.dtype)
            buf = self.bufs[key]
            start = (buf.data_ptr() - self.flat.data_ptr()) // 4
            t = src[start:start + buf.numel()].view(buf.shape)
            if key == "patch_w":
                t = t[:, :p[0].numel()]
            elif j is not None:
                t = t[j * D:(j + 1) * D]
            outs.append(t.reshape(p.shape))
        return outs


class _EncodeImagesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, images, out_dtype, n_tower, *params):
        dev = images.device
        pk = enc.packed(dev, training=True)
        lib = _lib.load()
        n = images.shape[0]
        T, Hp, D = pk.tokens_per_tile, pk.proj_hidden, pk.hidden
        kernel_out = out_dtype if out_dtype in _DT else torch.float32
        out = torch.empty(n, T, Hp, dtype=kernel_out, device=dev)
        chunks = []
        with torch.cuda.device(dev):
            stream = _stream_ptr(dev)
            step = max(1, enc.max_tiles_per_call)
            for s in range(0, n, step):
                m = min(step, n - s)
                need = lib.radvlm_encode_workspace_bytes(C.byref(pk.tower), C.byref(pk.projector), m)
                sbytes = lib.radvlm_tower_saved_bytes(C.byref(pk.tower), m)
                if need == 0 or sbytes == 0:
                    raise _lib.RadvlmError(_lib.ERR_BAD_ARGUMENT, _lib.last_error())
                ws = enc._workspace(dev, need)
                saved = torch.empty(sbytes, dtype=torch.uint8, device=dev)
                _lib.check(lib.radvlm_siglip_tower_forward_train(
                    C.byref(pk.tower), images[s:s + m].data_ptr(), _DT[images.dtype], m, saved.data_ptr(), sbytes,
                    ws.data_ptr(), ws.numel(), stream))
                hid_ptr = saved.data_ptr() + lib.radvlm_tower_saved_hidden_offset(C.byref(pk.tower), m)
                _lib.check(lib.radvlm_projector_forward(
                    C.byref(pk.projector), hid_ptr, m * T, out[s:s + m].data_ptr(), _DT[kernel_out], ws.data_ptr(),
                    ws.numel(), stream))
                chunks.append((s, m, saved))
        ctx.enc, ctx.pk, ctx.images, ctx.chunks, ctx.n_tower = enc, pk, images, chunks, n_tower
        ctx.param_names = enc._param_names()
        ctx.params = params
        return out if out.dtype == out_dtype else out.to(out_dtype)

    @staticmethod
    def backward(ctx, d_out):
        enc, pk, images = ctx.enc, ctx.pk, ctx.images
        dev = images.device
        lib = _lib.load()
        T, Hp, D = pk.tokens_per_tile, pk.proj_hidden, pk.hidden
        d_out = d_out.to(device=dev, dtype=torch.bfloat16).contiguous()
        gb = enc._grad_buffers(pk, dev)
        reducer = None
        if enc.grad_allreduce_group is not False:   # False = off; None = default process group (when initialised)
            from .dist import GradientAllReducer
            import torch.distributed as tdist
            if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(enc.grad_allreduce_group) > 1:
                reducer = GradientAllReducer(enc.grad_allreduce_group, enc.grad_bucket_bytes, average=True)
        with torch.cuda.device(dev):
            stream = _stream_ptr(dev)
            for s, m, saved in ctx.chunks:
                rows = m * T
                pws = lib.radvlm_projector_backward_workspace_bytes(C.byref(pk.projector), rows)
                tws = lib.radvlm_tower_backward_workspace_bytes(C.byref(pk.tower), m) if gb.tower_trainable else 0
                ws = enc._workspace(dev, max(pws, tws))
                hid_ptr = saved.data_ptr() + lib.radvlm_tower_saved_hidden_offset(C.byref(pk.tower), m)
                d_hidden = torch.empty(rows, D, dtype=torch.float32, device=dev) if gb.tower_trainable else None
                _lib.check(lib.radvlm_projector_backward(
                    C.byref(pk.projector), C.byref(gb.proj), hid_ptr, d_out[s:s + m].data_ptr(), rows,
                    None if d_hidden is None else d_hidden.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                last_chunk = (s, m) == (ctx.chunks[-1][0], ctx.chunks[-1][1])
                if last_chunk and reducer is not None:   # projector gradients are final: reduce them first
                    reducer.submit_flat(gb.projector_slice())
                if gb.tower_trainable:
                    NL = pk.tower.num_layers
                    step_l = max(1, enc.backward_layers_per_range) if (last_chunk and reducer is not None) else NL
                    hi = NL
                    while hi > 0:
                        lo = max(0, hi - step_l)
                        _lib.check(lib.radvlm_siglip_tower_backward_range(
                            C.byref(pk.tower), C.byref(gb.tower), images[s:s + m].data_ptr(), _DT[images.dtype], m,
                            saved.data_ptr(), saved.numel(), d_hidden.data_ptr(), ws.data_ptr(), ws.numel(), lo, hi,
                            stream))
                        if last_chunk and reducer is not None:   # overlap: these layers are done, the next range runs
                            reducer.submit_flat(gb.layer_slice(lo, hi, include_embeddings=(lo == 0)))
                        hi = lo
        if reducer is not None:
            reducer.finish()
        grads = gb.export(ctx.param_names, ctx.params, ctx.needs_input_grad[4:])
        ctx.chunks = None
        return (None, None, None, None, *grads)


def _encode_images_train(self: "B200VisionEncoder", images: torch.Tensor, out_dtype=None) -> torch.Tensor:
    if images.dim() != 4:
        raise ValueError("encode_images expects [n, C, S, S], got %s" % (tuple(images.shape),))
    if not torch.cuda.is_available():
        raise RuntimeError("radvlm_b200: no CUDA device; the encode path has no CPU fallback")
    dev = next(self.projector_module.parameters()).device
    out_dtype = out_dtype or images.dtype
    if images.dtype not in _DT:
        images = images.float()
    images = images.detach().to(dev).contiguous()
    tparams = [p for _, p in self.tower_module.named_parameters()]
    pparams = [p for _, p in self.projector_module.named_parameters()]
    return _EncodeImagesFn.apply(self, images, out_dtype, len(tparams), *tparams, *pparams)


def _param_names(self: "B200VisionEncoder"):
    return [(n, True) for n, _ in self.tower_module.named_parameters()] + \
           [(n, False) for n, _ in self.projector_module.named_parameters()]


def _grad_buffers(self: "B200VisionEncoder", pk: PackedWeights, dev) -> _GradBuffers:
    """The cached gradient accumulators (one flat fp32 allocation), zeroed; rebuilt when the packed weights or the
    set of trainable Parameters changed."""
    key = (id(pk), torch.device(dev), tuple(p.requires_grad for p in self._source_tensors()))
    gb = getattr(self, "_gb", None)
    if gb is None or self._gb_key != key:
        self._gb = gb = None   # release the old allocation first
        gb = _GradBuffers(self, pk, dev)   # freshly zero-filled
        self._gb, self._gb_key = gb, key
    else:
        gb.flat.zero_()
    return gb


B200VisionEncoder.encode_images_train = _encode_images_train
B200VisionEncoder._grad_buffers = _grad_buffers
B200VisionEncoder._param_names = _param_names


def flops_per_tile(hidden=1152, inter=4304, layers=26, seq=729, patch_k=588, proj=3584) -> float:
    """Algorithmic forward FLOPs per tile (2*M*N*K, unpadded)."""
    per_layer = 2 * seq * (3 * hidden * hidden + hidden * hidden + 2 * hidden * inter) + 4 * seq * seq * hidden
    return 2.0 * seq * patch_k * hidden + layers * per_layer + 2.0 * seq * (hidden * proj + proj * proj)


assert math.isclose(flops_per_tile() / 1e9, 666.55, rel_tol=1e-3)
