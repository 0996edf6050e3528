#!/usr/bin/env python
"""bench.py — visual tokens/s of the anyres_max_9 multimodal encode path (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus 1 --steps K ...   # the reference's CPU implementation (host cores)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: B chest-X-ray-shaped images
(1024x1024 grayscale replicated to RGB, uint8) per GPU -> fused preprocess (anyres tiling, 10 tiles) ->
SigLIP-so400m/14-384 tower (26 executed layers) -> mlp2x_gelu projector -> unpad / newline merge -> splice into
Qwen2-width (3584) input embeddings; 7371 visual tokens per image (BASELINE.json configs[1], replicated B
times per GPU; at N > 1 the images are sharded by rank and the embeddings are all-gathered, configs[2]).

`value`   : inputs resident in HBM when the timed region starts (device-timed, max over ranks).
`e2e`     : same metric through the public API with HOST (pinned) uint8 images, H2D inside the timed region and a
            D2H read of a slice of the result every step.
`roofline`: the dominant kernel (the fc2 tcgen05 GEMM) timed live with CUDA events on the launch stream, in a
            profiled pass of its own (the timed regions run with that instrumentation off).
`cpu_baseline`: the oracle port on the host cores, one whole image (10 tiles, nothing extrapolated).
Extra keys of the same line: `train` (configs[4]: forward + backward, N > 1 with the NCCL gradient all-reduce),
`c3_strong` (configs[2]: global batch 256, strong scaling), `batch1` (one 10-tile image per call), `gather_check`
(N > 1: checksum proof that every rank received every rank's rows).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOKENS_PER_IMAGE = 7371          # 729 + 81 * 82 for a 1024x1024 image
TILES_PER_IMAGE = 10
IMG = 1024
METRIC = "visual tokens/sec (anyres_max_9, SigLIP+projector)"
STEP_TIMES = os.environ.get("RADVLM_BENCH_STEP_TIMES", "0") == "1"   # diagnostics: per-step device / host times on stderr


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def _workload_config(args, world):
    return {
        "workload": "BASELINE.json configs[1] (anyres_max_9 encode of a synthetic 1024x1024 CXR: 10 tiles -> 7371 visual "
                    "tokens, SigLIP-so400m/14-384 26 layers + mlp2x_gelu 1152->3584->3584, unpad/newline merge, splice "
                    "with a 32-token prompt) x %d images per GPU per step" % args.batch
                    + ("" if world == 1 else "; images sharded by rank, inputs_embeds all-gathered over NCCL (configs[2])"),
        "images_per_gpu_per_step": args.batch,
        "global_images_per_step": args.batch * world,
        "tiles_per_image": TILES_PER_IMAGE,
        "visual_tokens_per_image": TOKENS_PER_IMAGE,
        "prompt_tokens": 32,
        "parallelism": "dp%d (image-sharded)" % world,
        "gather": ("none" if world == 1 else
                   {"peer": "fused merge + scatter kernel over peer memory (radvlm_merge_splice_scatter)",
                    "nccl": "NCCL all-gather of inputs_embeds, asynchronous"}.get(
                       getattr(args, "gather", "auto"),
                       "merge kernel writes into peer-mapped memory, copy engines push the slice to every peer over NVLink "
                       "(dist.PeerGather mode=ce)")),
        "l2": "per-step working set (0.83 GB bf16 weights + >1 GB activations) exceeds the 126 MB L2; input images rotate over 3 buffers",
    }


# ----------------------------------------------------------------------------------------------------------------
# CPU arms (oracle port / real reference)
# ----------------------------------------------------------------------------------------------------------------
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms run on rank 0 alone and may use the whole host."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


def _cpu_sample(use_reference: bool):
    """One bounded sample of the CPU path = ONE whole image of the workload, nothing extrapolated: preprocess a
    1024x1024 image (10 tiles), tower + projector on all 10 tiles (fp32), merge + splice with a 32-token prompt.
    Returns (seconds per image, description, kind, cores)."""
    import numpy as np
    import torch
    import golden_inputs as gi
    cores = _use_all_host_threads()
    rng = np.random.default_rng(0)
    gray = rng.integers(0, 256, size=(IMG, IMG), dtype=np.uint8)
    ids = torch.randint(1, 1000, (1, 32), generator=torch.Generator().manual_seed(5))
    ids[0, 7] = -200
    if use_reference:
        from PIL import Image
        st = _cpu_state(True)
        host, ref = st["host"], st["ref"]
        pil = Image.fromarray(np.repeat(gray[:, :, None], 3, axis=2))
        t0 = time.perf_counter()
        tiles = ref.mm_utils.process_anyres_image(pil, host.get_vision_tower().image_processor, gi.PINPOINTS)
        t_pre = time.perf_counter() - t0
        t0 = time.perf_counter()
        with torch.no_grad():   # the reference's own entry point: encode_images (10 tiles) + merge + splice inside
            out = host.prepare_inputs_labels_for_multimodal(ids, None, torch.ones_like(ids, dtype=torch.bool), None,
                                                            ids.clone(), [tiles], ["image"], [(IMG, IMG)])
        t_enc = time.perf_counter() - t0
        assert out[4].shape[1] == 31 + TOKENS_PER_IMAGE
        kind = "reference"
        desc = ("1 whole image: process_anyres_image (%.2fs) + prepare_inputs_labels_for_multimodal = SigLIP tower + "
                "projector fp32 on all 10 tiles + merge/splice (%.2fs)" % (t_pre, t_enc))
    else:
        from oracle import encoder_oracle as eo
        from oracle import resample_oracle as ro
        st = _cpu_state(False)
        t0 = time.perf_counter()
        tiles = torch.from_numpy(ro.process_anyres_image(gray, gi.PINPOINTS))
        t_pre = time.perf_counter() - t0
        t0 = time.perf_counter()
        with torch.no_grad():
            feats = eo.encode_images(st["tsd"], st["psd"], tiles)
            merged = eo.merge_image(feats, (IMG, IMG), st["newline"], gi.PINPOINTS)
            emb = eo.prepare_inputs_labels(st["table"], [merged], ids, torch.ones_like(ids, dtype=torch.bool), ids.clone(),
                                           32768, False)[0]
        t_enc = time.perf_counter() - t0
        assert emb.shape[1] == 31 + TOKENS_PER_IMAGE
        kind = "port"
        desc = ("1 whole image: preprocess (%.2fs) + SigLIP tower + projector fp32 on all 10 tiles + merge/splice (%.2fs), "
                "oracle port of the reference path" % (t_pre, t_enc))
    return t_pre + t_enc, desc, kind, cores


_CPU_STATE = {}


def _cpu_state(use_reference: bool):
    import torch
    key = "ref" if use_reference else "port"
    if key in _CPU_STATE:
        return _CPU_STATE[key]
    if use_reference:
        from oracle.ref_loader import build_reference_host
        host, ref = build_reference_host(vocab=4096, hidden_size=3584, seed=0)
        st = {"host": host, "ref": ref}
    else:
        from radvlm_b200 import synthetic
        host = synthetic.build_host(hidden_size=3584, vocab=4096, seed=0, dtype=torch.float32, device="cpu")
        st = {"tsd": host.model.vision_tower.vision_tower.state_dict(), "psd": host.model.mm_projector.state_dict(),
              "newline": host.model.image_newline.detach(), "table": host.model.embed_tokens.weight.detach()}
    _CPU_STATE[key] = st
    return st


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on the box's host cores (the real Python reference when
    /root/reference exists, else the oracle port).  Each step = one whole 1024x1024 image (10 tiles, nothing
    extrapolated).  Under torchrun rank 0 alone runs it, with ALL host threads; the other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.ref_loader import reference_available
    use_ref = reference_available()
    for _ in range(max(1, min(args.warmup, 2))):   # the first call pages the weights in; more warm-up buys nothing on a CPU
        _cpu_sample(use_ref)
    t_total, per_image_s, desc, kind, cores = 0.0, [], "", "", 1
    for _ in range(args.steps):
        t0 = time.perf_counter()
        s, desc, kind, cores = _cpu_sample(use_ref)
        t_total += time.perf_counter() - t0
        per_image_s.append(s)
    per_image = sum(per_image_s) / len(per_image_s)
    value = TOKENS_PER_IMAGE / per_image
    cfg = _workload_config(args, max(args.gpus, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": kind,
                         "sample": "each step = " + desc + "; value = 7371 tokens / mean seconds per image over %d steps" % args.steps},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU path, rank 0 only, %d threads; each step = 1 image of the workload (the CPU does not scale with the "
                "GPU count: the same value at every N)" % cores,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------------------
def _gpu_eager_incumbent(host, dev, n_tiles=80, reps=3):
    """The GPU incumbent: torch eager bf16 on the same B200, op for op what the reference runs for the tower and the
    projector (siglip_encoder.py:148-305: conv2d patch embedding, LayerNorm, q/k/v Linear, matmul * scale,
    softmax(fp32), matmul, out_proj, tanh-GELU MLP; builder.py:44-48) on the same random-init parameters.  Only the
    dense part: the reference's preprocessing and merge run on the CPU and are not charged to it.  Library kernels
    (cuBLAS / ATen) only - none of this repo's code.  Returns (visual tokens/s at 7371 per 10 tiles, seconds per call)."""
    import torch
    import torch.nn.functional as F
    vm = host.get_model().vision_tower.vision_tower.vision_model
    proj = host.get_model().mm_projector
    heads = host.get_model().vision_tower.config.num_attention_heads
    x0 = torch.randn(n_tiles, 3, 384, 384, device=dev, dtype=torch.bfloat16)

    @torch.no_grad()
    def fwd(px):
        x = vm.embeddings.patch_embedding(px).flatten(2).transpose(1, 2)
        x = x + vm.embeddings.position_embedding.weight[None]
        B, T, D = x.shape
        hd = D // heads
        for layer in vm.encoder.layers:
            h = layer.layer_norm1(x)
            a = layer.self_attn
            q = a.q_proj(h).view(B, T, heads, hd).transpose(1, 2)
            k = a.k_proj(h).view(B, T, heads, hd).transpose(1, 2)
            v = a.v_proj(h).view(B, T, heads, hd).transpose(1, 2)
            w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
            o = torch.matmul(w, v).transpose(1, 2).contiguous().reshape(B, T, D)
            x = x + a.out_proj(o)
            h = layer.layer_norm2(x)
            x = x + layer.mlp.fc2(F.gelu(layer.mlp.fc1(h), approximate="tanh"))
        return proj(x)

    fwd(x0[:8])
    fwd(x0)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fwd(x0)
    e1.record()
    torch.cuda.synchronize(dev)
    sec = e0.elapsed_time(e1) / 1e3 / reps
    return n_tiles / TILES_PER_IMAGE * TOKENS_PER_IMAGE / sec, sec


def _init_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def _prompt(B, dev, Lp=32):
    import torch
    ids = torch.randint(1, 4000, (B, Lp), generator=torch.Generator().manual_seed(5))
    ids[:, 7] = -200
    ids_dev = ids.to(dev)
    mask_dev = torch.ones_like(ids_dev, dtype=torch.bool)
    labels_dev = torch.where(ids_dev < 0, torch.full_like(ids_dev, -100), ids_dev)
    pos_dev = torch.arange(Lp, device=dev)[None].expand(B, -1).contiguous()
    return ids_dev, pos_dev, mask_dev, labels_dev


def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from radvlm_b200 import _lib, mm_utils, synthetic
    from radvlm_b200.encoder import flops_per_tile
    import golden_inputs as gi

    world, rank, local, dev = _init_dist()
    _lib.load()

    B = args.batch
    host = synthetic.build_host(hidden_size=3584, vocab=4096, seed=0, dtype=torch.bfloat16, device=dev)
    from radvlm_b200 import mm_arch
    enc = mm_arch._encoder_for(host).freeze()

    # synthetic chest-X-ray-shaped inputs: 1024x1024 grayscale replicated to RGB, uint8, 3 rotating batches
    rng = np.random.default_rng(1000 + rank)
    n_buf = 3
    host_imgs, dev_imgs = [], []
    for _ in range(n_buf):
        gray = rng.integers(0, 256, size=(B, IMG, IMG, 1), dtype=np.uint8)
        rgb = torch.from_numpy(np.repeat(gray, 3, axis=3).copy()).pin_memory()
        host_imgs.append(rgb)
        dev_imgs.append(rgb.to(dev))
    Lp = 32
    prompt = _prompt(B, dev, Lp)
    # N > 1: every rank's inputs_embeds are gathered on every rank, overlapping the next step's encode.
    #   --gather ce     : dist.PeerGather(mode="ce"): the merge kernel writes into peer-mapped memory, DMA engines push
    #                     the finished slice to every peer over NVLink (no SM, no collective call)
    #   --gather peer   : dist.PeerGather(mode="kernel"): the fused merge + scatter kernel stores every row to all peers
    #   --gather nccl   : asynchronous NCCL all_gather_into_tensor, two alternating slots
    slots = [{"work": None, "emb": None, "out": None} for _ in range(2)]
    step_no = [0]
    peer = [None]
    last = {"emb": None, "slot": None}

    def drain_gathers():
        if peer[0] is not None:
            peer[0].drain()
        for sl in slots:
            if sl["work"] is not None:
                sl["work"].wait()
                sl["work"] = None

    def step(images_u8, pr=prompt):
        n = len(images_u8)
        tiles, sizes, splits, _ = mm_utils.preprocess_anyres_batch(list(images_u8), gi.PINPOINTS, device=dev,
                                                                   dtype=torch.bfloat16)
        if peer[0] is not None:
            last["slot"] = peer[0].peek_slot()
        out = host.prepare_inputs_labels_for_multimodal(pr[0], pr[1], pr[2], None, pr[3],
                                                        list(torch.split(tiles, splits)), ["image"] * n, sizes)
        emb = out[4]
        last["emb"] = emb
        if world > 1 and args.gather in ("peer", "ce", "auto") and n == B:
            if peer[0] is None:   # first step: size the peer buffers from the embeddings, then redo the step into them
                from radvlm_b200.dist import PeerGather
                mode = "kernel" if args.gather == "peer" else "ce"
                try:
                    peer[0] = PeerGather(emb.shape[0] * emb.shape[1], emb.shape[2], emb.dtype, dev, mode=mode)
                except Exception as e:  # no cudaIpc / peer access on this box: the NCCL all-gather does the same job
                    if args.gather != "auto":
                        raise
                    print("bench: peer-memory gather unavailable (%s); using the NCCL all-gather" % e, file=sys.stderr)
                    args.gather = "nccl"
                    return step(images_u8, pr)
                args.gather = "peer" if mode == "kernel" else "ce"
                host.radvlm_b200_gather = peer[0]
                return step(images_u8, pr)
        elif world > 1 and n == B:
            sl = slots[step_no[0] & 1]
            step_no[0] += 1
            if sl["work"] is not None:
                sl["work"].wait()
            if sl["out"] is None:
                sl["out"] = torch.empty((world,) + tuple(emb.shape), dtype=emb.dtype, device=dev)
            sl["emb"] = emb
            sl["work"] = dist.all_gather_into_tensor(sl["out"], emb, async_op=True)
            last["slot"] = sl
        return emb

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, read_back=False):
        """device time of `steps` calls of fn(i), bracketed by barrier + synchronize, max over ranks"""
        rb_slots = [(torch.empty(1, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(2)] if read_back else None
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        chk = 0.0
        marks = []
        pending = None
        for i in range(steps):
            emb = fn(i)
            if read_back:
                # D2H read of a value computed from every step's result: copied asynchronously into pinned memory and
                # consumed one step later (after the next step has been enqueued), so the read does not drain the stream
                slot = rb_slots[i & 1]
                slot[0].copy_(emb[0, -1, :8].float().sum().reshape(1), non_blocking=True)
                slot[1].record()
                if pending is not None:
                    pending[1].synchronize()
                    chk += float(pending[0][0])
                pending = slot
            if STEP_TIMES:
                marks.append((torch.cuda.Event(enable_timing=True), time.time()))
                marks[-1][0].record()
        if pending is not None:
            pending[1].synchronize()
            chk += float(pending[0][0])
        drain_gathers()   # the timed region ends when the last exchange has landed
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if STEP_TIMES and marks:   # diagnostics: device time and host enqueue time of every step of this region
            dev_ms = [e0.elapsed_time(marks[0][0])] + [marks[k - 1][0].elapsed_time(marks[k][0]) for k in range(1, len(marks))]
            host_ms = [(marks[0][1] - t_wall0) * 1e3] + [(marks[k][1] - marks[k - 1][1]) * 1e3 for k in range(1, len(marks))]
            print("step times: device ms %s | host enqueue ms %s" % (" ".join("%.1f" % v for v in dev_ms),
                                                                     " ".join("%.1f" % v for v in host_ms)), file=sys.stderr)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t_wall0, time.time(), chk

    for i in range(args.warmup):
        step(dev_imgs[i % n_buf])
    drain_gathers()
    sync()

    # ---- device-resident arm: the headline `value`.  The per-class CUDA-event instrumentation is OFF here.
    clocks = ClockSampler(local) if rank == 0 else None
    from radvlm_b200 import mm_arch as _mm_arch
    _enc = _mm_arch._encoder_for(host)
    g0 = (_enc.n_graph_replays, _enc.n_graph_captures, _enc.n_eager_launches)
    ms, tw0, tw1, _ = timed(lambda i: step(dev_imgs[i % n_buf]), args.steps)
    clk = clocks.stop(tw0, tw1) if clocks else None
    # how the encode calls of the timed region were issued: replays of captured CUDA graphs / captures / eager launches
    graph_stats = {"enabled": bool(_enc.graph_mode), "encode_calls_replayed": _enc.n_graph_replays - g0[0],
                   "graphs_captured": _enc.n_graph_captures - g0[1], "encode_calls_eager": _enc.n_eager_launches - g0[2],
                   "graphs_captured_before": g0[1], "capture_host_ms_total": _enc.graph_capture_seconds * 1e3}

    # ---- end-to-end arm: pinned host uint8 in, slice of the result out, every step
    for i in range(min(2, args.warmup)):
        step(host_imgs[i % n_buf])
    ms_e2e, _, _, _ = timed(lambda i: step(host_imgs[i % n_buf]), args.steps, read_back=True)

    # ---- profiled pass (separate from the timed regions): live duration / launch count of every kernel class,
    #      CUDA events recorded around each launch on the launch stream (radvlm_profile_*)
    prof_steps = max(1, min(args.steps, 5))
    _lib.profile_enable(True)
    _lib.profile_read()
    ms_prof, _, _, _ = timed(lambda i: step(dev_imgs[i % n_buf]), prof_steps)
    prof_ms, prof_n = _lib.profile_read()
    _lib.profile_enable(False)

    # ---- N > 1: prove the exchange delivered every rank's rows (checksum of checksums over the gathered buffer)
    gather_check = None
    if world > 1:
        emb = step(dev_imgs[0])
        drain_gathers()
        torch.cuda.synchronize(dev)
        rows = emb.shape[0] * emb.shape[1]
        if peer[0] is not None:
            gathered = peer[0].gathered(last["slot"])[:, :rows]
        else:
            gathered = last["slot"]["out"].view(world, rows, -1)
        mine = emb.reshape(rows, -1).view(torch.int32).sum(dtype=torch.int64).reshape(1)    # bit-level checksum
        theirs = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(theirs, mine)
        seen = torch.stack([gathered[r].reshape(rows, -1).view(torch.int32).sum(dtype=torch.int64) for r in range(world)])
        ok = torch.tensor([1 if torch.equal(seen, theirs) and len(set(theirs.tolist())) == world else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gather_check = "ok" if int(ok.item()) == 1 else "FAILED"
        gather_check += (": every rank holds all %d ranks' [%d, %d] slices, int32-sum checksums equal to the owners' "
                         "(all ranks agree, %d distinct batches)" % (world, rows, emb.shape[2], len(set(theirs.tolist()))))

    # ---- strong scaling, BASELINE.json configs[2]: a fixed global batch of 256 images, 256 / N per rank in
    #      micro-batches of B, embeddings exchanged per micro-batch
    c3 = None
    if args.c3 and 256 % (world * B) == 0:
        micro = 256 // (world * B)
        reps = 2
        timed(lambda i: step(dev_imgs[i % n_buf]), micro)   # one untimed global batch
        ms_c3, _, _, _ = timed(lambda i: step(dev_imgs[i % n_buf]), micro * reps)
        c3 = {"value": 256 * TOKENS_PER_IMAGE / (ms_c3 / reps / 1e3), "unit": "tokens/s", "scaling": "strong",
              "ms_per_global_batch": ms_c3 / reps, "global_batch": 256, "images_per_rank": 256 // world,
              "micro_batches_per_rank": micro, "repeats": reps,
              "workload": "BASELINE.json configs[2]: 256 synthetic 1024x1024 CXR = 2560 tiles -> 1,886,976 visual tokens "
                          "per global batch, image-sharded over %d GPU(s)" % world}

    # ---- batch 1 (the reference trains and serves at per-device batch 1: ONE 10-tile image per call)
    b1 = None
    if args.batch1:
        p1 = _prompt(1, dev, Lp)
        one = [dev_imgs[k][j:j + 1] for k in range(n_buf) for j in range(min(B, 4))]
        saved_gather = getattr(host, "radvlm_b200_gather", None)
        host.radvlm_b200_gather = None
        for i in range(5):
            step(one[i % len(one)], p1)
        reps1 = 40
        ms_b1, _, _, _ = timed(lambda i: step(one[i % len(one)], p1), reps1)
        host.radvlm_b200_gather = saved_gather
        peaks1, _ = _peaks()
        pk1 = float(peaks1.get("bf16_tflops_sustained", peaks1["bf16_tflops"]))
        b1 = {"ms_per_image": ms_b1 / reps1, "value": TOKENS_PER_IMAGE / (ms_b1 / reps1 / 1e3), "unit": "tokens/s",
              "tiles_per_call": TILES_PER_IMAGE, "calls": reps1,
              "path_frac_of_peak": flops_per_tile() * TILES_PER_IMAGE / (ms_b1 / reps1 / 1e3) / 1e12 / pk1,
              "what": "one 1024x1024 image per call through preprocess + prepare_inputs_labels_for_multimodal, device "
                      "resident, per rank (no exchange)"}
        # encode_images alone on those 10 tiles: eager launches (~140 kernel launches + their TMA descriptors per call)
        # against ONE replay of the call captured in a CUDA graph (B200VisionEncoder.capture)
        if world == 1:   # N = 1 only: a rank-local failure inside `timed` (barrier + all-reduce) would hang the others
            try:
                from radvlm_b200 import mm_arch
                t1, _, _, _ = mm_utils.preprocess_anyres_batch(list(one[0]), gi.PINPOINTS, device=dev, dtype=torch.bfloat16)
                enc1 = mm_arch._encoder_for(host)
                with torch.no_grad():
                    mode1, enc1.graph_mode = enc1.graph_mode, False
                    ref1 = host.encode_images(t1)
                    ms_eager, _, _, _ = timed(lambda i: host.encode_images(t1), reps1)
                    enc1.graph_mode = mode1
                    for _ in range(6):
                        same_auto = bool(torch.equal(host.encode_images(t1), ref1))
                    ms_auto, _, _, _ = timed(lambda i: host.encode_images(t1), reps1)
                    graphed = enc1.capture(int(t1.shape[0]), in_dtype=torch.bfloat16)
                    same = bool(torch.equal(graphed(t1), ref1))
                    ms_graph, _, _, _ = timed(lambda i: graphed(t1), reps1)
                b1["encode_only"] = {"eager_ms": ms_eager / reps1, "auto_graph_ms": ms_auto / reps1,
                                     "auto_graph_equals_eager": same_auto,
                                     "cuda_graph_ms": ms_graph / reps1, "graph_equals_eager": same,
                                     "tma_descriptor_cache": dict(zip(("hits", "misses"), _lib.tmap_cache_stats()))}
            except Exception as e:   # a failed capture must not cost the bench line
                b1["encode_only"] = {"error": "%s: %s" % (type(e).__name__, e)}

    line = None
    if rank == 0:
        peaks, peak_src = _peaks()
        tokens_step = B * world * TOKENS_PER_IMAGE
        value = tokens_step / (ms / args.steps / 1e3)
        e2e = tokens_step / (ms_e2e / args.steps / 1e3)
        # ---- rooflines.  Algorithmic work per launch / average launch duration measured live
        # (CUDA events on the launch stream, radvlm_profile_*).  `roofline` is the dominant single kernel by time:
        # the fc2 GEMM ([tiles*729, 4304] x [1152, 4304]^T with the fp32 residual epilogue).
        attn_flops = 26 * 4.0 * 729 * 729 * 1152
        gemm_flops_step = (flops_per_tile() - attn_flops) * TILES_PER_IMAGE * B
        gemm_total_ms = sum(v for k, v in prof_ms.items() if k.startswith("gemm"))
        gemm_ms_step = gemm_total_ms / prof_steps
        achieved_all = gemm_flops_step / (gemm_ms_step * 1e-3) / 1e12 if gemm_ms_step > 0 else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        hbm_peak = float(peaks.get("hbm_gbs", 6545.0))
        total_ms = sum(prof_ms.values()) or 1.0
        rows_step = B * TILES_PER_IMAGE * 729
        olp_calls = max(prof_n.get("gemm_out", 0), 1) / prof_steps / 26.0   # tower calls per step (weights re-read per call)
        traffic = _ncu_traffic()

        def per_launch(cls, work_step):
            n = max(prof_n.get(cls, 0), 1)
            return work_step * prof_steps / n, prof_ms.get(cls, 0.0) / n, n / prof_steps

        def traffic_of(cls, launches_per_step):
            """DRAM bytes of one launch from the ncu capture, if it was taken at this run's tiles-per-launch."""
            t = traffic.get(cls)
            tower_calls = launches_per_step / (52.0 if cls == "layernorm" else 26.0)
            if not t or tower_calls <= 0:
                return None
            if abs(B * TILES_PER_IMAGE / tower_calls - traffic.get("tiles_per_launch", -1)) > 1e-6:
                return None
            return (t["dram_read_mb"] + t["dram_write_mb"]) * 1e6

        def tensor_roofline(cls, kernel, flops_step_cls):
            fl, ms_l, lps = per_launch(cls, flops_step_cls)
            ach = fl / (ms_l * 1e-3) / 1e12 if ms_l > 0 else 0.0
            return {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": traffic_of(cls, lps), "algorithmic_flops_per_launch": fl,
                    "ms_per_launch": ms_l, "launches_per_step": lps,
                    "share_of_step": prof_ms.get(cls, 0.0) / total_ms}

        fc2 = tensor_roofline("gemm_fc2", "tcgen05 cta_group::2 GEMM as fc2 ([rows,4304] x [1152,4304]^T, fp32 residual "
                              "epilogue)", 26 * 2.0 * rows_step * 1152 * 4304)
        fc2["peak_source"] = "%s bf16_tflops_sustained (kernel timed inside a long step)" % peak_src
        fc2["measured"] = ("separate profiled pass of %d steps right after the timed regions (CUDA events around every "
                           "launch; the timed regions run with the instrumentation off)" % prof_steps)
        folded = prof_n.get("layernorm", 0) == 0
        # bytes per output element: bf16 A 2 + (delta epilogue, the default with the fold: bf16 stream copy in 2, bf16
        # branch out 2, bf16 LayerNorm-2 input out 2 = 8 | fp32 stream in / out + bf16 copy = 12 | without the fold 10)
        delta = folded and os.environ.get("RADVLM_B200_OUTPROJ", "") != "f32"
        out_bpe = 8.0 if delta else (12.0 if folded else 10.0)
        ob, oms, olps = per_launch("gemm_out", 26 * (out_bpe * rows_step * 1152 + 2.0 * 1152 * 1152 * olp_calls))
        out_hbm = {"kernel": "out_proj GEMM (%s)" % ("bf16 branch + LayerNorm-2 input + row statistics out, %g B per element" % out_bpe
                                                     if delta else "fp32 residual epilogue%s" % (" + bf16 stream copy + row statistics" if folded else "")),
                   "bound": "hbm", "achieved": ob / (oms * 1e-3) / 1e9 if oms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                   "traffic": traffic_of("gemm_out", olps), "algorithmic_bytes_per_launch": ob, "ms_per_launch": oms}
        out_hbm["frac"] = out_hbm["achieved"] / hbm_peak
        ln_bytes, ln_ms, ln_lps = per_launch("layernorm", 52 * 6.0 * rows_step * 1152)
        ln_ach = ln_bytes / (ln_ms * 1e-3) / 1e9 if ln_ms > 0 else 0.0
        launches_step = sum(prof_n.values()) / prof_steps
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": _workload_config(args, world),
            "e2e": {"value": e2e, "unit": "tokens/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": B * IMG * IMG * 3 + B * 128 + 4096,
                    "d2h_bytes_per_step": B * Lp * 8 + B * Lp + 4,
                    "how": "pinned-host uint8 images H2D on a copy stream every step; input_ids / attention_mask D2H for the "
                           "splice plan (enqueued before the encode kernels, waited for after them); a checksum of every "
                           "step's inputs_embeds slice D2H into pinned memory, consumed one step later"},
            "gpu_launches": int(round(launches_step * args.steps)),
            "clocks": clk,
            "roofline": fc2,
            "roofline_all_gemms": {"kernel": "all tcgen05 GEMM launches of a step (patch, qkv, out, fc1, fc2, projector)",
                                   "bound": "tensor", "achieved": achieved_all, "peak": peak, "unit": "TFLOP/s",
                                   "frac": achieved_all / peak, "traffic": None,
                                   "algorithmic_flops_per_step": gemm_flops_step, "kernel_ms_per_step": gemm_ms_step,
                                   "share_of_step": gemm_total_ms / total_ms},
            "roofline_gemm_qkv": tensor_roofline("gemm_qkv", "QKV GEMM (head-split epilogue)", 26 * 2.0 * rows_step * 1152 * 3456),
            "roofline_gemm_out": tensor_roofline("gemm_out", "out_proj GEMM (bf16-delta epilogue with the residual prefetch ring; fp32 residual epilogue with RADVLM_B200_OUTPROJ=f32)", 26 * 2.0 * rows_step * 1152 * 1152),
            "roofline_gemm_fc1": tensor_roofline("gemm_fc1", "fc1 GEMM (GELU-tanh epilogue)", 26 * 2.0 * rows_step * 1152 * 4304),
            "roofline_attention": tensor_roofline("attention", "siglip_attention_pp_kernel (MUFU issue co-limited, DESIGN.md)",
                                                  attn_flops * TILES_PER_IMAGE * B),
            "roofline_layernorm": ({"kernel": "layernorm_f32_to_bf16_kernel", "bound": "hbm", "achieved": ln_ach,
                                    "peak": hbm_peak, "unit": "GB/s", "frac": ln_ach / hbm_peak,
                                    "traffic": traffic_of("layernorm", ln_lps),
                                    "algorithmic_bytes_per_launch": ln_bytes, "ms_per_launch": ln_ms,
                                    "share_of_step": prof_ms.get("layernorm", 0.0) / total_ms}
                                   if prof_n.get("layernorm", 0) > 0 else
                                   {"kernel": "none: LayerNorm is folded into the QKV / fc1 GEMMs (row statistics and the "
                                              "bf16 copy of the stream come out of the residual GEMM epilogues); "
                                              "RADVLM_B200_LN=kernel restores the stand-alone kernels",
                                    "launches_per_step": 0.0, "share_of_step": 0.0}),
            # out_proj is the GEMM of a layer closest to the HBM bound (K = 1152: 155 GFLOP against 8 B per output
            # element, 12 B with RADVLM_B200_OUTPROJ=f32), so it also gets an HBM roofline
            "roofline_gemm_out_hbm": out_hbm,
            "kernel_ms_per_step": {k: v / prof_steps for k, v in prof_ms.items()},
            "kernel_launches_per_step": {k: v / prof_steps for k, v in prof_n.items()},
            "profiled_pass": {"steps": prof_steps, "ms_per_step": ms_prof / prof_steps,
                              "sum_of_kernels_ms_per_step": total_ms / prof_steps},
            "path_tflops": flops_per_tile() * TILES_PER_IMAGE * B * world / (ms / args.steps / 1e3) / 1e12,
            "path_frac_of_peak": flops_per_tile() * TILES_PER_IMAGE * B / (ms / args.steps / 1e3) / 1e12 / peak,
        }
        line["cuda_graph"] = graph_stats
        if gather_check is not None:
            line["gather_check"] = gather_check
        if c3 is not None:
            line["c3_strong"] = c3
        if b1 is not None:
            line["batch1"] = b1
        if world == 1 and not args.no_cpu_baseline:
            s_img, desc, kind, cores = _cpu_sample(False)
            line["cpu_baseline"] = {"value": TOKENS_PER_IMAGE / s_img, "unit": "tokens/s", "cores": cores, "kind": kind,
                                    "sample": desc}
            try:   # context only: the same architecture in torch eager bf16 on this GPU (tower + projector alone)
                v, sec = _gpu_eager_incumbent(host, dev)
                line["gpu_eager_incumbent"] = {
                    "value": v, "unit": "tokens/s",
                    "what": "torch eager bf16 (cuBLAS / ATen), the reference's op sequence for SigLIP tower + mm_projector "
                            "only (its CPU preprocessing and merge not charged), 80 tiles per call, %.1f ms" % (sec * 1e3)}
            except Exception as e:  # e.g. out of memory on a shared box: the headline numbers do not depend on it
                line["gpu_eager_incumbent"] = {"unavailable": str(e)[:200]}
    if world > 1 and peer[0] is not None:
        host.radvlm_b200_gather = None
        peer[0].close()   # unmap the peers' buffers, then free this rank's
        peer[0] = None
    # ---- training mode (BASELINE.json configs[4]) as a key of the same line
    if args.train_steps > 0:
        del dev_imgs, host_imgs
        torch.cuda.empty_cache()
        tr = _train_measure(args, host, world, rank, local, dev, steps=args.train_steps, warmup=3)
        if line is not None:
            line["train"] = tr
    if rank == 0:
        if args.workload == "c3" and line.get("c3_strong"):   # the strong-scaling configuration as the headline of this line
            c = line["c3_strong"]
            line["weak_scaling_value"] = line["value"]
            line.update({"value": c["value"], "scaling": "strong", "ms_per_step": c["ms_per_global_batch"],
                         "steps": c["repeats"], "e2e": None})
            line["config"]["workload"] = c["workload"] + " (a step = one global batch)"
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------------------------
# training mode (BASELINE.json configs[4]): a `train` key of the main line, or a line of its own with --mode train
# ----------------------------------------------------------------------------------------------------------------
def _train_measure(args, host, world, rank, local, dev, steps, warmup):
    """Step = preprocess + tower/projector/merge/splice forward (activations saved) + backward with a random upstream
    gradient on inputs_embeds + (N > 1) NCCL all-reduce (average) of the tower / projector gradients, issued per layer
    range from inside the backward, in place on the flat gradient buffer."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from radvlm_b200 import _lib, mm_arch, mm_utils
    from radvlm_b200.encoder import flops_per_tile
    import golden_inputs as gi

    B = args.train_batch
    host.model.vision_tower.requires_grad_(True)      # mm_tunable_parts = mm_vision_tower, mm_mlp_adapter
    host.model.mm_projector.requires_grad_(True)
    host.model.image_newline.requires_grad_(True)
    host.train()
    enc = mm_arch._encoder_for(host).freeze(False)
    enc.grad_allreduce_group = None if world > 1 else False
    rng = np.random.default_rng(2000 + rank)
    n_buf = 2
    dev_imgs = []
    for _ in range(n_buf):
        gray = rng.integers(0, 256, size=(B, IMG, IMG, 1), dtype=np.uint8)
        dev_imgs.append(torch.from_numpy(np.repeat(gray, 3, axis=3).copy()).to(dev))
    pr = _prompt(B, dev)
    upstream = [None]

    def step(images_u8, timed_fwd=None):
        tiles, sizes, splits, _ = mm_utils.preprocess_anyres_batch(list(images_u8), gi.PINPOINTS, device=dev,
                                                                   dtype=torch.bfloat16)
        out = host.prepare_inputs_labels_for_multimodal(pr[0], pr[1], pr[2], None, pr[3],
                                                        list(torch.split(tiles, splits)), ["image"] * B, sizes)
        emb = out[4]
        if upstream[0] is None:
            upstream[0] = torch.randn(emb.shape, device=dev, dtype=emb.dtype, generator=torch.Generator(device=dev).manual_seed(3))
        if timed_fwd is not None:
            timed_fwd.record()
        emb.backward(upstream[0])
        host.zero_grad(set_to_none=True)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(warmup):
        step(dev_imgs[i % n_buf])
    sync()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    tw0 = time.time()
    e0.record()
    for i in range(steps):
        starts[i].record()
        step(dev_imgs[i % n_buf], mids[i])
    e1.record()
    sync()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    fwd_ms = sum(s.elapsed_time(m) for s, m in zip(starts, mids)) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop(tw0, tw1) if clocks else None
    # per-class kernel times: separate profiled pass
    prof_steps = max(1, min(steps, 3))
    _lib.profile_enable(True)
    _lib.profile_read()
    for i in range(prof_steps):
        step(dev_imgs[i % n_buf])
    sync()
    prof_ms, prof_n = _lib.profile_read()
    _lib.profile_enable(False)
    host.model.vision_tower.requires_grad_(False)
    host.model.mm_projector.requires_grad_(False)
    host.model.image_newline.requires_grad_(False)
    host.eval()
    peaks, _ = _peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    tokens_step = B * world * TOKENS_PER_IMAGE
    step_ms = ms / steps
    # algorithmic FLOPs of a training step, nothing recomputed: forward + (dgrad + wgrad) of every Linear and the
    # 4 backward GEMMs of attention (dV, dP, dQ, dK) = 3 x forward, minus the dgrad the patch embedding does not need
    flops_step = (3.0 * flops_per_tile() - 2.0 * 729 * 588 * 1152) * TILES_PER_IMAGE * B
    return {
        "metric": "visual tokens/sec, training step (tower + projector + merge forward/backward)", "mode": "train",
        "value": tokens_step / (step_ms / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": step_ms, "forward_ms_per_step": fwd_ms,
        "backward_ms_per_step": step_ms - fwd_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[4]: %d synthetic 1024x1024 CXR per GPU per step (%d tiles), "
                               "SigLIP-so400m 26 layers + mlp2x_gelu + unpad/newline merge + splice, forward (activations "
                               "kept, nothing recomputed) + backward with a random upstream gradient on inputs_embeds"
                               % (B, B * TILES_PER_IMAGE)
                               + ("" if world == 1 else "; NCCL gradient all-reduce (average, in place on the flat fp32 "
                                                        "gradient buffer) per layer range, overlapped with the backward"),
                   "images_per_gpu_per_step": B, "parallelism": "dp%d (replicas)" % world},
        "clocks": clk,
        "kernel_ms_per_step": {k: round(v / prof_steps, 3) for k, v in prof_ms.items() if v > 0},
        "gpu_launches": int(sum(prof_n.values()) / prof_steps * steps),
        "algorithmic_flops_per_step": flops_step,
        "path_tflops": flops_step * world / (step_ms / 1e3) / 1e12,
        "path_frac_of_peak": flops_step / (step_ms / 1e3) / 1e12 / peak,
    }


def run_train_arm(args):
    import torch
    import torch.distributed as dist
    from radvlm_b200 import _lib, synthetic
    world, rank, local, dev = _init_dist()
    _lib.load()
    host = synthetic.build_host(hidden_size=3584, vocab=4096, seed=0, dtype=torch.bfloat16, device=dev)
    args.train_batch = args.train_batch if args.batch == 16 else args.batch   # --mode train --batch B keeps working
    tr = _train_measure(args, host, world, rank, local, dev, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        print(json.dumps(tr))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _ncu_traffic():
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/r02_traffic.json, the latest committed capture); {} if absent."""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    p = os.path.join(d, "r02_traffic.json")
    for older in ("r01b_traffic.json", "r01_traffic.json"):
        if not os.path.exists(p):
            p = os.path.join(d, older)
    try:
        with open(p) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "peer", "ce"],
                    help="N > 1: how the embeddings are all-gathered: ce = copy-engine push over peer memory (auto), "
                         "peer = fused merge + scatter kernel over peer memory, nccl = NCCL all-gather")
    ap.add_argument("--train-steps", type=int, default=10, help="timed training steps for the `train` key (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=4, help="images per GPU per training step")
    ap.add_argument("--no-c3", dest="c3", action="store_false", help="skip the strong-scaling (global batch 256) key")
    ap.add_argument("--no-batch1", dest="batch1", action="store_false", help="skip the batch-1 latency key")
    ap.add_argument("--workload", default="c1", choices=["c1", "c3"],
                    help="c1 (default): BASELINE configs[1] x batch per GPU, weak scaling; c3: configs[2], a fixed global "
                         "batch of 256 images (strong scaling) as the line's value")
    ap.add_argument("--mode", default="encode", choices=["encode", "train"],
                    help="encode: the headline metric (BASELINE configs[1..3]); train: configs[4], forward + backward")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.mode == "train":
        return run_train_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
