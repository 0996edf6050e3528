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
`roofline`: the dominant kernel class (tcgen05 GEMM) timed live with CUDA events on the launch stream.
`cpu_baseline`: the oracle port (or the real reference when /root/reference exists) on the host cores, bounded sample.
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
                   ("fused merge + scatter over peer memory (radvlm_merge_splice_scatter)" if getattr(args, "gather", "nccl") in ("peer", "auto")
                    else "NCCL all-gather of inputs_embeds, asynchronous")),
        "l2": "per-step working set (0.83 GB bf16 weights + >1 GB activations) exceeds the 126 MB L2; input images rotate over 3 buffers",
    }


# ----------------------------------------------------------------------------------------------------------------
# CPU arms (oracle port / real reference)
# ----------------------------------------------------------------------------------------------------------------
def _cpu_sample(use_reference: bool):
    """One bounded sample of the CPU path: preprocess 1 image, tower+projector on 1 tile, merge+splice of 1 image.
    Returns (seconds per image extrapolated to 10 tiles, description, kind, cores)."""
    import numpy as np
    import torch
    import golden_inputs as gi
    cores = torch.get_num_threads()
    rng = np.random.default_rng(0)
    gray = rng.integers(0, 256, size=(IMG, IMG), dtype=np.uint8)
    if use_reference:
        from PIL import Image
        st = _cpu_state(True)
        host, ref = st["host"], st["ref"]
        pil = Image.fromarray(np.repeat(gray[:, :, None], 3, axis=2))
        t0 = time.perf_counter()
        tiles = ref.mm_utils.process_anyres_image(pil, host.get_vision_tower().image_processor, gi.PINPOINTS)
        t_pre = time.perf_counter() - t0
        t0 = time.perf_counter()
        with torch.no_grad():
            feat1 = host.encode_images(tiles[:1])
        t_tile = time.perf_counter() - t0
        feats = feat1.expand(TILES_PER_IMAGE, -1, -1).contiguous()
        host.encode_images = lambda images, _f=feats: _f
        ids = torch.randint(1, 1000, (1, 32))
        ids[0, 5] = -200
        t0 = time.perf_counter()
        with torch.no_grad():
            host.prepare_inputs_labels_for_multimodal(ids, None, torch.ones_like(ids, dtype=torch.bool), None, ids.clone(),
                                                      [tiles], ["image"], [(IMG, IMG)])
        t_merge = time.perf_counter() - t0
        del host.encode_images
        kind = "reference"
    else:
        from oracle import encoder_oracle as eo
        from oracle import resample_oracle as ro
        st = _cpu_state(False)
        t0 = time.perf_counter()
        tiles = torch.from_numpy(ro.process_anyres_image(gray, gi.PINPOINTS))
        t_pre = time.perf_counter() - t0
        t0 = time.perf_counter()
        with torch.no_grad():
            feat1 = eo.encode_images(st["tsd"], st["psd"], tiles[:1])
        t_tile = time.perf_counter() - t0
        feats = feat1.expand(TILES_PER_IMAGE, -1, -1).contiguous()
        ids = torch.randint(1, 1000, (1, 32))
        ids[0, 5] = -200
        t0 = time.perf_counter()
        merged = eo.merge_image(feats, (IMG, IMG), st["newline"], gi.PINPOINTS)
        eo.prepare_inputs_labels(st["table"], [merged], ids, torch.ones_like(ids, dtype=torch.bool), ids.clone(), 32768, False)
        t_merge = time.perf_counter() - t0
        kind = "port"
    per_image = t_pre + TILES_PER_IMAGE * t_tile + t_merge
    desc = ("1 image: process_anyres_image (%.2fs) + tower/projector fp32 on 1 of 10 tiles (%.2fs, x10 extrapolated) + "
            "merge/splice (%.2fs)" % (t_pre, t_tile, t_merge))
    return per_image, desc, kind, cores


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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.ref_loader import reference_available
    use_ref = reference_available()
    for _ in range(args.warmup):
        _cpu_sample(use_ref)
    t_total, per_image_s, desc, kind, cores = 0.0, [], "", "", 1
    for _ in range(args.steps):
        t0 = time.perf_counter()
        s, desc, kind, cores = _cpu_sample(use_ref)
        t_total += time.perf_counter() - t0
        per_image_s.append(s)
    per_image = sorted(per_image_s)[len(per_image_s) // 2]
    value = TOKENS_PER_IMAGE / per_image
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _workload_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": kind,
                         "sample": "each step = " + desc + "; value = 7371 tokens / median extrapolated seconds per image"},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
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


def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from radvlm_b200 import _lib, mm_utils, synthetic
    from radvlm_b200.encoder import flops_per_tile
    import golden_inputs as gi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B = args.batch
    host = synthetic.build_host(hidden_size=3584, vocab=4096, seed=0, dtype=torch.bfloat16, device=dev)
    from radvlm_b200 import mm_arch
    mm_arch._encoder_for(host).freeze()

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
    ids = torch.randint(1, 4000, (B, Lp), generator=torch.Generator().manual_seed(5))
    ids[:, 7] = -200
    ids_dev = ids.to(dev)
    mask_dev = torch.ones_like(ids_dev, dtype=torch.bool)
    labels_dev = torch.where(ids_dev < 0, torch.full_like(ids_dev, -100), ids_dev)
    pos_dev = torch.arange(Lp, device=dev)[None].expand(B, -1).contiguous()
    # N > 1: the all-gather of step i runs asynchronously (NCCL's own stream) and overlaps the encode of step i+1;
    # two (embeddings, gathered) slots alternate, a slot is reused only after its collective has completed.
    # --gather peer: no collective call at all - the merge kernel itself writes every embedding row into the other
    # ranks' gathered buffers over NVLink (radvlm_merge_splice_scatter, dist.PeerGather), on a side stream.
    slots = [{"work": None, "emb": None, "out": None} for _ in range(2)]
    step_no = [0]
    peer = [None]

    def drain_gathers():
        if peer[0] is not None:
            peer[0].drain()
        for sl in slots:
            if sl["work"] is not None:
                sl["work"].wait()
                sl["work"] = None

    def step(images_u8):
        tiles, sizes, splits, _ = mm_utils.preprocess_anyres_batch(list(images_u8), gi.PINPOINTS, device=dev,
                                                                   dtype=torch.bfloat16)
        out = host.prepare_inputs_labels_for_multimodal(ids_dev, pos_dev, mask_dev, None, labels_dev,
                                                        list(torch.split(tiles, splits)), ["image"] * B, sizes)
        emb = out[4]
        if world > 1 and args.gather in ("peer", "auto"):
            if peer[0] is None:   # first step: size the peer buffers from the embeddings, then redo the step into them
                from radvlm_b200.dist import PeerGather
                try:
                    peer[0] = PeerGather(emb.shape[0] * emb.shape[1], emb.shape[2], emb.dtype, dev)
                except Exception as e:  # no cudaIpc / peer access on this box: the NCCL all-gather does the same job
                    if args.gather == "peer":
                        raise
                    print("bench: peer-memory gather unavailable (%s); using the NCCL all-gather" % e, file=sys.stderr)
                    args.gather = "nccl"
                    return step(images_u8)
                args.gather = "peer"
                host.radvlm_b200_gather = peer[0]
                return step(images_u8)
        elif world > 1:
            sl = slots[step_no[0] & 1]
            step_no[0] += 1
            if sl["work"] is not None:
                sl["work"].wait()
            if sl["out"] is None:
                sl["out"] = torch.empty((world,) + tuple(emb.shape), dtype=emb.dtype, device=dev)
            sl["emb"] = emb
            sl["work"] = dist.all_gather_into_tensor(sl["out"], emb, async_op=True)
        return emb

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(src_list, steps, read_back):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        chk = 0.0
        for i in range(steps):
            emb = step(src_list[i % n_buf])
            if read_back:
                chk += float(emb[0, -1, :8].float().sum().item())   # D2H read of a slice of the result
        drain_gathers()   # the timed region ends when the last all-gather has landed
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t_wall0, time.time(), chk

    for i in range(args.warmup):
        step(dev_imgs[i % n_buf])
    drain_gathers()
    sync()

    # ---- device-resident arm, with live per-kernel-class timing (CUDA events on the launch stream)
    _lib.profile_enable(True)
    _lib.profile_read()
    clocks = ClockSampler(local) if rank == 0 else None
    ms, tw0, tw1, _ = timed(dev_imgs, args.steps, read_back=False)
    clk = clocks.stop(tw0, tw1) if clocks else None
    prof_ms, prof_n = _lib.profile_read()
    _lib.profile_enable(False)

    # ---- end-to-end arm: pinned host uint8 in, slice of the result out, every step
    for i in range(min(2, args.warmup)):
        step(host_imgs[i % n_buf])
    ms_e2e, _, _, _ = timed(host_imgs, args.steps, read_back=True)

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
        gemm_ms_step = gemm_total_ms / args.steps
        achieved_all = gemm_flops_step / (gemm_ms_step * 1e-3) / 1e12 if gemm_ms_step > 0 else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        hbm_peak = float(peaks.get("hbm_gbs", 6545.0))
        total_ms = sum(prof_ms.values()) or 1.0
        rows_step = B * TILES_PER_IMAGE * 729
        traffic = _ncu_traffic()

        def per_launch(cls, work_step):
            n = max(prof_n.get(cls, 0), 1)
            return work_step * args.steps / n, prof_ms.get(cls, 0.0) / n, n / args.steps

        def traffic_of(cls, launches_per_step):
            """DRAM bytes of one launch from the ncu capture, if it was taken at this run's tiles-per-launch."""
            t = traffic.get(cls)
            tower_calls = launches_per_step / (52.0 if cls == "layernorm" else 26.0)
            if not t or tower_calls <= 0:
                return None
            if abs(B * TILES_PER_IMAGE / tower_calls - traffic.get("tiles_per_launch", -1)) > 1e-6:
                return None
            return (t["dram_read_mb"] + t["dram_write_mb"]) * 1e6

        fc2_flops, fc2_ms, fc2_lps = per_launch("gemm_fc2", 26 * 2.0 * rows_step * 1152 * 4304)
        fc2_ach = fc2_flops / (fc2_ms * 1e-3) / 1e12 if fc2_ms > 0 else 0.0
        ln_bytes, ln_ms, ln_lps = per_launch("layernorm", 52 * 6.0 * rows_step * 1152)
        ln_ach = ln_bytes / (ln_ms * 1e-3) / 1e9 if ln_ms > 0 else 0.0
        at_flops, at_ms, at_lps = per_launch("attention", attn_flops * TILES_PER_IMAGE * B)
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": _workload_config(args, world),
            "e2e": {"value": e2e, "unit": "tokens/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": B * IMG * IMG * 3 + B * 128 + 4096,
                    "d2h_bytes_per_step": B * Lp * 8 + 4},
            "gpu_launches": int(sum(prof_n.values())),
            "clocks": clk,
            "roofline": {"kernel": "gemm_bf16_tn_2cta_kernel<192,3> as the fc2 GEMM (tcgen05 cta_group::2, fp32 residual "
                                   "epilogue)", "bound": "tensor",
                         "achieved": fc2_ach, "peak": peak, "unit": "TFLOP/s", "frac": fc2_ach / peak,
                         "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peak_src,
                         "traffic": traffic_of("gemm_fc2", fc2_lps),
                         "algorithmic_flops_per_launch": fc2_flops, "ms_per_launch": fc2_ms,
                         "launches_per_step": fc2_lps,
                         "share_of_step": prof_ms.get("gemm_fc2", 0.0) / total_ms},
            "roofline_all_gemms": {"kernel": "all tcgen05 GEMM launches of a step (patch, qkv, out, fc1, fc2, projector)",
                                   "bound": "tensor", "achieved": achieved_all, "peak": peak, "unit": "TFLOP/s",
                                   "frac": achieved_all / peak, "traffic": None,
                                   "algorithmic_flops_per_step": gemm_flops_step, "kernel_ms_per_step": gemm_ms_step,
                                   "share_of_step": gemm_total_ms / total_ms},
            "roofline_attention": {"kernel": "siglip_attention_pp_kernel", "bound": "tensor (MUFU issue co-limited, DESIGN.md)",
                                   "achieved": at_flops / (at_ms * 1e-3) / 1e12 if at_ms > 0 else 0.0, "peak": peak,
                                   "unit": "TFLOP/s", "frac": (at_flops / (at_ms * 1e-3) / 1e12 / peak) if at_ms > 0 else 0.0,
                                   "traffic": traffic_of("attention", at_lps),
                                   "share_of_step": prof_ms.get("attention", 0.0) / total_ms},
            "roofline_layernorm": {"kernel": "layernorm_f32_to_bf16_kernel", "bound": "hbm", "achieved": ln_ach,
                                   "peak": hbm_peak, "unit": "GB/s", "frac": ln_ach / hbm_peak,
                                   "traffic": traffic_of("layernorm", ln_lps),
                                   "algorithmic_bytes_per_launch": ln_bytes, "ms_per_launch": ln_ms,
                                   "share_of_step": prof_ms.get("layernorm", 0.0) / total_ms},
            "kernel_ms_per_step": {k: v / args.steps for k, v in prof_ms.items()},
            "kernel_launches_per_step": {k: v / args.steps for k, v in prof_n.items()},
            "path_tflops": flops_per_tile() * TILES_PER_IMAGE * B * world / (ms / args.steps / 1e3) / 1e12,
            "path_frac_of_peak": flops_per_tile() * TILES_PER_IMAGE * B / (ms / args.steps / 1e3) / 1e12 / peak,
        }
        if world == 1 and not args.no_cpu_baseline:
            s, desc, kind, cores = _cpu_sample(False)
            line["cpu_baseline"] = {"value": TOKENS_PER_IMAGE / s, "unit": "tokens/s", "cores": cores, "kind": kind,
                                    "sample": desc}
            try:   # context only: the same architecture in torch eager bf16 on this GPU (tower + projector alone)
                v, sec = _gpu_eager_incumbent(host, dev)
                line["gpu_eager_incumbent"] = {
                    "value": v, "unit": "tokens/s",
                    "what": "torch eager bf16 (cuBLAS / ATen), the reference's op sequence for SigLIP tower + mm_projector "
                            "only (its CPU preprocessing and merge not charged), 80 tiles per call, %.1f ms" % (sec * 1e3)}
            except Exception as e:  # e.g. out of memory on a shared box: the headline numbers do not depend on it
                line["gpu_eager_incumbent"] = {"unavailable": str(e)[:200]}
        print(json.dumps(line))
    if world > 1:
        if peer[0] is not None:
            host.radvlm_b200_gather = None
            peer[0].close()   # unmap the peers' buffers, then free this rank's
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------------------------
# training-mode arm (BASELINE.json configs[4]): --mode train.  Not the driver's headline metric: an extra line.
# ----------------------------------------------------------------------------------------------------------------
def run_train_arm(args):
    """Step = preprocess + tower/projector/merge/splice forward (activations saved) + backward with a random upstream
    gradient on inputs_embeds + (N > 1) bucketed NCCL all-reduce of the tower / projector gradients, overlapped."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from radvlm_b200 import _lib, mm_arch, mm_utils, synthetic
    from radvlm_b200.encoder import flops_per_tile
    import golden_inputs as gi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    B = args.batch
    host = synthetic.build_host(hidden_size=3584, vocab=4096, seed=0, dtype=torch.bfloat16, device=dev)
    host.model.vision_tower.requires_grad_(True)      # mm_tunable_parts = mm_vision_tower, mm_mlp_adapter
    host.model.mm_projector.requires_grad_(True)
    host.model.image_newline.requires_grad_(True)
    host.train()
    enc = mm_arch._encoder_for(host)
    enc.grad_allreduce_group = None if world > 1 else False
    rng = np.random.default_rng(2000 + rank)
    n_buf = 2
    dev_imgs = []
    for _ in range(n_buf):
        gray = rng.integers(0, 256, size=(B, IMG, IMG, 1), dtype=np.uint8)
        dev_imgs.append(torch.from_numpy(np.repeat(gray, 3, axis=3).copy()).to(dev))
    Lp = 32
    ids = torch.randint(1, 4000, (B, Lp), generator=torch.Generator().manual_seed(5))
    ids[:, 7] = -200
    ids_dev = ids.to(dev)
    mask_dev = torch.ones_like(ids_dev, dtype=torch.bool)
    labels_dev = torch.where(ids_dev < 0, torch.full_like(ids_dev, -100), ids_dev)
    pos_dev = torch.arange(Lp, device=dev)[None].expand(B, -1).contiguous()
    upstream = [None]
    t_fwd = [0.0]

    def step(images_u8, timed_fwd=None):
        tiles, sizes, splits, _ = mm_utils.preprocess_anyres_batch(list(images_u8), gi.PINPOINTS, device=dev,
                                                                   dtype=torch.bfloat16)
        out = host.prepare_inputs_labels_for_multimodal(ids_dev, pos_dev, mask_dev, None, labels_dev,
                                                        list(torch.split(tiles, splits)), ["image"] * B, sizes)
        emb = out[4]
        if upstream[0] is None:
            upstream[0] = torch.randn(emb.shape, device=dev, dtype=emb.dtype, generator=torch.Generator(device=dev).manual_seed(3))
        if timed_fwd is not None:
            timed_fwd.record()
        emb.backward(upstream[0])
        host.zero_grad(set_to_none=True)

    for i in range(args.warmup):
        step(dev_imgs[i % n_buf])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    _lib.profile_enable(True)
    _lib.profile_read()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    tw0 = time.time()
    e0.record()
    for i in range(args.steps):
        starts[i].record()
        step(dev_imgs[i % n_buf], mids[i])
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    fwd_ms = sum(s.elapsed_time(m) for s, m in zip(starts, mids)) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop(tw0, tw1) if clocks else None
    prof_ms, prof_n = _lib.profile_read()
    _lib.profile_enable(False)
    if rank == 0:
        peaks, _ = _peaks()
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        tokens_step = B * world * TOKENS_PER_IMAGE
        step_ms = ms / args.steps
        # algorithmic FLOPs of a training step with per-layer recompute: forward + recomputed
        # QKV / out_proj / fc1 + 2x every Linear + 2.5x attention for the tower, 3x for the projector
        hidden, inter, seq, proj, L = 1152, 4304, 729, 3584, 26
        lin = 2.0 * seq * (4 * hidden * hidden + 2 * hidden * inter)
        attn = 4.0 * seq * seq * hidden
        tower_fwd = L * (lin + attn) + 2.0 * seq * 588 * hidden
        recompute = L * 2.0 * seq * (4 * hidden * hidden + hidden * inter)
        bwd = L * (2 * lin + 2.5 * attn) + 2.0 * seq * 588 * hidden
        proj_f = 2.0 * seq * (hidden * proj + proj * proj)
        flops_step = (tower_fwd + recompute + bwd + proj_f * (1 + 0.35 + 2)) * TILES_PER_IMAGE * B
        print(json.dumps({
            "metric": "visual tokens/sec, training step (tower + projector + merge forward/backward)", "mode": "train",
            "value": tokens_step / (step_ms / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "forward_ms_per_step": fwd_ms,
            "backward_ms_per_step": step_ms - fwd_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[4]: %d synthetic 1024x1024 CXR per GPU per step, SigLIP-so400m 26 "
                                   "layers + mlp2x_gelu + unpad/newline merge + splice, forward + backward (layer "
                                   "recompute) with a random upstream gradient on inputs_embeds" % B
                                   + ("" if world == 1 else "; bucketed NCCL gradient all-reduce overlapped with the backward"),
                       "images_per_gpu_per_step": B, "parallelism": "dp%d (replicas)" % world},
            "clocks": clk,
            "kernel_ms_per_step": {k: round(v / args.steps, 3) for k, v in prof_ms.items() if v > 0},
            "gpu_launches": int(sum(prof_n.values())),
            "path_tflops": flops_step * world / (step_ms / 1e3) / 1e12,
            "path_frac_of_peak": flops_step / (step_ms / 1e3) / 1e12 / peak,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _ncu_traffic():
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/r01b_traffic.json, the latest committed capture); {} if absent."""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    p = os.path.join(d, "r01b_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(d, "r01_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "peer"],
                    help="N > 1: how the embeddings are all-gathered (NCCL all-gather, or the fused merge + scatter "
                         "kernel over peer memory)")
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
