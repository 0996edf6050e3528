"""GPU tests of the training-mode path (BASELINE.json config 5; run on a real B200: ``pytest -m gpu``).

Every call goes through the C ABI.  Checkers: plain PyTorch fp32 autograd of the same op for the building blocks,
and the committed gradients of the REAL reference (tests/golden/grad_golden.npz, produced by
tests/golden/make_golden.py grad: fp32 CPU autograd through process_anyres_image ->
prepare_inputs_labels_for_multimodal) for the end-to-end case.

Tolerance for gradients (bf16 tensor-core operands, fp32 accumulation and fp32 residual-stream gradient, against an
fp32 reference): cosine >= 0.999 and max|delta| / max|ref| <= 3e-2 per parameter tensor.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi

pytestmark = pytest.mark.gpu

GRAD_COS_MIN = 0.999
GRAD_RELMAX = 3e-2


def _metrics(got, ref):
    g, r = got.double().flatten().cpu(), ref.double().flatten().cpu()
    cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-300))
    relmax = float((g - r).abs().max() / (r.abs().max() + 1e-300))
    return cos, relmax


@pytest.fixture(scope="module")
def lib():
    from radvlm_b200 import _lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ============================================================================ building blocks
@pytest.mark.parametrize("M,N,K", [(512, 256, 128), (1458, 1152, 4304), (300, 200, 136)])
def test_gemm_dgrad_form_vs_torch(lib, M, N, K):
    """dX[M, in] = dY[M, out] W[out, in]: W is read as stored (MN-major B operand)."""
    from radvlm_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(1)
    dY = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(K, N, device="cuda", generator=g) * 0.1).bfloat16()     # [out = K, in = N]
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    _lib.check(lib.radvlm_gemm_bf16_ex(dY.data_ptr(), K, 0, W.data_ptr(), N, 1, M, N, K, None, _lib.EPI_BIAS_F32,
                                       out.data_ptr(), N, None, 0, 1, _stream()))
    torch.testing.assert_close(out, dY.float() @ W.float(), rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize("rows,out_dim,in_dim,splits", [(1458, 1152, 4304, 1), (1458, 4304, 1152, 3), (1000, 200, 304, 4)])
def test_gemm_wgrad_form_vs_torch(lib, rows, out_dim, in_dim, splits):
    """dW[out, in] += dY^T X: both operands MN-major, split-K partial sums added atomically to a running sum."""
    from radvlm_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(2)
    dY = torch.randn(rows, out_dim, device="cuda", generator=g).bfloat16()
    X = torch.randn(rows, in_dim, device="cuda", generator=g).bfloat16()
    acc = torch.full((out_dim, in_dim), 0.5, device="cuda", dtype=torch.float32)
    _lib.check(lib.radvlm_gemm_bf16_ex(dY.data_ptr(), out_dim, 1, X.data_ptr(), in_dim, 1, out_dim, in_dim, rows, None,
                                       _lib.EPI_ATOMIC_F32, acc.data_ptr(), in_dim, None, 0, splits, _stream()))
    torch.testing.assert_close(acc, 0.5 + dY.float().t() @ X.float(), rtol=1e-3, atol=5e-2)


def test_colsum_gelu_layernorm_backward_vs_torch(lib):
    from radvlm_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(3)
    rows, D, I = 1458, 1152, 4304
    # bias gradient
    dy = torch.randn(rows, I, device="cuda", generator=g).bfloat16()
    db = torch.full((I,), 2.0, device="cuda")
    _lib.check(lib.radvlm_colsum_bf16(dy.data_ptr(), rows, I, I, db.data_ptr(), _stream()))
    torch.testing.assert_close(db, 2.0 + dy.float().sum(0), rtol=1e-4, atol=1e-2)
    # GELU forward + backward, both forms
    for erf_form, approx in ((0, "tanh"), (1, "none")):
        u = (torch.randn(rows, I, device="cuda", generator=g) * 2).bfloat16()
        da = torch.randn(rows, I, device="cuda", generator=g).bfloat16()
        uf = u.float().requires_grad_(True)
        a_ref = torch.nn.functional.gelu(uf, approximate=approx)
        a_ref.backward(da.float())
        du = da.clone()
        a = torch.empty_like(u)
        _lib.check(lib.radvlm_gelu_fwd_bwd_bf16(u.data_ptr(), du.data_ptr(), a.data_ptr(), u.numel(), erf_form, _stream()))
        torch.testing.assert_close(a.float(), a_ref.detach(), rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(du.float(), uf.grad, rtol=1e-2, atol=1e-2)
    # LayerNorm backward (dx added to the running residual gradient)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 1
    gamma = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    beta = 0.1 * torch.randn(D, device="cuda", generator=g)
    dyn = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    dres0 = torch.randn(rows, D, device="cuda", generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-6).backward(dyn.float())
    dres = dres0.clone()
    dgamma, dbeta = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    stats = torch.empty(rows, 2, device="cuda")
    _lib.check(lib.radvlm_layernorm_bwd(x.data_ptr(), gamma.data_ptr(), dyn.data_ptr(), dres.data_ptr(), dgamma.data_ptr(),
                                        dbeta.data_ptr(), stats.data_ptr(), rows, D, 1e-6, _stream()))
    torch.testing.assert_close(dres, dres0 + xr.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dgamma, gr.grad, rtol=1e-3, atol=2e-2)
    torch.testing.assert_close(dbeta, br.grad, rtol=1e-3, atol=2e-2)


def test_attention_backward_vs_torch_autograd(lib):
    from radvlm_b200 import _lib
    tiles, heads, hd, T, Tp, hp = 2, 16, 72, 729, 768, 80
    D = heads * hd
    g = torch.Generator(device="cuda").manual_seed(6)
    q = torch.zeros(tiles, heads, Tp, hp, device="cuda", dtype=torch.bfloat16)
    k = torch.zeros_like(q)
    vt = torch.zeros_like(q)                                     # V, same layout as q / k
    q[:, :, :T, :hd] = torch.randn(tiles, heads, T, hd, device="cuda", generator=g)
    k[:, :, :T, :hd] = torch.randn(tiles, heads, T, hd, device="cuda", generator=g)
    vt[:, :, :T, :hd] = torch.randn(tiles, heads, T, hd, device="cuda", generator=g)
    dout = torch.randn(tiles * T, D, device="cuda", generator=g).bfloat16()
    vt_fwd = vt.clone()
    vt_fwd[:, :, :T, hd] = 1   # the forward sums P with a ones column; the backward wants plain zeros there
    out = torch.empty(tiles * T, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(tiles * heads, Tp, device="cuda", dtype=torch.float32)
    scale = hd ** -0.5
    _lib.check(lib.radvlm_attention_fwd_lse(q.data_ptr(), k.data_ptr(), vt_fwd.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                            tiles, heads, T, Tp, hd, hp, scale, _stream()))
    qf = q[:, :, :T, :hd].float().requires_grad_(True)
    kf = k[:, :, :T, :hd].float().requires_grad_(True)
    vf = vt[:, :, :T, :hd].float().contiguous().requires_grad_(True)
    s = qf @ kf.transpose(-1, -2) * scale
    ref = (torch.softmax(s, dim=-1) @ vf).transpose(1, 2).reshape(tiles * T, D)
    torch.testing.assert_close(lse.view(tiles, heads, Tp)[:, :, :T], torch.logsumexp(s, dim=-1).detach() * 1.4426950408889634,
                               rtol=1e-3, atol=1e-2)
    ref.backward(dout.float())
    wsb = lib.radvlm_attention_bwd_workspace_bytes(tiles, heads, Tp)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    dqkv = torch.empty(tiles * T, 3 * D, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.radvlm_attention_bwd(q.data_ptr(), k.data_ptr(), vt.data_ptr(), dout.data_ptr(), out.data_ptr(),
                                        lse.data_ptr(), dqkv.data_ptr(), ws.data_ptr(), wsb, tiles, heads, T, Tp, hd, hp,
                                        scale, _stream()))
    got = dqkv.float().view(tiles, T, 3, heads, hd).permute(2, 0, 3, 1, 4)    # [3, tiles, heads, T, hd]
    for name, gg, rr in (("dQ", got[0], qf.grad), ("dK", got[1], kf.grad), ("dV", got[2], vf.grad)):
        cos, relmax = _metrics(gg, rr)
        assert cos >= GRAD_COS_MIN and relmax <= GRAD_RELMAX, "%s: cos=%.6f relmax=%.3e" % (name, cos, relmax)


# ============================================================================ end to end vs the real reference
def _small_train_host():
    from radvlm_b200 import synthetic
    v = dict(gi.SMALL_VISION)
    v["num_hidden_layers"] -= 1
    host = synthetic.build_host(hidden_size=gi.SMALL_PROJ, vocab=64, seed=gi.GRAD_SEED, dtype=torch.float32, device="cuda",
                                vision_cfg=synthetic.siglip_config(**v))
    host.requires_grad_(True)
    host.train()
    return host


def _grad_batch(dtype):
    from radvlm_b200 import mm_utils
    imgs = [torch.from_numpy(gi.grad_image(n)) for n in gi.GRAD_IMAGES]
    tiles, sizes, splits, _ = mm_utils.preprocess_anyres_batch(imgs, gi.PINPOINTS, dtype=dtype)
    L = max(len(r) for r in gi.GRAD_IDS)
    ids = torch.zeros(len(gi.GRAD_IDS), L, dtype=torch.long)
    mask = torch.zeros(len(gi.GRAD_IDS), L, dtype=torch.bool)
    for b, r in enumerate(gi.GRAD_IDS):
        ids[b, :len(r)] = torch.tensor(r)
        mask[b, :len(r)] = True
    labels = torch.where(ids < 0, torch.full_like(ids, -100), ids)
    pos = torch.arange(L)[None].expand(len(gi.GRAD_IDS), -1).contiguous()
    return list(torch.split(tiles, splits)), sizes, ids.cuda(), pos.cuda(), mask.cuda(), labels.cuda()


def test_training_path_gradients_vs_reference_golden(lib, golden_dir):
    """Tower + projector + merge + splice under autograd against the REAL reference's fp32 gradients
    (2x1-grid image: crop without pooling; 4x4-grid image: bilinear pooling; newline and text rows)."""
    z = np.load(os.path.join(golden_dir, "grad_golden.npz"))
    host = _small_train_host()
    images, sizes, ids, pos, mask, labels = _grad_batch(torch.float32)
    assert [int(t.shape[0]) for t in images] == [int(v) for v in z["tile_counts"]]
    out = host.prepare_inputs_labels_for_multimodal(ids, pos, mask, None, labels, images, ["image"] * len(images), sizes)
    emb = out[4]
    assert tuple(emb.shape) == tuple(int(v) for v in z["embeds_shape"]) and emb.requires_grad
    cos, relmax = _metrics(emb.detach()[:, torch.from_numpy(z["rows"]).cuda()], torch.from_numpy(z["embeds_rows"]))
    assert cos >= 0.999 and relmax <= 2e-2, "forward (train mode): cos=%.6f relmax=%.3e" % (cos, relmax)
    R = gi.grad_loss_weights(emb.shape).cuda()
    (emb * R).sum().backward()
    worst = (1.0, 0.0, "")
    checked = 0
    scale_ref = max(float(np.abs(z[k]).max()) for k in z.files if k.startswith("grad/"))
    for name, p in host.named_parameters():
        key = "grad/" + name
        if key not in z.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, "%s: gradient for a parameter off the path" % name
            continue
        ref = torch.from_numpy(z[key])
        assert p.grad is not None, "%s: no gradient" % name
        assert p.grad.shape == ref.shape
        if float(ref.abs().max()) < 1e-5 * scale_ref:     # analytically zero (k_proj.bias: softmax shift invariance)
            assert float(p.grad.abs().max()) <= 1e-3 * scale_ref, name
            continue
        cos, relmax = _metrics(p.grad, ref)
        assert cos >= GRAD_COS_MIN and relmax <= GRAD_RELMAX, "%s: cos=%.6f relmax=%.3e" % (name, cos, relmax)
        if cos < worst[0]:
            worst = (cos, relmax, name)
        checked += 1
    assert checked >= 54
    print("gradient parity vs reference: %d tensors, worst cos=%.6f (relmax %.3e) at %s" % ((checked,) + worst))


def test_training_path_accumulates_and_respects_frozen_parts(lib):
    """Two backward passes accumulate into .grad like autograd does; frozen parameters get no gradient
    (mm_tunable_parts = projector only must not run the tower backward)."""
    host = _small_train_host()
    images, sizes, ids, pos, mask, labels = _grad_batch(torch.float32)

    def run():
        emb = host.prepare_inputs_labels_for_multimodal(ids, pos, mask, None, labels, images, ["image"] * len(images), sizes)[4]
        (emb * gi.grad_loss_weights(emb.shape).cuda()).sum().backward()

    run()
    g1 = {n: p.grad.clone() for n, p in host.named_parameters() if p.grad is not None}
    run()
    scale = max(float(v.abs().max()) for v in g1.values())
    for n, p in host.named_parameters():
        if n in g1 and float(g1[n].abs().max()) > 1e-4 * scale:   # (k_proj.bias is analytically zero: rounding noise)
            cos, relmax = _metrics(p.grad, 2 * g1[n])
            assert cos > 0.9999 and relmax < 1e-2, "%s: accumulation cos=%.6f relmax=%.3e" % (n, cos, relmax)
    host.zero_grad(set_to_none=True)
    host.model.vision_tower.requires_grad_(False)
    host.model.embed_tokens.requires_grad_(False)
    run()
    for n, p in host.named_parameters():
        if "vision_tower" in n or "embed_tokens" in n:
            assert p.grad is None, n
    for n in ("model.mm_projector.0.weight", "model.mm_projector.2.bias", "model.image_newline"):
        cos, relmax = _metrics(dict(host.named_parameters())[n].grad, g1[n])
        assert cos > 0.9999 and relmax < 1e-2, "%s with a frozen tower: cos=%.6f relmax=%.3e" % (n, cos, relmax)


def test_backward_full_size_vs_oracle_autograd(lib):
    """BASELINE.json configs[4] at the real widths (SigLIP-so400m: 1152 / 4304 / 16 heads x 72, 26 layers; projector
    1152 -> 3584 -> 3584): gradients of two tiles against fp32 CPU autograd of the oracle on the same (bf16) weights."""
    from oracle import encoder_oracle as eo
    from radvlm_b200 import synthetic
    host = synthetic.build_host(hidden_size=3584, vocab=64, seed=gi.FULL_SEED, dtype=torch.bfloat16, device="cuda")
    host.model.vision_tower.requires_grad_(True)
    host.model.mm_projector.requires_grad_(True)
    host.train()
    x = gi.encoder_pixels(2, seed=31).bfloat16()
    feat = host.encode_images(x.cuda())
    assert feat.requires_grad and tuple(feat.shape) == (2, 729, 3584)
    R = (torch.randn(feat.shape, generator=torch.Generator().manual_seed(8)) * 0.1).bfloat16()
    feat.backward(R.cuda())
    torch.cuda.synchronize()
    tower = dict(host.model.vision_tower.vision_tower.named_parameters())
    proj = dict(host.model.mm_projector.named_parameters())
    watch_t = ["vision_model.encoder.layers.25.mlp.fc2.weight", "vision_model.encoder.layers.25.self_attn.out_proj.bias",
               "vision_model.encoder.layers.12.layer_norm1.weight", "vision_model.encoder.layers.12.mlp.fc1.weight",
               "vision_model.encoder.layers.0.self_attn.q_proj.weight", "vision_model.encoder.layers.0.self_attn.v_proj.weight",
               "vision_model.embeddings.patch_embedding.weight", "vision_model.embeddings.position_embedding.weight"]
    watch_p = ["0.weight", "2.weight", "2.bias"]
    tsd = {k: v.detach().float().cpu() for k, v in host.model.vision_tower.vision_tower.state_dict().items()}
    psd = {k: v.detach().float().cpu() for k, v in host.model.mm_projector.state_dict().items()}
    for k in watch_t:
        tsd[k].requires_grad_(True)
    for k in watch_p:
        psd[k].requires_grad_(True)
    ref = eo.encode_images(tsd, psd, x.float())
    ref.backward(R.float())
    worst = (1.0, "")
    for name, got, want in [(k, tower[k].grad, tsd[k].grad) for k in watch_t] + [(k, proj[k].grad, psd[k].grad) for k in watch_p]:
        assert got is not None and torch.isfinite(got).all(), name
        cos, relmax = _metrics(got.float(), want)
        # bf16 .grad tensors (the Parameters are bf16) of a 26-layer bf16-operand backward
        assert cos >= 0.999 and relmax <= 5e-2, "%s: cos=%.6f relmax=%.3e" % (name, cos, relmax)
        if cos < worst[0]:
            worst = (cos, name)
    print("full-size gradient parity vs oracle autograd: worst cos=%.6f at %s" % worst)


@pytest.mark.parametrize("padding_side", ["right", "left"])
def test_splice_backward_with_truncation_vs_oracle(lib, padding_side):
    """tokenizer_model_max_length shorter than a sample (llava_arch.py:495-498): the text tokens that were cut off get
    ZERO gradient in embed_tokens (their rows of the compact d_text buffer are written by no segment), the surviving
    ones, the visual features and image_newline get the oracle's fp32 autograd gradient."""
    from oracle import encoder_oracle as eo
    from radvlm_b200 import synthetic
    vcfg = synthetic.siglip_config(hidden_size=144, intermediate_size=272, num_hidden_layers=1, num_attention_heads=2)
    host = synthetic.build_host(hidden_size=gi.MERGE_HIDDEN, vocab=gi.MERGE_VOCAB, seed=0, dtype=torch.float32,
                                device="cuda", vision_cfg=vcfg)
    with torch.no_grad():
        host.model.embed_tokens.weight.copy_(gi.merge_embed_table())
        host.model.image_newline.copy_(gi.merge_newline())
    host.model.embed_tokens.requires_grad_(True)
    host.model.image_newline.requires_grad_(True)
    host.config.tokenizer_padding_side = padding_side
    sizes, tiles = [(500, 300), (1024, 1024)], [3, 10]
    g = torch.Generator().manual_seed(77)
    feats = torch.randn(sum(tiles), 729, gi.MERGE_HIDDEN, generator=g)
    lengths, L = [30, 50], 50
    ids = torch.zeros(2, L, dtype=torch.long)
    mask = torch.zeros(2, L, dtype=torch.bool)
    for b, n in enumerate(lengths):
        ids[b, :n] = torch.randint(1, gi.MERGE_VOCAB, (n,), generator=g)
        mask[b, :n] = True
    ids[0, 4] = -200      # sample 0: 4 text + image + 25 text; the cut falls inside the trailing text
    ids[1, 45] = -200     # sample 1: 45 text + image + 4 text; the cut falls inside the image, all trailing text is lost
    labels = torch.where(ids < 0, torch.full_like(ids, -100), ids)
    # oracle side (fp32 CPU autograd)
    f_ref = feats.clone().requires_grad_(True)
    nl_ref = gi.merge_newline().clone().requires_grad_(True)
    tab_ref = gi.merge_embed_table().clone().requires_grad_(True)
    per_image, base = [], 0
    for n, size in zip(tiles, sizes):
        per_image.append(eo.merge_image(f_ref[base:base + n], size, nl_ref, gi.PINPOINTS))
        base += n
    n0 = int(per_image[0].shape[0])
    max_length = 4 + n0 + 10           # sample 0 keeps 10 of its 25 trailing text tokens
    assert 45 + int(per_image[1].shape[0]) > max_length > 45 + 100
    host.config.tokenizer_model_max_length = max_length
    r_emb, _, r_mask, _ = eo.prepare_inputs_labels(tab_ref, per_image, ids, mask, labels, max_length, padding_side == "left")
    R = torch.randn(r_emb.shape, generator=g)
    (r_emb * R).sum().backward()
    # product side
    f_dev = feats.cuda().requires_grad_(True)
    host.encode_images = lambda images: f_dev
    images = [torch.zeros(n, 3, 2, 2) for n in tiles]
    out = host.prepare_inputs_labels_for_multimodal(ids.cuda(), None, mask.cuda(), None, labels.cuda(), images,
                                                    ["image", "image"], sizes)
    emb = out[4]
    assert tuple(emb.shape) == tuple(r_emb.shape) and emb.shape[1] == max_length
    assert torch.equal(emb.detach().cpu(), r_emb.detach())
    (emb * R.cuda()).sum().backward()
    torch.cuda.synchronize()
    got = host.model.embed_tokens.weight.grad.cpu()
    assert torch.isfinite(got).all()
    torch.testing.assert_close(got, tab_ref.grad, rtol=1e-5, atol=1e-5)
    # tokens that only occur behind the cut have exactly zero gradient
    kept = set(ids[0, :4].tolist()) | set(ids[0, 5:15].tolist()) | set(ids[1, :45].tolist())
    lost = (set(ids[0, 15:30].tolist()) | set(ids[1, 46:50].tolist())) - kept
    assert lost and all(float(got[t].abs().max()) == 0.0 for t in lost)
    torch.testing.assert_close(f_dev.grad.cpu(), f_ref.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(host.model.image_newline.grad.cpu(), nl_ref.grad, rtol=1e-4, atol=1e-4)
