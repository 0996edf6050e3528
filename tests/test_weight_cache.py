"""Host logic of the packed-weight cache (radvlm_b200.encoder.PackedWeights / B200VisionEncoder.packed) on CPU.

The reference trains under DeepSpeed (scripts/zero3.json, finetune_radio_7b.sh:62), whose optimizers update Parameters
through ``p.data`` — which does NOT bump autograd's ``_version`` counter.  The packed bf16 copies must follow anyway."""
import pytest
import torch

from radvlm_b200 import synthetic
from radvlm_b200.encoder import B200VisionEncoder


def _host(dtype):
    vcfg = synthetic.siglip_config(hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=2,
                                   image_size=28, patch_size=14)
    return synthetic.build_host(hidden_size=16, vocab=8, seed=1, dtype=dtype, device="cpu", vision_cfg=vcfg)


def _enc(host):
    return B200VisionEncoder(host.model.vision_tower.vision_tower, host.model.mm_projector, num_heads=2, image_size=28)


def _packed_tensor(pk, ptr):
    return next(t for t in pk._keep if t.data_ptr() == ptr)


def test_packed_layout_and_padding():
    host = _host(torch.float32)
    pk = _enc(host).packed("cpu")
    vm = host.model.vision_tower.vision_tower.vision_model
    pw = _packed_tensor(pk, pk.tower.patch_w)
    assert pw.shape == (32, 640) and pw.dtype == torch.bfloat16                      # 3*14*14 = 588 -> 640
    assert torch.equal(pw[:, :588], vm.embeddings.patch_embedding.weight.detach().reshape(32, 588).bfloat16())
    assert float(pw[:, 588:].abs().max()) == 0.0
    a = vm.encoder.layers[1].self_attn
    qkv = _packed_tensor(pk, pk._layers[1].qkv_w)
    assert torch.equal(qkv, torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]).detach().bfloat16())
    qkv_b = _packed_tensor(pk, pk._layers[1].qkv_b)
    assert qkv_b.dtype == torch.float32 and torch.equal(qkv_b, torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]).detach())
    assert pk.tower.num_layers == 2 and pk.tower.patch_k_pad == 640 and pk.projector.hidden == 16


def test_data_updates_are_seen_while_training():
    """DeepSpeed-style ``p.data.copy_`` / ``p.data.add_`` (no version bump) with trainable parameters."""
    host = _host(torch.float32)
    host.model.vision_tower.requires_grad_(True)
    enc = _enc(host)
    pk = enc.packed("cpu")
    fc1 = host.model.vision_tower.vision_tower.vision_model.encoder.layers[0].mlp.fc1.weight
    ptr = pk._layers[0].fc1_w
    v0 = fc1._version
    fc1.data.add_(1.0)
    fc1.data.copy_(fc1.data * 0.5)
    assert fc1._version == v0                       # the update is invisible to the version counter
    pk2 = enc.packed("cpu")
    assert pk2 is pk and pk2._layers[0].fc1_w == ptr and enc.n_repacks == 1      # refreshed in place, same pointers
    assert torch.equal(_packed_tensor(pk2, ptr), fc1.detach().bfloat16())


def test_frozen_weights_version_bump_and_invalidate():
    host = _host(torch.float32)                     # nothing requires grad: pure inference
    enc = _enc(host)
    pk = enc.packed("cpu")
    w2 = host.model.mm_projector[2].weight
    ptr = pk.projector.w2
    n0 = enc.n_refreshes
    enc.packed("cpu")
    assert enc.n_refreshes == n0                    # unchanged weights: no copy at all
    with torch.no_grad():
        w2.mul_(2.0)                                # torch.optim-style in-place update: version bump -> refresh
    enc.packed("cpu")
    assert enc.n_refreshes == n0 + 1 and torch.equal(_packed_tensor(pk, ptr), w2.detach().bfloat16())
    w2.data.mul_(3.0)                               # silent update at inference: needs invalidate()
    enc.packed("cpu")
    assert not torch.equal(_packed_tensor(pk, ptr), w2.detach().bfloat16())
    enc.invalidate().packed("cpu")
    assert torch.equal(_packed_tensor(pk, ptr), w2.detach().bfloat16())


def test_bf16_parameters_are_aliased_and_storage_swap_rebuilds():
    host = _host(torch.bfloat16)
    enc = _enc(host)
    pk = enc.packed("cpu")
    fc2 = host.model.vision_tower.vision_tower.vision_model.encoder.layers[1].mlp.fc2.weight
    assert pk.n_alias > 0 and pk._layers[1].fc2_w == fc2.data_ptr()      # zero-copy: the struct points at the Parameter
    fc2.data = (fc2.data * 2).clone()               # ZeRO gather / checkpoint load: new storage -> rebuild
    pk2 = enc.packed("cpu")
    assert pk2 is not pk and enc.n_repacks == 2 and pk2._layers[1].fc2_w == fc2.data_ptr()


def test_partitioned_parameters_raise():
    host = _host(torch.float32)
    enc = _enc(host)
    p = host.model.mm_projector[0].weight
    p.data = torch.empty(0)                         # what ZeRO-3 leaves outside of its gather context
    with pytest.raises(RuntimeError, match="partitioned"):
        enc.packed("cpu")
