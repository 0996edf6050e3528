"""Container-only: the drop-in boundary against the REAL reference classes (skipped where /root/reference is absent).

``mm_arch.attach()`` is applied to ``Host(nn.Module, LlavaMetaForCausalLM)`` built from the unmodified reference
mixins (oracle/ref_loader.py: LlavaMetaModel builds the real SigLipVisionTower and the real mm_projector).  No GPU
here, so nothing is encoded: the test pins the parts of the contract that are host logic — every real parameter
name on the path is consumed by the weight packer (and only post_layernorm is left out), the merge table is built
from the real config / tower properties, the call without a CUDA device fails loudly instead of falling back."""
import pytest
import torch

import golden_inputs as gi
from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_host():
    from oracle.ref_loader import build_reference_host
    host, ref = build_reference_host(vocab=64, hidden_size=3584, seed=0)
    return host, ref


def test_attach_binds_the_reference_method_names(ref_host):
    from radvlm_b200 import mm_arch
    host, ref = ref_host
    cls = type(host)
    assert cls.encode_images is ref.llava_arch.LlavaMetaForCausalLM.encode_images      # unpatched class
    mm_arch.attach(host)
    assert host.encode_images.__func__ is mm_arch.encode_images
    assert host.prepare_inputs_labels_for_multimodal.__func__ is mm_arch.prepare_inputs_labels_for_multimodal
    assert cls.encode_images is ref.llava_arch.LlavaMetaForCausalLM.encode_images      # only the instance changed
    # the early-return contract needs no device (llava_arch.py:254-255)
    ids = torch.tensor([[5]])
    out = host.prepare_inputs_labels_for_multimodal(ids, None, None, "pkv", None, [torch.zeros(1, 3, 384, 384)], ["image"], [(384, 384)])
    assert out[0] is ids and out[3] == "pkv" and out[4] is None


def test_packing_covers_every_real_parameter_on_the_path(ref_host):
    from radvlm_b200 import mm_arch
    host, _ = ref_host
    mm_arch.attach(host)
    enc = mm_arch._encoder_for(host)
    tower = host.get_vision_tower()
    assert enc.tower_module is tower.vision_tower and enc.projector_module is host.get_model().mm_projector
    assert len(tower.vision_tower.vision_model.encoder.layers) == 26                   # load_model dropped the 27th
    pk = enc.packed("cpu")
    assert pk.tower.num_layers == 26 and pk.tower.hidden == 1152 and pk.tower.intermediate == 4304
    assert pk.tower.heads == 16 and pk.tower.patch_k_pad == 640 and pk.tokens_per_tile == 729
    assert pk.projector.in_dim == 1152 and pk.projector.hidden == 3584
    # every source Parameter that reached a packed buffer / alias, by storage address
    consumed = {s.data_ptr() for _, s in pk._copies} | {t.data_ptr() for t in pk._keep}
    def covered(p):
        # a source is consumed as a whole tensor or (patch weight / q,k,v) as a reshaped view with the same storage start
        return p.data_ptr() in consumed
    missing = [n for n, p in enc._source_named() if not covered(p)]
    assert sorted(missing) == ["vision_model.post_layernorm.bias", "vision_model.post_layernorm.weight"], missing
    # packed values == the reference parameters rounded to bf16 (spot check incl. the fused qkv layout)
    a = tower.vision_tower.vision_model.encoder.layers[7].self_attn
    qkv = next(t for t in pk._keep if t.data_ptr() == pk._layers[7].qkv_w)
    assert torch.equal(qkv, torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]).detach().bfloat16())
    # gradients: the names autograd will be handed back cover the same set
    names = [n for n, _ in enc._param_names()]
    assert len(names) == len(set(names)) == len(enc._source_tensors())


def test_merge_table_from_the_real_config_and_no_cpu_fallback(ref_host):
    from oracle import planner_oracle as po
    from radvlm_b200 import _lib, mm_arch
    host, _ = ref_host
    mm_arch.attach(host)
    sizes = list(gi.PARITY_SIZES)
    plans = [po.plan_image(s, gi.PINPOINTS) for s in sizes]
    table, tokens = mm_arch._merge_table(host, [p["n_tiles"] for p in plans], sizes, False)
    assert tokens == [p["n_tokens"] for p in plans]
    for m, p in zip(table, plans):
        assert (m.mode, m.grid_w, m.crop_r0, m.crop_c0, m.crop_h, m.crop_w, m.pool, m.out_h, m.out_w) == (
            _lib.MERGE_ANYRES, p["grid_w"], p["crop_r0"], p["crop_c0"], p["crop_h"], p["crop_w"], p["pool"], p["out_h"], p["out_w"])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA"):
            host.encode_images(torch.zeros(1, 3, 384, 384))
        ids = torch.tensor([[3, -200, 4]])
        with pytest.raises(RuntimeError):
            host.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, [torch.zeros(2, 3, 384, 384)], ["image"], [(384, 384)])
