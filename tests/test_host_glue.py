"""Host-side glue of ``prepare_inputs_labels_for_multimodal`` that needs no GPU: the zero-copy concatenation of per-image
tile blocks (``llava_arch.py:272``'s ``torch.cat``), the non-blocking ids / mask fetch, and the CUDA-graph capture policy
of the encoder (how captures are rationed)."""
import numpy as np
import torch

from radvlm_b200 import mm_arch


def test_concat_tiles_is_a_view_when_blocks_are_adjacent():
    t = torch.randn(25, 3, 8, 8)
    parts = list(t.split([10, 5, 10]))
    c = mm_arch._concat_tiles(parts)
    assert c.data_ptr() == t.data_ptr() and torch.equal(c, t)
    c = mm_arch._concat_tiles(parts[1:])                       # starts inside the allocation
    assert c.data_ptr() == t[10:].data_ptr() and torch.equal(c, t[10:])
    assert mm_arch._concat_tiles(parts[:1]) is parts[0]
    c = mm_arch._concat_tiles(list(t.split([10, 0, 15])))      # an image without tiles in the middle
    assert torch.equal(c, t)


def test_concat_tiles_copies_when_blocks_are_not_adjacent():
    t = torch.randn(25, 3, 8, 8)
    a, b, c = t.split([10, 5, 10])
    for parts in ([a, c], [b, a], [a.clone(), b], [a, b.clone()], [a[:, :, ::2], b[:, :, ::2]]):
        got = mm_arch._concat_tiles(parts)
        assert torch.equal(got, torch.cat(parts, 0))
        assert got.untyped_storage().data_ptr() != t.untyped_storage().data_ptr()   # a fresh tensor, not a view of t
    g = a.clone().requires_grad_(True)
    out = mm_arch._concat_tiles([g, b])
    assert out.requires_grad                                    # autograd inputs go through torch.cat


def test_ids_and_mask_fetch_on_host_tensors():
    ids = torch.tensor([[5, -200, 7], [1, 2, 3]], dtype=torch.int32)
    host, ev = mm_arch._to_host_async(ids, torch.int64)
    assert ev is None and host.dtype == torch.int64
    assert np.array_equal(mm_arch._host_result((host, ev)), ids.numpy().astype(np.int64))
    for mask in (torch.tensor([[1, 0, 2], [0, 0, 1]]), torch.tensor([[True, False, True], [False, False, True]]),
                 torch.tensor([[1.0, 0.0, 0.5], [0.0, 0.0, 3.0]])):
        m = mm_arch._host_result(mm_arch._to_host_async(mask, torch.uint8))
        assert m.dtype == np.uint8 and np.array_equal(m, np.array([[1, 0, 1], [0, 0, 1]], dtype=np.uint8))


def test_graph_capture_budget_policy():
    """encoder._launch_encode captures a key only after `graph_min_sightings` eager launches, and beyond the free captures
    only one per `graph_replays_per_capture` replays already served (host logic restated on the counters)."""
    from radvlm_b200.encoder import B200VisionEncoder
    from radvlm_b200 import synthetic
    vcfg = synthetic.siglip_config(hidden_size=32, intermediate_size=48, num_hidden_layers=1, num_attention_heads=2,
                                   image_size=28, patch_size=14)
    host = synthetic.build_host(hidden_size=16, vocab=8, seed=1, dtype=torch.float32, device="cpu", vision_cfg=vcfg)
    enc = B200VisionEncoder(host.model.vision_tower.vision_tower, host.model.mm_projector, num_heads=2, image_size=28)
    assert enc.graph_mode is False or enc.graph_mode is True    # opt-in through RADVLM_B200_GRAPH
    assert enc.graph_min_sightings >= 1 and enc.graph_free_captures >= 1 and enc.graph_replays_per_capture >= 16

    def allowed(captures, replays):
        return captures < enc.graph_free_captures + replays // enc.graph_replays_per_capture

    assert allowed(0, 0) and allowed(enc.graph_free_captures - 1, 0)
    assert not allowed(enc.graph_free_captures, enc.graph_replays_per_capture - 1)
    assert allowed(enc.graph_free_captures, enc.graph_replays_per_capture)
