"""2+ GPU check of the fused merge + scatter (radvlm_merge_splice_scatter + dist.PeerGather) against a NCCL all-gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/peer_gather_check.py

Every rank builds a different batch (different features / ids), runs prepare_inputs_labels_for_multimodal twice per slot
through the peer path and compares the gathered [world, rows, H] buffer bit for bit with all_gather_into_tensor of the
plain path's inputs_embeds."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import golden_inputs as gi
    from radvlm_b200 import synthetic
    vcfg = synthetic.siglip_config(hidden_size=144, intermediate_size=272, num_hidden_layers=1, num_attention_heads=2)
    host = synthetic.build_host(hidden_size=64, vocab=gi.MERGE_VOCAB, seed=0, dtype=torch.bfloat16, device=dev, vision_cfg=vcfg)
    case = gi.merge_cases()["mixed"]
    ids, mask, labels = gi.merge_ids(case)
    pos = torch.arange(ids.shape[1])[None].expand(ids.shape[0], -1).contiguous()
    images = [torch.zeros(n, 3, 2, 2) for n in case["tiles"]]
    ok = True
    for mode in ("ce", "kernel"):
        ok = run_mode(mode, host, case, ids, mask, labels, pos, images, dev, rank, world) and ok
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER GATHER CHECK", "PASSED" if int(t.item()) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if int(t.item()) == 1 else 1


def run_mode(mode, host, case, ids, mask, labels, pos, images, dev, rank, world):
    """5 steps (both slots, reused twice -> the consumer-release barrier is exercised) with a different batch per rank
    and step; `mode`: "ce" = copy-engine push of the finished slice, "kernel" = the fused merge + scatter kernel."""
    from radvlm_b200.dist import PeerGather
    ok = True
    gather = None
    for it in range(5):
        g = torch.Generator().manual_seed(100 * it + rank)
        feats = torch.randn(sum(case["tiles"]), 729, 64, generator=g).to(dev, torch.bfloat16)
        host.encode_images = lambda images, _f=feats: _f
        args = (ids.to(dev), pos.to(dev), mask.to(dev), None, labels.to(dev), images)
        kw = dict(modalities=["image"] * ids.shape[0], image_sizes=case["sizes"])
        if hasattr(host, "radvlm_b200_gather"):
            del host.radvlm_b200_gather
        ref = host.prepare_inputs_labels_for_multimodal(*args, **kw)[4]
        want = torch.empty((world,) + tuple(ref.shape), dtype=ref.dtype, device=dev)
        dist.all_gather_into_tensor(want, ref.contiguous())
        if gather is None:
            gather = PeerGather(ref.shape[0] * ref.shape[1], ref.shape[2], ref.dtype, dev, mode=mode, timeout_s=60)
        host.radvlm_b200_gather = gather
        slot = gather.peek_slot()
        emb = host.prepare_inputs_labels_for_multimodal(*args, **kw)[4]
        got = gather.wait(slot)
        torch.cuda.synchronize(dev)
        same_local = torch.equal(emb, ref)
        same_all = torch.equal(got.view(world, *ref.shape), want)
        ok = ok and same_local and same_all
        print("rank %d mode %s iter %d slot %d: local %s gathered %s" % (rank, mode, it, slot, same_local, same_all), flush=True)
    host.radvlm_b200_gather = None
    gather.close()
    return ok


if __name__ == "__main__":
    sys.exit(main())
