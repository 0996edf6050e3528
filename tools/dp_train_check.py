"""2-GPU check of the data-parallel training path (run under torchrun on a GPU box):
every rank runs forward + backward on ITS images with the in-backward NCCL gradient all-reduce switched on, then the
averaged gradients are compared with a single-process computation of both shards on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_inputs as gi  # noqa: E402
from radvlm_b200 import mm_arch, mm_utils, synthetic  # noqa: E402


def grads_for(host, images_u8, dev, allreduce):
    enc = mm_arch._encoder_for(host)
    enc.grad_allreduce_group = None if allreduce else False
    host.zero_grad(set_to_none=True)
    tiles, sizes, splits, _ = mm_utils.preprocess_anyres_batch(images_u8, gi.PINPOINTS, device=dev, dtype=torch.bfloat16)
    feat = host.encode_images(tiles)
    R = torch.randn(feat.shape, generator=torch.Generator().manual_seed(7)).to(dev).to(feat.dtype)
    (feat.float() * R.float()).sum().backward()
    tower = host.model.vision_tower.vision_tower
    return {n: p.grad.float().clone() for n, p in list(tower.named_parameters()) + list(host.model.mm_projector.named_parameters())
            if p.grad is not None}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    v = dict(gi.SMALL_VISION)
    v["num_hidden_layers"] -= 1
    host = synthetic.build_host(hidden_size=gi.SMALL_PROJ, vocab=64, seed=3, dtype=torch.bfloat16, device=dev,
                                vision_cfg=synthetic.siglip_config(**v))
    host.model.vision_tower.requires_grad_(True)
    host.model.mm_projector.requires_grad_(True)
    host.train()
    rng = np.random.default_rng(5)
    imgs = [torch.from_numpy(rng.integers(0, 256, size=(500, 300 + 50 * i, 3), dtype=np.uint8)) for i in range(world)]
    reduced = grads_for(host, [imgs[rank]], dev, allreduce=True)
    ok = True
    if rank == 0:
        per_rank = [grads_for(host, [imgs[r]], dev, allreduce=False) for r in range(world)]
        gmax = max(float(sum(pr[n] for pr in per_rank).abs().max()) / world for n in reduced)
        for n, g in reduced.items():
            want = sum(pr[n] for pr in per_rank) / world
            # k_proj.bias gradients are analytically zero (softmax shift invariance): judge them on the global scale
            scale = max(float(want.abs().max()), 1e-3 * gmax)
            err = float((g - want).abs().max()) / scale
            if err > 2e-2:
                ok = False
                print("MISMATCH", n, err)
        print("dp_train_check world=%d: %d gradient tensors averaged over ranks %s" % (world, len(reduced), "OK" if ok else "FAILED"))
    else:
        for r in range(world):   # keep the collective-free local work symmetric (rank 0 recomputes all shards alone)
            pass
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
