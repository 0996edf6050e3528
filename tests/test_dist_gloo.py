"""Host-side logic of the data-parallel encode (radvlm_b200.dist) on CPU: world_size-2 gloo processes.
Covers the LPT sharding and the all-gather-v reassembly; the per-rank encode is faked by a deterministic
function of the image index (the GPU kernels are covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radvlm_b200 import dist as rdist


def test_lpt_sharding_balanced_and_deterministic():
    tiles = [10, 2, 37, 13, 13, 31, 3, 10, 17, 5]
    for world in (1, 2, 4, 8):
        owned = rdist.shard_images_lpt(tiles, world)
        assert sorted(i for o in owned for i in o) == list(range(len(tiles)))   # a partition
        loads = [sum(tiles[i] for i in o) for o in owned]
        assert max(loads) - min(loads) <= max(tiles)                              # LPT bound
        assert owned == rdist.shard_images_lpt(tiles, world)
    assert rdist.shard_images_lpt([10] * 8, 8) == [[i] for i in range(8)]
    assert rdist.shard_images_lpt([], 2) == [[], []]


def _fake_tokens(i, n, H):
    g = torch.Generator().manual_seed(100 + i)
    return torch.randn(n, H, generator=g)


def _worker(rank, world, port, tiles, counts, H, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def encode_and_merge(indices):
            if not indices:
                return torch.zeros(0, H)
            return torch.cat([_fake_tokens(i, counts[i], H) for i in indices], dim=0)

        out = rdist.encode_images_sharded([None] * len(tiles), [None] * len(tiles), tiles, counts, encode_and_merge)
        ok = all(torch.equal(out[i], _fake_tokens(i, counts[i], H)) for i in range(len(tiles)))
        ret[rank] = bool(ok and len(out) == len(tiles))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tiles", [[10, 2, 13, 31, 3], [10, 10], [2]])
def test_all_gather_v_world2_gloo(tiles):
    counts = [729 + 7 * t for t in tiles]   # ragged token counts
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, tiles, counts, 16, ret), nprocs=2, join=True)
    assert ret.get(0) is True and ret.get(1) is True


def test_single_process_passthrough():
    counts = [5, 9, 3]
    toks = torch.cat([_fake_tokens(i, n, 8) for i, n in enumerate(counts)])
    out = rdist.gather_visual_tokens(toks, [[0, 1, 2]], counts)
    assert all(torch.equal(out[i], _fake_tokens(i, counts[i], 8)) for i in range(3))


# ---------------------------------------------------------------------------------------------- training mode
def test_gradient_bucket_plan():
    assert rdist.plan_gradient_buckets([], 100) == []
    assert rdist.plan_gradient_buckets([10, 20, 30], 100) == [[0, 1, 2]]
    assert rdist.plan_gradient_buckets([60, 50, 50, 500, 10], 100) == [[0], [1, 2], [3], [4]]   # an oversized tensor stands alone


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = [(7, 5), (3,), (64, 33), (1,), (129,)]
        mk = lambda r: [torch.full(s, float(r + 1)) + torch.arange(int(torch.tensor(s).prod())).reshape(s) * 0.01 for s in shapes]
        mine = mk(rank)
        want = [sum(mk(r)[i] for r in range(world)) / world for i in range(len(shapes))]
        red = rdist.GradientAllReducer(bucket_bytes=256 * 4, average=True)   # several buckets, one oversized tensor
        red.submit(mine[:2])       # two submissions, as the layer-range backward does
        red.submit(mine[2:])
        red.finish()
        ok = all(torch.allclose(a, b, rtol=0, atol=1e-6) for a, b in zip(mine, want))
        mine2 = mk(rank)
        rdist.allreduce_gradients(mine2, average=False)
        ok = ok and all(torch.allclose(a, b * world, rtol=0, atol=1e-5) for a, b in zip(mine2, want))
        # in-place form: a contiguous slice of one flat buffer, cut into buckets, averaged without any copy
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        ptr = flat.data_ptr()
        red = rdist.GradientAllReducer(bucket_bytes=96 * 4, average=True)
        red.submit_flat(flat[100:700])
        red.submit_flat(flat[700:1000])
        red.finish()
        mean = sum(r + 1 for r in range(world)) / world
        want_flat = torch.arange(1000, dtype=torch.float32)
        want_flat = torch.cat([want_flat[:100] * (rank + 1), want_flat[100:] * mean])
        ok = ok and flat.data_ptr() == ptr and torch.allclose(flat, want_flat, rtol=1e-6, atol=1e-4)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) is True and ret.get(1) is True
