"""Host-side multi-rank logic on CPU (gloo, world_size 2 and 3): ray sharding + the single gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coponerf_b200.dist import gather_rays, render_sharded, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 512, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _fake_forward(inp, **kw):
    """Stands in for CoPoNeRF.forward: per-ray outputs that depend only on that ray's uv."""
    uv = inp["query"]["uv"]                      # (B, 1, n, 2)
    b, _, n, _ = uv.shape
    rgb = torch.stack((uv[..., 0] * 3 + 1, uv[..., 1] - 7, uv[..., 0] * uv[..., 1]), dim=-1)
    return {"rgb": rgb, "mask_c2": (uv[:, 0, :, 0] % 2 == 0), "at_wt": uv[:, 0, :, :1].repeat(2, 1, 4),
            "pixel_val": uv[:, 0, :, None, :].repeat(2, 1, 3, 1)}


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        uv = torch.randint(0, 256, (2, 1, n, 2), generator=g).float()
        inp = {"context": {}, "query": {"uv": uv}}
        keys = ("rgb", "mask_c2", "at_wt", "pixel_val")
        full = render_sharded(_fake_forward, inp, keys=keys)
        ref = _fake_forward(inp)
        for k in keys:
            assert full[k].shape == ref[k].shape and full[k].dtype == ref[k].dtype, k
            assert torch.equal(full[k], ref[k]), k
        lo, hi = shard_range(n, world, rank)
        assert torch.equal(gather_rays(uv[:, :, lo:hi], n, 2), uv)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n", [(2, 513), (3, 64), (2, 1)])
def test_ray_sharded_render_equals_single_process(world, n):
    mp.spawn(_worker, args=(world, _free_port(), n), nprocs=world, join=True)
