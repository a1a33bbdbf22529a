"""Multi-GPU plumbing for the render path: one process per GPU, torch.distributed for the single exchange step.

Rays (and stereo pairs) are independent in CoPoNeRF.forward(), so the path shards with no
data-path collective; the only exchange is one gather of the final pixels. Two partitionings:
  * pairs  >= world: each rank renders whole pairs (bench.py --gpus N, BASELINE config 3);
  * one pair, many ranks: each rank renders a contiguous slice of the rays (`render_sharded`).
Every kernel reduces per ray in a fixed order, so a rank's slice equals the slice of a single-GPU render bit for bit.
"""
import torch
import torch.distributed as dist


def shard_range(n, world, rank):
    """Contiguous [lo, hi) of `n` items for `rank`: sizes differ by at most one, earlier ranks take the extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rays(local, n_total, dim, group=None):
    """All-gather a ray-sharded tensor (sharded along `dim` by shard_range) into the full tensor on every rank.

    One collective: slices are padded to the largest shard so a single all_gather_into_tensor moves them, and the
    padding is dropped afterwards.
    """
    world = dist.get_world_size(group)
    if world == 1:
        return local
    dim = dim % local.dim()
    biggest = -(-n_total // world)
    moved = local.movedim(dim, 0).contiguous()
    pad = biggest - moved.shape[0]
    if pad:
        moved = torch.cat((moved, moved.new_zeros((pad,) + tuple(moved.shape[1:]))), dim=0)
    out = moved.new_empty((world * biggest,) + tuple(moved.shape[1:]))
    dist.all_gather_into_tensor(out, moved, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        parts.append(out[r * biggest:r * biggest + (hi - lo)])
    return torch.cat(parts, dim=0).movedim(0, dim)


# dims along which the reference concatenates ray chunks (test.py:200-212)
RAY_DIM = {"rgb": -2, "valid_mask": -2, "depth_ray": -2, "at_wt": -2, "at_wt_max": -2, "pixel_val": -3, "coords": -2,
           "T_to_C1_pts": -2, "T_to_C2_pts": -2, "C2_pts_to_C1": -2, "mask_c2": -1, "matchability_cycle_mask": -1}


def render_sharded(forward, inp, keys=("rgb",), group=None, **kw):
    """Render one batch of pairs with the rays split over the ranks of `group`.

    `forward(inp, **kw)` is CoPoNeRF.forward (or anything with its contract). Every rank gets the gathered `keys`.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uv = inp["query"]["uv"]
    n = uv.shape[2]
    lo, hi = shard_range(n, world, rank)
    sub = {"context": inp["context"], "query": dict(inp["query"])}
    sub["query"]["uv"] = uv[:, :, lo:hi].contiguous()
    if "rgb" in sub["query"]:
        sub["query"]["rgb"] = inp["query"]["rgb"][:, :, lo:hi]
    out = forward(sub, **kw)
    full = dict(out)
    for k in keys:
        t = out[k]
        if t.dtype == torch.bool:
            full[k] = gather_rays(t.to(torch.uint8), n, RAY_DIM[k], group).bool()
        else:
            full[k] = gather_rays(t, n, RAY_DIM[k], group)
    full["uv"] = uv
    return full
