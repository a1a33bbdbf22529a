"""Two-GPU check (NCCL): a ray-sharded render of one pair, gathered with one collective, equals the single-GPU render
bit for bit. Skipped when fewer than two devices are visible."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cases import cuda_model, make_case, to_device
    from coponerf_b200.dist import render_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        H = W = 64
        inp, z, rel, flow = make_case(H, W, 777, seed=2)
        from coponerf_b200.model import CoPoNeRF
        from coponerf_b200 import synth
        m = CoPoNeRF(n_view=2, npoints=64, chunk_rays=128)
        m.load_state_dict(synth.render_state_dict(0), strict=False)
        m = m.to(dev).eval()
        m.H, m.W = H, W
        kw = dict(z=to_device(z, dev), rel_pose=rel.to(dev), flow=to_device(flow, dev), val=True)
        inp_d = to_device(inp, dev)
        keys = ("rgb", "at_wt", "at_wt_max", "mask_c2", "depth_ray")
        full = render_sharded(m, inp_d, keys=keys, **kw)
        ref = m(inp_d, **kw)
        for k in keys:
            assert torch.equal(full[k], ref[k]), k
    finally:
        dist.destroy_process_group()


def test_two_gpu_ray_sharding_bit_exact():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
