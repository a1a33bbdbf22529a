"""Device time of get_z() and of its three parts on one B200 (CUDA events, after warm-up).

    python scripts/bench_getz.py [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coponerf_b200 import pair_stage, pose_native, synth, ufc_native  # noqa: E402
from coponerf_b200.model import CoPoNeRF  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    m = CoPoNeRF(n_view=2).eval()
    m.load_state_dict(synth.full_state_dict(0), strict=True)
    m = m.to(dev)
    inp = synth.make_input(256, 256, None, seed=10)
    inp = {g: {k: v.to(dev) for k, v in d.items()} for g, d in inp.items()}
    for _ in range(3):
        z, rel, flow = m.get_z(inp)
    ops = m._ufc_ops

    def timeit(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.iters, out

    t_all, _ = timeit(lambda: m.get_z(inp))
    m.graph_get_z = True
    for _ in range(2):
        m.get_z(inp)
    t_graph, _ = timeit(lambda: m.get_z(inp))
    m.graph_get_z = False
    sd_ufc, sd_pose = m._sd_cache["ufc"], m._sd_cache["pose"]
    t_enc_unfolded, _ = timeit(lambda: pair_stage.encode_images(m, inp["context"]["rgb"]))
    t_enc, (pyr, zc) = timeit(lambda: pair_stage.encode_images(m, inp["context"]["rgb"], m._sd_cache["encoder"]))
    t_ufc, (feats, flows, c) = timeit(lambda: ufc_native.ufc_forward(sd_ufc, pyr, 2, ops))
    tokens = feats[-1].flatten(-2, -1).transpose(-1, -2)
    t_pose, _ = timeit(lambda: pose_native.pose_from_features(sd_pose, tokens, c, inp["context"]["intrinsics"], 256, ops))
    print(json.dumps({"get_z_graph_ms": t_graph, "get_z_ms": t_all, "encoder_ms": t_enc, "encoder_unfolded_bn_ms": t_enc_unfolded, "cost_aggregation_ms": t_ufc, "pose_ms": t_pose,
                      "iters": args.iters}))


if __name__ == "__main__":
    main()
