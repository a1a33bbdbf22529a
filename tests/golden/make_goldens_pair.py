"""Golden vectors of the per-pair stage and of the whole pipeline, from the UNMODIFIED reference (/root/reference):
get_z() (image encoder, cost aggregation, CrossBlock pose features, pose head) followed by forward(val=True) on a
256x256 pair -- BASELINE config 5 ("joint pose + correspondence + render forward pass").

Build container only. Inputs and all 744 weights come from coponerf_b200.synth (numpy PCG64), so tests rebuild them
from the seed and only the reference OUTPUTS are stored.

    python tests/golden/make_goldens_pair.py            # writes tests/golden/pair_256.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_goldens import FLOAT_KEYS, INT_KEYS, import_reference  # noqa: E402

CASE = dict(H=256, W=256, n_rays=256, seed=7, pose_set="mild", weights_seed=0)
# strided samples of the big tensors: z levels (2, C, n, n) and the averaged correlation c (1, 1, 64, 64, 64, 64)
Z_STRIDE = ((4, 2, 2), (4, 4, 4), (4, 8, 8), (4, 16, 16))
C_STRIDE = 8


def main():
    from coponerf_b200 import synth
    ref = import_reference().CoPoNeRF(n_view=2).eval()
    ref.load_state_dict(synth.full_state_dict(CASE["weights_seed"]), strict=True)
    inp = synth.make_input(CASE["H"], CASE["W"], CASE["n_rays"], seed=CASE["seed"], pose_set=CASE["pose_set"])
    with torch.no_grad():
        z, rel_pose, flow = ref.get_z(inp)
        out = ref(inp, z=z, rel_pose=rel_pose, flow=flow, val=True)
        # the module's third UFC output, for the pose-feature operators
        c = ref.feature_cost_aggregation(ref.encoder.forward(_norm(inp), None, 2)[:3], 2)[2]
    rec = {"rel_pose": rel_pose.numpy()}
    for i, f in enumerate(flow):
        rec[f"flow{i}"] = f.numpy()
    for i, (t, (sc, sy, sx)) in enumerate(zip(z, Z_STRIDE)):
        rec[f"z{i}"] = t[:, ::sc, ::sy, ::sx].contiguous().numpy()
    rec["c"] = c[:, :, ::C_STRIDE, ::C_STRIDE, ::C_STRIDE, ::C_STRIDE].contiguous().numpy()
    for k in FLOAT_KEYS:
        rec[k] = out[k].detach().cpu().numpy().astype(np.float32)
    for k in INT_KEYS:
        rec[k] = out[k].detach().cpu().numpy()
    np.savez_compressed(os.path.join(HERE, "pair_256.npz"), **rec)
    print("pair_256: rel_pose\n", rel_pose[0].numpy(), "\nvalid", float(out["valid_mask"].mean()),
          "rgb range", float(out["rgb"].min()), float(out["rgb"].max()))


def _norm(inp):
    x = torch.flatten(inp["context"]["rgb"], 0, 1).permute(0, -1, 1, 2)
    x = ((x + 1) / 2.).clone()
    for ch, (m, s) in enumerate(((0.485, 0.229), (0.456, 0.224), (0.406, 0.225))):
        x[:, ch] = (x[:, ch] - m) / s
    return x


if __name__ == "__main__":
    main()
