"""Adds the per-ray conditioning of the whole call (BASELINE config 5) to tests/golden/pair_256.npz, measured on the
UNMODIFIED reference: how far forward(val=True)'s outputs move when its per-pair inputs (z, rel_pose, flow from get_z) are
perturbed by what two correct fp32 evaluations of get_z differ by -- features 2e-6 relative, rel_pose 4e-6 absolute,
flows 5e-5 px (the levels the CUDA get_z and the CPU restatement sit at against the reference) --
plus the one-ulp camera-pose perturbation of make_goldens.sensitivity. Stored as <key>_sens (max over 4 seeded runs of
max|delta| per ray), used by tests/cases.check_against exactly like the render goldens' sensitivities. The stored
reference outputs themselves are left untouched. Build container only.

    python tests/golden/add_pair_sens.py
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_goldens import SENS_KEYS, import_reference  # noqa: E402
from make_goldens_pair import CASE  # noqa: E402

RUNS = 4
Z_REL, POSE_ABS, FLOW_PX = 2e-6, 4e-6, 5e-5


def main():
    from coponerf_b200 import synth
    path = os.path.join(HERE, "pair_256.npz")
    rec = dict(np.load(path))
    ref = import_reference().CoPoNeRF(n_view=2).eval()
    ref.load_state_dict(synth.full_state_dict(CASE["weights_seed"]), strict=True)
    inp = synth.make_input(CASE["H"], CASE["W"], CASE["n_rays"], seed=CASE["seed"], pose_set=CASE["pose_set"])
    g = torch.Generator().manual_seed(4321)
    rn = lambda t: torch.randn(t.shape, generator=g)
    with torch.no_grad():
        z, rel_pose, flow = ref.get_z(inp)
        base = ref(inp, z=z, rel_pose=rel_pose, flow=flow, val=True)
        assert np.abs(base["rgb"].numpy() - rec["rgb"]).max() <= 1e-6 * np.abs(rec["rgb"]).max(), "reference run differs from the stored golden"
        sens = {}
        for _ in range(RUNS):
            pin = copy.deepcopy(inp)
            for grp in ("context", "query"):
                t = pin[grp]["cam2world"]
                pin[grp]["cam2world"] = t * (1 + 6e-8 * rn(t))
            zp = [t * (1 + Z_REL * rn(t)) for t in z]
            rp = rel_pose.clone()
            rp[:, :3, :] += POSE_ABS * rn(rp[:, :3, :])
            fp = tuple(f + (FLOW_PX if i < 2 else FLOW_PX / 32.0) * rn(f) for i, f in enumerate(flow))
            out = ref(pin, z=zp, rel_pose=rp, flow=fp, val=True)
            for k in SENS_KEYS:
                d = (out[k] - base[k]).abs().numpy().astype(np.float32)
                if k == "rgb":
                    d = d.max(axis=-1)[:, 0]
                elif k == "at_wt":
                    d = d.max(axis=-1)
                else:
                    d = d.reshape(d.shape[0], d.shape[1], -1).max(axis=-1)
                sens[k + "_sens"] = np.maximum(sens.get(k + "_sens", 0), d)
    rec.update(sens)
    np.savez_compressed(path, **rec)
    rs = sens["rgb_sens"] / np.abs(rec["rgb"]).max()
    print(f"pair_256: rgb sensitivity to get_z-level input noise: median {np.median(rs):.1e} p90 {np.quantile(rs, 0.9):.1e} "
          f"max {rs.max():.1e}; rays above 1e-5: {(rs > 1e-5).mean():.3f}")


if __name__ == "__main__":
    main()
