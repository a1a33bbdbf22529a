"""Generate golden vectors by running the UNMODIFIED reference (/root/reference).

Runs only in the build container (the GPU box has no /root/reference). The
reference is imported with the three shims of SURVEY.md section 8(c): stub
`timm`, stub `matplotlib`, identity `Tensor.cuda`. Inputs and weights come
from coponerf_b200.synth (numpy PCG64), so tests rebuild them from the seed
and only the reference OUTPUTS are stored here.

    python tests/golden/make_goldens.py            # writes tests/golden/*.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("COPONERF_REFERENCE", "/root/reference")


def import_reference():
    sys.path.insert(0, REF)
    timm = types.ModuleType("timm")
    tm = types.ModuleType("timm.models")
    tl = types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models = tm
    tm.layers = tl
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    mpl = types.ModuleType("matplotlib")
    mc = types.ModuleType("matplotlib.colors")
    mpl.colors = mc
    sys.modules.update({"matplotlib": mpl, "matplotlib.colors": mc})
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    from models import CoPoNeRF  # noqa: E402

    return CoPoNeRF


# (name, H, W, n_rays, S, pose_set, seed, val)
RENDER_CASES = [
    ("render_64_frontal", 64, 64, 512, 64, "frontal", 1, True),
    ("render_64_oblique", 64, 64, 512, 64, "oblique", 2, True),
    ("render_64_oblique_train", 64, 64, 256, 64, "oblique", 3, False),
    ("render_256_mild", 256, 256, 384, 64, "mild", 4, True),
    ("render_128_s128", 128, 128, 128, 128, "mild", 5, True),
]

# Float outputs stored per case. pixel_val is stored too (it is a CPU tensor in the reference).
FLOAT_KEYS = ["rgb", "valid_mask", "depth_ray", "at_wt", "pixel_val", "coords", "T_to_C1_pts",
              "T_to_C2_pts", "C2_pts_to_C1", "rel_pose_flip", "gt_rel_pose", "gt_rel_pose_flip"]
INT_KEYS = ["at_wt_max", "mask_c2", "matchability_cycle_mask"]


def run_render_case(model, case):
    from coponerf_b200 import synth

    name, H, W, n_rays, S, pose_set, seed, val = case
    inp = synth.make_input(H, W, n_rays, seed=seed, pose_set=pose_set)
    z, rel_pose, flow = synth.make_features(H, W, seed=seed)
    model.H, model.W, model.npoints = H, W, S
    with torch.no_grad():
        out = model(inp, z=z, rel_pose=rel_pose, flow=flow, val=val)
    rec = {}
    for k in FLOAT_KEYS:
        rec[k] = out[k].detach().cpu().numpy().astype(np.float32)
    for k in INT_KEYS:
        rec[k] = out[k].detach().cpu().numpy()
    rec["meta"] = np.array([H, W, n_rays, S, seed, int(val)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    valid = rec["valid_mask"].mean()
    print(f"{name}: rgb range [{rec['rgb'].min():.3f}, {rec['rgb'].max():.3f}] valid {valid:.3f} "
          f"at_wt max {rec['at_wt'].max():.3f}")


def main():
    from coponerf_b200 import synth

    CoPoNeRF = import_reference()
    torch.manual_seed(0)
    model = CoPoNeRF.CoPoNeRF(n_view=2).eval()
    missing, unexpected = model.load_state_dict(synth.render_state_dict(0), strict=False)
    assert not unexpected, unexpected
    for case in RENDER_CASES:
        run_render_case(model, case)


if __name__ == "__main__":
    main()
