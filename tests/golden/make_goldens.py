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


SENS_KEYS = ["rgb", "at_wt", "depth_ray", "T_to_C1_pts", "T_to_C2_pts", "C2_pts_to_C1"]
SENS_RUNS = 4


def sensitivity(model, inp, z, rel_pose, flow, val, base):
    """Per-ray conditioning of the reference itself: how far its fp32 outputs move when the camera poses are
    perturbed by about one ulp (relative 6e-8, seeded). Triangulating near-parallel rays amplifies such
    rounding-sized changes by orders of magnitude on a few rays (tests/cases.py uses this to widen the gate
    on exactly those rays, and nowhere else). Stored as <key>_sens = max over runs of max|delta| per ray."""
    import copy
    g = torch.Generator().manual_seed(1234)
    sens = {}
    for _ in range(SENS_RUNS):
        pin = copy.deepcopy(inp)
        for grp in ("context", "query"):
            t = pin[grp]["cam2world"]
            pin[grp]["cam2world"] = t * (1 + 6e-8 * torch.randn(t.shape, generator=g))
        rp = rel_pose * (1 + 6e-8 * torch.randn(rel_pose.shape, generator=g))
        with torch.no_grad():
            out = model(pin, z=z, rel_pose=rp, flow=flow, val=val)
        for k in SENS_KEYS:
            d = (out[k] - base[k]).abs().detach().cpu().numpy().astype(np.float32)
            if k == "rgb":
                d = d.max(axis=-1)[:, 0]          # (B, N)
            elif k == "at_wt":
                d = d.max(axis=-1)                # (2B, N)
            else:
                d = d.reshape(d.shape[0], d.shape[1], -1).max(axis=-1)   # (B, N)
            sens[k + "_sens"] = np.maximum(sens.get(k + "_sens", 0), d)
    return sens


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
    rec.update(sensitivity(model, inp, z, rel_pose, flow, val, out))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    valid = rec["valid_mask"].mean()
    rs = rec["rgb_sens"] / np.abs(rec["rgb"]).max()
    print(f"{name}: rgb range [{rec['rgb'].min():.3f}, {rec['rgb'].max():.3f}] valid {valid:.3f} "
          f"at_wt max {rec['at_wt'].max():.3f}; 1-ulp pose sensitivity of rgb: median {np.median(rs):.1e} "
          f"max {rs.max():.1e}, rays above 3e-6: {(rs > 3e-6).mean():.3f}")


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
