"""Oracle sensitivity to 1-ulp perturbations of the query/context ray geometry (CPU)."""
import os, sys, glob
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import make_case, rel_err, GOLDEN_DIR
from coponerf_b200 import synth
from oracle import render_oracle
torch.manual_seed(0)
orig = render_oracle.plucker_embedding
def noisy(c2w, uv, K):
    out = orig(c2w, uv, K)
    return out * (1 + 6e-8 * torch.randn_like(out))
for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "render_*.npz"))):
    g = dict(np.load(p)); H, W, n, S, seed, val = [int(v) for v in g["meta"]]
    inp, z, rel, flow = make_case(H, W, n, seed)
    sd = synth.render_state_dict(0)
    render_oracle.plucker_embedding = orig
    base = render_oracle.render_forward(sd, inp, z, rel, flow, H, W, S, bool(val))
    render_oracle.plucker_embedding = noisy
    pert = render_oracle.render_forward(sd, inp, z, rel, flow, H, W, S, bool(val))
    e = np.abs(pert["rgb"].numpy() - base["rgb"].numpy()).max(axis=-1)[0, 0]
    print(os.path.basename(p), {k: f"{rel_err(pert[k].numpy(), base[k].numpy()):.2e}" for k in ("rgb", "at_wt", "depth_ray", "T_to_C2_pts", "pixel_val")},
          "worst rays", np.argsort(e)[-3:], np.sort(e)[-3:])
