"""How much do the oracle outputs move under 1-ulp-sized perturbations of weights and features? (CPU)"""
import os, sys, glob
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import make_case, rel_err, GOLDEN_DIR
from coponerf_b200 import synth
from oracle import render_oracle
torch.manual_seed(0)
for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "render_*.npz"))):
    g = dict(np.load(p)); H, W, n, S, seed, val = [int(v) for v in g["meta"]]
    inp, z, rel, flow = make_case(H, W, n, seed)
    sd = synth.render_state_dict(0)
    base = render_oracle.render_forward(sd, inp, z, rel, flow, H, W, S, bool(val))
    sd2 = {k: v * (1 + 6e-8 * torch.randn_like(v)) for k, v in sd.items()}
    z2 = [t * (1 + 6e-8 * torch.randn_like(t)) for t in z]
    pert = render_oracle.render_forward(sd2, inp, z2, rel, flow, H, W, S, bool(val))
    print(os.path.basename(p), {k: f"{rel_err(pert[k].numpy(), base[k].numpy()):.2e}" for k in ("rgb", "at_wt", "depth_ray", "T_to_C2_pts")},
          "at_wt max", float(base["at_wt"].max()))
