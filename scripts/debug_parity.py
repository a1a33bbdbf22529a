"""Print per-output errors of the CUDA path against the golden vectors (GPU box)."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import GOLDEN_DIR, GPU_TOL, rel_err, run_cuda
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "render_*.npz"))):
    g = dict(np.load(p))
    H, W, n, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n, S, seed, val, flags=flags)
    print(os.path.basename(p), "flags", flags)
    for k in GPU_TOL:
        a = out[k].numpy(); b = g[k]
        extra = ""
        if k + "_sens" in g:
            from cases import per_ray, SENS_FACTOR, SENS_FLOOR
            sc = np.abs(b).max(); err = per_ray(a.astype(np.float64) - b, k); sens = g[k + "_sens"]
            allowed = GPU_TOL[k] * sc + SENS_FACTOR * np.maximum(0, sens - SENS_FLOOR * sc)
            st = sens <= SENS_FLOOR * sc
            extra = f" over-gate={int((err > allowed).sum())} stable-frac={st.mean():.2f} max-err-on-stable={(err[st].max() / sc if st.any() else 0):.2e} worst err/allowed={(err / allowed).max():.2f}"
        print(f"   {k:18s} {rel_err(a,b):.3e}  (tol {GPU_TOL[k]})  nan={np.isnan(a).sum()}{extra}")
    am, gm = out["at_wt_max"].numpy(), g["at_wt_max"]
    print("   at_wt_max mismatches", int((am != gm).sum()), "of", am.size,
          " mask_c2", int((out["mask_c2"].numpy() != g["mask_c2"]).sum()),
          " match", int((out["matchability_cycle_mask"].numpy() != g["matchability_cycle_mask"]).sum()))
    # where is the rgb error?
    e = np.abs(out["rgb"].numpy() - g["rgb"]).max(axis=-1)[0, 0]
    worst = np.argsort(e)[-5:]
    print("   worst rays", worst, e[worst], "valid", g["valid_mask"][0, worst, 0])
    ew = np.abs(out["at_wt"].numpy() - g["at_wt"]).max(axis=-1)
    print("   at_wt err per view max", ew.max(axis=1), " at rays", ew.argmax(axis=1))
