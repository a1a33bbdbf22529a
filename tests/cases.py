"""Shared helpers of the parity tests: seeded cases, error metrics, oracle / CUDA runners."""
import os

import numpy as np
import torch

from coponerf_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED_POSE = {1: "frontal", 2: "oblique", 3: "oblique", 4: "mild", 5: "mild"}

# max|a-b| / max|b| gates for the CUDA path against the reference's fp32 outputs. north_star asks for 1e-4
# relative on the fp32 outputs; depth_ray / T_to_C*_pts are ill-conditioned (the reference's own fp32 and
# fp64 runs differ by 6.5e-4 and 3.5e-3) and get the documented looser gates.
GPU_TOL = {"rgb": 1e-4, "pixel_val": 1e-5, "coords": 1e-6, "at_wt": 1e-4, "valid_mask": 0.0,
           "depth_ray": 2e-3, "T_to_C1_pts": 1e-2, "T_to_C2_pts": 1e-2, "C2_pts_to_C1": 1e-2,
           "rel_pose_flip": 1e-6, "gt_rel_pose": 1e-6, "gt_rel_pose_flip": 1e-6}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make_case(H, W, n_rays, seed, pose=None, batch=1):
    pose = pose or SEED_POSE.get(seed, "mild")
    inp = synth.make_input(H, W, n_rays, seed=seed, pose_set=pose, batch=batch)
    z, rel_pose, flow = synth.make_features(H, W, seed=seed, batch=batch)
    return inp, z, rel_pose, flow


def to_device(obj, dev):
    if isinstance(obj, torch.Tensor):
        return obj.to(dev)
    if isinstance(obj, dict):
        return {k: to_device(v, dev) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_device(v, dev) for v in obj)
    return obj


SENS_KEYS = ("rgb", "at_wt", "depth_ray", "T_to_C1_pts", "T_to_C2_pts", "C2_pts_to_C1")
SENS_FACTOR = 4.0      # four perturbation samples understate the worst case by a few x (measured on B200: <= 4)
SENS_FLOOR = 1e-6      # rays whose reference output moves less than this (relative) get the plain gate
PARITY_LOG = []        # one line per (case, key) with the raw error statistics; printed by conftest.py at the end of a run


def per_ray(d, key):
    """max |delta| per ray, reduced like tests/golden/make_goldens.py:sensitivity."""
    d = np.abs(np.asarray(d, dtype=np.float64))
    if key == "rgb":
        return d.max(axis=-1)[:, 0]
    if key == "at_wt":
        return d.max(axis=-1)
    return d.reshape(d.shape[0], d.shape[1], -1).max(axis=-1)


def run_oracle(H, W, n_rays, S, seed, val, chunk=None, pose=None, batch=1, with_sens=False, ray_idx=None):
    from oracle import render_oracle
    inp, z, rel_pose, flow = make_case(H, W, n_rays, seed, pose, batch)
    if ray_idx is not None:     # a subset of the rays of the case (rays are independent)
        inp["query"]["uv"] = inp["query"]["uv"][:, :, ray_idx].contiguous()
        inp["query"]["rgb"] = inp["query"]["rgb"][:, :, ray_idx].contiguous()
    sd = synth.render_state_dict(0)
    base = render_oracle.render_forward(sd, inp, z, rel_pose, flow, H, W, S, bool(val), chunk=chunk)
    if with_sens:   # the same 1-ulp pose perturbation the golden generator applies to the reference
        import copy
        g = torch.Generator().manual_seed(1234)
        for _ in range(4):
            pin = copy.deepcopy(inp)
            for grp in ("context", "query"):
                t = pin[grp]["cam2world"]
                pin[grp]["cam2world"] = t * (1 + 6e-8 * torch.randn(t.shape, generator=g))
            rp = rel_pose * (1 + 6e-8 * torch.randn(rel_pose.shape, generator=g))
            out = render_oracle.render_forward(sd, pin, z, rp, flow, H, W, S, bool(val), chunk=chunk)
            for k in SENS_KEYS:
                d = per_ray(out[k].numpy() - base[k].numpy(), k)
                base[k + "_sens"] = np.maximum(base.get(k + "_sens", 0), d)
    return base


_MODELS = {}


def cuda_model(S=64, chunk_rays=2048):
    """coponerf_b200.CoPoNeRF with the seeded render weights on cuda:0 (cached per S / chunk)."""
    from coponerf_b200.model import CoPoNeRF
    key = (S, chunk_rays)
    if key not in _MODELS:
        m = CoPoNeRF(n_view=2, npoints=S, chunk_rays=chunk_rays)
        missing, unexpected = m.load_state_dict(synth.render_state_dict(0), strict=False)
        assert not unexpected, unexpected
        _MODELS[key] = m.cuda().eval()
    return _MODELS[key]


def run_cuda(H, W, n_rays, S, seed, val, chunk_rays=2048, pose=None, batch=1, flags=0, ray_slice=None):
    """The CUDA path through the drop-in forward() (which calls the C-ABI). Returns CPU tensors."""
    inp, z, rel_pose, flow = make_case(H, W, n_rays, seed, pose, batch)
    if ray_slice is not None:
        inp["query"]["uv"] = inp["query"]["uv"][:, :, ray_slice].contiguous()
    dev = torch.device("cuda:0")
    m = cuda_model(S, chunk_rays)
    m.H, m.W = H, W
    m.engine().flags = flags
    out = m(to_device(inp, dev), z=to_device(z, dev), rel_pose=rel_pose.to(dev), flow=to_device(flow, dev), val=bool(val))
    torch.cuda.synchronize()
    return {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def check_against(out, ref, tag, tol=GPU_TOL, min_stable=0.5):
    """Float outputs within `tol` (relative to max|ref|); integer / boolean outputs exact up to float-noise ties.

    The reference is ill-conditioned on a minority of rays: triangulating near-parallel rays amplifies
    one-ulp differences of the fp32 pose arithmetic by up to 1e4 (measured per ray on
    the reference itself by the golden generator, `<key>_sens`). When `ref` carries that measurement the gate
    of a ray is tol + SENS_FACTOR * max(0, sens - floor): the plain gate wherever the reference is stable.
    """
    get = lambda d, k: np.asarray(d[k].numpy() if isinstance(d[k], torch.Tensor) else d[k])
    allowed_by_key = {}
    for k, t in tol.items():
        a, b = get(out, k), get(ref, k)
        assert a.shape == b.shape, (tag, k, a.shape, b.shape)
        scale = max(float(np.abs(b).max()), 1e-30) if b.size else 1.0
        if k + "_sens" in ref and b.size:
            sens = np.asarray(ref[k + "_sens"], dtype=np.float64)
            err = per_ray(a.astype(np.float64) - b.astype(np.float64), k)
            allowed = t * scale + SENS_FACTOR * np.maximum(0.0, sens - SENS_FLOOR * scale)
            bad = err > allowed
            allowed_by_key[k] = allowed
            # raw statistics next to the gate: how large the error really is, how many rays sit on a widened gate and how
            # many of them needed it (error above the plain gate)
            widened = allowed > 2.0 * t * scale
            needed = err > t * scale
            PARITY_LOG.append(f"{tag}:{k} max {(err / scale).max():.2e} p99 {np.quantile(err / scale, 0.99):.2e} "
                              f"median {np.median(err / scale):.2e} (gate {t:.0e}); rays with gate > 2x plain "
                              f"{widened.mean():.4f}, rays above the plain gate {needed.mean():.4f}")
            assert not bad.any(), (f"{tag}:{k} {int(bad.sum())} rays over the gate; worst err/scale "
                                   f"{(err / scale).max():.3e}, worst excess {(err - allowed).max() / scale:.3e}")
            # the widening is an allowance for a minority: at most 5 % of the rays may need it at all
            assert needed.mean() <= 0.05, f"{tag}:{k} {needed.mean():.3f} of the rays are above the plain gate {t}"
            if k == "rgb" and min_stable is not None:   # the plain gate must cover most of the image, or the case is a poor test
                stable = sens <= SENS_FLOOR * scale
                assert stable.mean() >= min_stable, f"{tag}: only {stable.mean():.2f} of the rays are well-conditioned"
        else:
            e = rel_err(a, b)
            PARITY_LOG.append(f"{tag}:{k} max {e:.2e} (gate {t:.0e})")
            assert e <= t, f"{tag}:{k} rel err {e:.3e} > {t}"
    # argmax of the round-1 weights: may differ only where the top two reference weights tie within noise
    am, gm, w = get(out, "at_wt_max"), get(ref, "at_wt_max"), get(ref, "at_wt")
    assert am.shape == gm.shape and am.dtype == gm.dtype, (tag, am.shape, gm.shape, am.dtype, gm.dtype)
    d = np.nonzero(am[..., 0] != gm[..., 0])
    if d[0].size:
        top2 = np.sort(w[d[0], d[1]], axis=-1)[:, -2:]
        # a tie within noise: 2e-4 relative, plus twice the weight error this ray is allowed (ill-conditioned rays only)
        slack = 2.0 * allowed_by_key["at_wt"][d[0], d[1]] if "at_wt" in allowed_by_key else 0.0
        assert np.all((top2[:, 1] - top2[:, 0]) <= 2e-4 * top2[:, 1] + slack), f"{tag}: at_wt_max differs off a tie"
        assert d[0].size <= max(2, 0.01 * am.size), f"{tag}: too many at_wt_max ties ({d[0].size})"
    # masks looked up at trunc(T_to_C2_pts): may differ only where that point sits on a pixel boundary
    c2 = get(ref, "T_to_C2_pts").astype(np.float64)
    near = (np.abs(c2 - np.round(c2)) <= 2e-3 * np.maximum(1.0, np.abs(c2))).any(axis=-1)
    for k in ("mask_c2", "matchability_cycle_mask"):
        a, b = get(out, k), get(ref, k)
        assert a.shape == b.shape and a.dtype == b.dtype, (tag, k)
        bad = (a != b) & ~near
        assert not bad.any(), f"{tag}:{k} differs at {int(bad.sum())} rays away from pixel boundaries"
