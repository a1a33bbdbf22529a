"""Pin the UFC closing-stage oracle against outputs of the unmodified reference functions (tests/golden/ufc_tail_*.npz)."""
import glob
import os

import numpy as np
import pytest

from coponerf_b200 import synth
from oracle import ufc_oracle

CASES = sorted(os.path.basename(p)[:-4] for p in
               glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ufc_tail_*.npz")))


def run_case(g):
    sizes, out, batch, seed = tuple(int(v) for v in g["meta"][:3]), int(g["meta"][3]), int(g["meta"][4]), int(g["meta"][5])
    src, trg = synth.ufc_tail_features(sizes, batch, seed)
    return ufc_oracle.ufc_tail(src, trg, sizes, out)


@pytest.mark.parametrize("case", CASES)
def test_ufc_tail_oracle_matches_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    (flow, flow_flip, t2s, s2t), c = run_case(g)
    assert np.abs(c.reshape(-1)[g["c_idx"]].numpy() - g["c_val"]).max() <= 1e-6
    assert abs(float(c.double().mean()) - float(g["c_mean"])) <= 1e-8
    assert abs(float((c.double() ** 2).mean()) - float(g["c_sq"])) <= 1e-8
    for name, got in (("flow", flow), ("flow_flip", flow_flip), ("flow_t_to_s", t2s), ("flow_s_to_t", s2t)):
        assert got.shape == g[name].shape
        tol = 1e-5 if name.startswith("flow_") and "_to_" in name else 1e-5 * g[name].shape[-1]
        assert np.abs(got.numpy() - g[name]).max() <= tol, name


CONV_CASES = sorted(os.path.basename(p)[:-4] for p in
                    glob.glob(os.path.join(os.path.dirname(__file__), "golden", "conv4d_*.npz")))


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv4d_oracle_matches_reference(case, golden_dir):
    from oracle import conv4d_oracle
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    x, layers, stride, pad = synth.conv4d_case(case)
    y = x
    for p in layers:
        y = conv4d_oracle.encoder4d_layer(y, p, stride, pad)
    assert tuple(y.shape) == tuple(g["shape"])
    assert np.abs(y.reshape(-1)[g["idx"]].numpy() - g["val"]).max() <= 1e-5
    assert abs(float(y.double().mean()) - float(g["mean"])) <= 1e-6


LINATT = sorted(os.path.basename(p)[:-4] for p in
                glob.glob(os.path.join(os.path.dirname(__file__), "golden", "linatt_*.npz")))


@pytest.mark.parametrize("case", LINATT)
def test_linear_attention_oracle_matches_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    y = ufc_oracle.linear_attention(*synth.linatt_case(case))
    assert tuple(y.shape) == tuple(g["shape"])
    assert np.abs(y.reshape(-1)[g["idx"]].numpy() - g["val"]).max() <= 1e-6 * float(np.abs(g["val"]).max())
