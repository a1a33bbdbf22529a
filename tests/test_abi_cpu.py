"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the host mirror keeps the reference's state_dict names. No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from coponerf_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        from coponerf_b200.build import build_library
        build_library()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    header = open(os.path.join(ROOT, "include", "coponerf_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cpn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/coponerf_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_struct_layout_and_version():
    _ensure_built()
    lib = _lib.load()
    assert lib.cpn_version() >= 100
    assert lib.cpn_sizeof_render_args() == ctypes.sizeof(_lib.RenderArgs)


def test_weight_table_matches_reference_state_dict_names():
    _ensure_built()
    lib = _lib.load()
    names = _lib.weight_names()
    assert names == list(synth.RENDER_PARAM_SHAPES)
    total = 0
    for i, n in enumerate(names):
        numel = 1
        for d in synth.RENDER_PARAM_SHAPES[n]:
            numel *= d
        assert lib.cpn_weight_numel(i) == numel, n
        total += numel
    assert lib.cpn_raw_weights_floats() == total
    assert lib.cpn_packed_weights_bytes() >= total * 4


def test_workspace_query_scales_with_rows():
    _ensure_built()
    lib = _lib.load()
    a = lib.cpn_render_workspace_bytes(1, 4096, 256, 64, 1)
    b = lib.cpn_render_workspace_bytes(2, 4096, 256, 64, 1)
    c = lib.cpn_render_workspace_bytes(1, 4096, 256, 128, 1)
    assert 0 < a < b and abs(b - c) < 0.05 * b
    assert a < lib.cpn_render_workspace_bytes(1, 4096, 256, 64, 3) < 3 * a
    assert lib.cpn_render_workspace_bytes(0, 4096, 256, 64, 1) == 0


def test_host_model_keeps_reference_parameter_names():
    from coponerf_b200.model import CoPoNeRF
    m = CoPoNeRF(n_view=2)
    sd = m.state_dict()
    for name, shape in synth.RENDER_PARAM_SHAPES.items():
        assert name in sd and tuple(sd[name].shape) == shape, name
    # parameters the reference declares but forward() never reads still load (models/CoPoNeRF.py:73,89-99)
    for name in ("corr_embed.weight", "latent_avg_query.weight", "latent_avg_key_2.bias",
                 "latent_avg_repeat_query.weight"):
        assert name in sd


def test_no_cpu_fallback():
    from coponerf_b200.model import CoPoNeRF
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = CoPoNeRF(n_view=2)
    m.H = m.W = 64
    inp = synth.make_input(64, 64, 8)
    z, rel, flow = synth.make_features(64, 64)
    with pytest.raises(RuntimeError):
        m(inp, z=z, rel_pose=rel, flow=flow, val=True)
