"""The state_dict-driven orchestration of the cost aggregation (coponerf_b200/ufc_native.py, with the PyTorch operator set)
and the independent oracle against the unmodified reference UFC.forward on CPU. Needs /root/reference (build container only)."""
import os
import sys

import pytest
import torch

REF = os.environ.get("COPONERF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_state_dict_driven_ufc_matches_reference_module():
    """ufc_native.ufc_forward (no reference modules, only the state_dict and an operator set) == UFC.forward."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_goldens import import_reference
    import_reference()
    from models.aggregation import UFC
    from coponerf_b200 import ufc_native
    from oracle.ufc_ops_torch import TorchOps
    torch.manual_seed(1)
    fca = UFC().eval()
    g = torch.Generator().manual_seed(4)
    feat = [torch.randn(2, 512, 16, 16, generator=g), torch.randn(2, 256, 32, 32, generator=g),
            torch.randn(2, 128, 64, 64, generator=g)]
    with torch.no_grad():
        ref_feats, ref_flows, ref_c = fca(feat, 2)
    got_feats, got_flows, got_c = ufc_native.ufc_forward(fca.state_dict(), feat, 2, TorchOps())
    for a, b in zip(got_feats, ref_feats):
        assert a.shape == b.shape and torch.allclose(a, b, atol=2e-5, rtol=1e-5)
    assert got_c.shape == ref_c.shape and (got_c - ref_c).abs().max() <= 1e-5
    for a, b in zip(got_flows, ref_flows):
        assert a.shape == b.shape and (a - b).abs().max() <= 1e-3 * 64


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_synthetic_ufc_state_dict_has_the_reference_names_and_shapes():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_goldens import import_reference
    import_reference()
    from models.aggregation import UFC
    from coponerf_b200 import synth
    ref = {k: tuple(v.shape) for k, v in UFC().state_dict().items()}
    assert synth.ufc_param_shapes() == ref


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_independent_ufc_oracle_matches_reference_module():
    """oracle/ufc_forward_oracle.py (reference formulation, no product code) == the unmodified UFC.forward."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_goldens import import_reference
    import_reference()
    from models.aggregation import UFC
    from oracle import ufc_forward_oracle
    torch.manual_seed(2)
    fca = UFC().eval()
    g = torch.Generator().manual_seed(5)
    feat = [torch.randn(2, 512, 16, 16, generator=g), torch.randn(2, 256, 32, 32, generator=g),
            torch.randn(2, 128, 64, 64, generator=g)]
    with torch.no_grad():
        ref_feats, ref_flows, ref_c = fca(feat, 2)
    got_feats, got_flows, got_c = ufc_forward_oracle.ufc_forward(fca.state_dict(), feat, 2)
    for a, b in zip(got_feats, ref_feats):
        assert a.shape == b.shape and torch.allclose(a, b, atol=2e-5, rtol=1e-5)
    assert got_c.shape == ref_c.shape and (got_c - ref_c).abs().max() <= 1e-5
    for a, b in zip(got_flows, ref_flows):
        assert a.shape == b.shape and (a - b).abs().max() <= 1e-3 * 64
