"""ufc_forward (the host orchestration around the native closing stage) against the unmodified reference
UFC.forward, on CPU with the oracle standing in for the CUDA tail. Needs /root/reference (build container only)."""
import os
import sys

import pytest
import torch

REF = os.environ.get("COPONERF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_ufc_forward_matches_reference_module():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_goldens import import_reference
    import_reference()
    from models.aggregation import UFC
    from coponerf_b200.ufc import ufc_forward
    from oracle import conv4d_oracle, ufc_oracle

    def conv_block(x, wq, bq, ws, bs, gamma, beta, stride, pad):
        return conv4d_oracle.encoder4d_layer(x, dict(wq=wq, bq=bq, ws=ws, bs=bs, gamma=gamma, beta=beta), stride, pad)
    torch.manual_seed(0)
    fca = UFC().eval()
    g = torch.Generator().manual_seed(3)
    feat = [torch.randn(2, 512, 16, 16, generator=g), torch.randn(2, 256, 32, 32, generator=g),
            torch.randn(2, 128, 64, 64, generator=g)]
    with torch.no_grad():
        ref_feats, ref_flows, ref_c = fca(feat, 2)
        got_feats, got_flows, got_c = ufc_forward(fca, feat, 2, tail=ufc_oracle.ufc_tail, conv_block=conv_block,
                                                      attention=ufc_oracle.linear_attention)
    assert all('forward' not in m.__dict__ for m in fca.modules())   # patches removed
    for a, b in zip(got_feats, ref_feats):
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-5, rtol=1e-5)
    assert got_c.shape == ref_c.shape and (got_c - ref_c).abs().max() <= 1e-5
    for a, b in zip(got_flows, ref_flows):
        assert a.shape == b.shape and (a - b).abs().max() <= 1e-3 * 64


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
