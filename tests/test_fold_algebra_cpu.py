"""The algebraic identities behind the folded render path, checked in float64 on CPU against the layer-by-layer form of
models/CoPoNeRF.py:387-485 (no GPU, no reference checkout). The CUDA side computes the same folded matrices in
cpn_pack_weights (coponerf_b200/csrc/weights.cu: fold_kernel, bilinear_fold_kernel); the GPU tests hold the result to the
reference goldens."""
import torch

from coponerf_b200 import synth


def _w(sd, name):
    w = sd[name + ".weight"].double()
    return w.reshape(w.shape[0], -1), sd[name + ".bias"].double()


def test_folded_value_and_key_layers_equal_the_three_layer_chain():
    """query_encode_latent_2 has no activation, so latent_value / key_map of the concatenated (primary, secondary)
    encodings are single layers on the concatenated 832-wide hidden vectors."""
    sd = synth.render_state_dict(0)
    W2, b2 = _w(sd, "query_encode_latent_2")
    Wv, bv = _w(sd, "latent_value")
    Wk, bk = _w(sd, "key_map")
    g = torch.Generator().manual_seed(0)
    hp, hs = torch.rand(50, 832, generator=g, dtype=torch.float64), torch.rand(50, 832, generator=g, dtype=torch.float64)
    e = torch.cat((hp @ W2.T + b2, hs @ W2.T + b2), dim=1)                  # CoPoNeRF.py:393-401
    for Wx, bx in ((Wv, bv), (Wk, bk)):
        WF = torch.cat((Wx[:, :416] @ W2, Wx[:, 416:] @ W2), dim=1)        # fold_kernel
        bF = bx + Wx[:, :416] @ b2 + Wx[:, 416:] @ b2
        assert torch.allclose(e @ Wx.T + bx, torch.cat((hp, hs), dim=1) @ WF.T + bF, rtol=1e-12, atol=1e-12)


def test_late_readout_equals_readout_of_values():
    """sum_rows w V = WVF (sum_rows w h) + b when the weights sum to one; z = R2 + 2 R1 (CoPoNeRF.py:456-485)."""
    g = torch.Generator().manual_seed(1)
    WVF, b = torch.randn(416, 1664, generator=g, dtype=torch.float64), torch.randn(416, generator=g, dtype=torch.float64)
    h = torch.rand(128, 1664, generator=g, dtype=torch.float64)            # the 2S rows of one ray
    w1 = torch.softmax(torch.randn(128, generator=g, dtype=torch.float64) * 4, 0)
    w2 = torch.softmax(torch.randn(128, generator=g, dtype=torch.float64) * 4, 0)
    V = h @ WVF.T + b
    R1, R2 = w1 @ V, w2 @ V
    assert torch.allclose(R1, (w1 @ h) @ WVF.T + b, rtol=1e-12, atol=1e-12)
    z_ref = (w2[:64] @ V[:64] + R1) + (w2[64:] @ V[64:] + R1)              # per view, then summed over the views
    assert torch.allclose(z_ref, ((w2 @ h) @ WVF.T + b) + 2 * R1, rtol=1e-12, atol=1e-12)
    assert torch.allclose(z_ref, R2 + 2 * R1, rtol=1e-12, atol=1e-12)


def test_bilinear_logits_equal_the_two_layer_dot_product():
    """<Wa k + ba, Wq q + bq> = k^T (WM q + BM) + (WS . q + CS) for key_map_2 and query_repeat_embed_2 against
    query_embed_2 (CoPoNeRF.py:408,446,450,473-474)."""
    sd = synth.render_state_dict(0)
    Wq, bq = _w(sd, "query_embed_2")
    g = torch.Generator().manual_seed(2)
    k, q = torch.rand(40, 128, generator=g, dtype=torch.float64), torch.rand(40, 128, generator=g, dtype=torch.float64)
    for name in ("key_map_2", "query_repeat_embed_2"):
        Wa, ba = _w(sd, name)
        ref = ((k @ Wa.T + ba) * (q @ Wq.T + bq)).sum(-1)
        WM, BM, WS, CS = Wa.T @ Wq, Wa.T @ bq, Wq.T @ ba, ba @ bq        # bilinear_fold_kernel
        got = (k * (q @ WM.T + BM)).sum(-1) + q @ WS + CS
        assert torch.allclose(ref, got, rtol=1e-12, atol=1e-12)


def test_round2_query_bias_is_a_linear_map_of_the_hidden_layer_and_one_readout_gives_z():
    """latent_value, encode_latent and the z_embed columns of query_repeat_embed have no activation between them
    (CoPoNeRF.py:404,463-472), so the per-ray bias of round 2 is sum_rows w1 (G h + g0) with G = Wqr[:, :128] We WVF
    (weights.cu: gfold_kernel), and z = R2 + 2 R1 = WVF sum_rows (w2 + 2 w1) h + 3 b."""
    sd = synth.render_state_dict(0)
    W2, b2 = _w(sd, "query_encode_latent_2")
    Wv, bv = _w(sd, "latent_value")
    We, be = _w(sd, "encode_latent")
    Wqr, bqr = _w(sd, "query_repeat_embed")
    WVF = torch.cat((Wv[:, :416] @ W2, Wv[:, 416:] @ W2), dim=1)
    bVF = bv + Wv[:, :416] @ b2 + Wv[:, 416:] @ b2
    g = torch.Generator().manual_seed(3)
    h = torch.rand(128, 1664, generator=g, dtype=torch.float64)
    w1 = torch.softmax(torch.randn(128, generator=g, dtype=torch.float64) * 4, 0)
    w2 = torch.softmax(torch.randn(128, generator=g, dtype=torch.float64) * 4, 0)
    local = torch.rand(128, 16, generator=g, dtype=torch.float64)
    V = h @ WVF.T + bVF
    R1 = w1 @ V                                                             # CoPoNeRF.py:456-461
    z_embed = We @ R1 + be                                                  # :467
    q2_ref = torch.relu(torch.cat((z_embed.expand(128, 128), local), dim=1) @ Wqr.T + bqr)   # :468-473
    G = Wqr[:, :128] @ We @ WVF
    g0 = Wqr[:, :128] @ (We @ bVF + be) + bqr
    rbias = w1 @ (h @ G.T + g0)
    q2 = torch.relu(local @ Wqr[:, 128:].T + rbias)
    assert torch.allclose(q2_ref, q2, rtol=1e-11, atol=1e-11)
    R2 = w2 @ V
    z = ((w2 + 2 * w1) @ h) @ WVF.T + 3 * bVF
    assert torch.allclose(R2 + 2 * R1, z, rtol=1e-11, atol=1e-11)
