"""Parity of the sm_100a render path (through the C-ABI, via the drop-in forward()) against
(1) the committed outputs of the unmodified reference (tests/golden/*.npz) and (2) the CPU oracle,
plus the size-independent properties: chunk / shard invariance and ray independence, bit-exact."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from cases import GOLDEN_DIR, GPU_TOL, check_against, rel_err, run_cuda, run_oracle

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "render_*.npz")))
EXACT_KEYS = ("rgb", "valid_mask", "depth_ray", "at_wt", "at_wt_max", "pixel_val", "coords", "T_to_C1_pts",
              "T_to_C2_pts", "C2_pts_to_C1", "mask_c2", "matchability_cycle_mask")


@pytest.mark.parametrize("case", CASES)
def test_matches_reference_golden(case):
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val)
    check_against(out, g, case)


@pytest.mark.parametrize("case", CASES)
def test_f16x3_scheme_matches_reference_golden(case):
    """Tensor-core GEMMs with three fp16 MMAs per product (CPN_FLAG_F16X3) instead of fp16 + two e4m3 corrections."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val, flags=2)
    check_against(out, g, case + "/f16x3")


@pytest.mark.parametrize("case", CASES)
def test_simt_cross_check_path_matches_reference_golden(case):
    """The fp32 CUDA-core GEMM path (CPN_FLAG_SIMT_ONLY) is held to the same gates."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val, flags=1)
    check_against(out, g, case + "/simt")


def test_matches_oracle_batch2_ragged():
    """Two pairs, a ray count that is not a multiple of any tile, oblique pose (invalid rays present)."""
    args = dict(H=64, W=64, n_rays=333, S=64, seed=2, val=True, batch=2)
    ref = run_oracle(with_sens=True, **args)
    out = run_cuda(chunk_rays=100, **args)
    check_against(out, ref, "batch2")
    v = ref["valid_mask"].numpy()
    assert 0.05 < v.mean() < 0.95  # the degenerate-ray paths are exercised


def test_chunk_invariance_bit_exact():
    """A ray's outputs do not depend on the chunk it is rendered in."""
    args = dict(H=64, W=64, n_rays=512, S=64, seed=2, val=True)
    a = run_cuda(chunk_rays=2048, **args)
    b = run_cuda(chunk_rays=96, **args)
    for k in EXACT_KEYS:
        assert torch.equal(a[k], b[k]), k


def test_ray_shard_invariance_bit_exact():
    """Rendering a slice of the rays alone (what a rank does) equals the slice of the full render."""
    args = dict(H=64, W=64, n_rays=512, S=64, seed=1, val=True)
    full = run_cuda(**args)
    lo, hi = 130, 389
    part = run_cuda(ray_slice=slice(lo, hi), **args)
    cat_dim = {"pixel_val": 1, "at_wt": 1, "at_wt_max": 1, "coords": 1, "mask_c2": 1, "matchability_cycle_mask": 1,
               "rgb": 2}
    for k in EXACT_KEYS:
        d = cat_dim.get(k, 1)
        assert torch.equal(part[k], full[k].narrow(d, lo, hi - lo)), k


def test_full_resolution_properties():
    """BASELINE config 2 size (256x256, all 65 536 rays): properties that need no oracle run."""
    out = run_cuda(256, 256, None, 64, seed=4, val=True, chunk_rays=2048)
    n = 65536
    assert out["rgb"].shape == (1, 1, n, 3) and out["at_wt"].shape == (2, n, 64)
    assert torch.isfinite(out["rgb"]).all()
    w = out["at_wt"].view(1, 2, n, 64)
    assert torch.allclose(w.sum(dim=(1, 3)), torch.ones(1, n), atol=1e-5)   # joint softmax over 2 x S
    inval = out["valid_mask"][0, :, 0] == 0
    assert (out["rgb"][0, 0][inval] == 1).all()                            # white where no epipolar segment
    assert out["at_wt_max"].min() >= 0 and out["at_wt_max"].max() < 64
    assert out["depth_ray"].min() >= 0 and out["depth_ray"].max() <= 10
    # the committed golden is a 384-ray subset of this very case: same uv order is not guaranteed, so compare
    # through a second full render in a different chunking instead (bit-exact)
    again = run_cuda(256, 256, None, 64, seed=4, val=True, chunk_rays=1000)
    for k in ("rgb", "at_wt_max", "pixel_val", "mask_c2"):
        assert torch.equal(out[k], again[k]), k


def test_empty_ray_set():
    out = run_cuda(64, 64, 16, 64, seed=1, val=True, ray_slice=slice(0, 0))
    assert out["rgb"].shape == (1, 1, 0, 3) and out["at_wt"].shape == (2, 0, 64)


def test_bad_arguments_fail_loudly():
    from coponerf_b200 import _lib
    with pytest.raises(_lib.CpnError):
        run_cuda(64, 64, 16, 40, seed=1, val=True)   # S not a multiple of 32
    lib = _lib.load()
    assert lib.cpn_render_rays(None, None) != 0
    assert b"null" in lib.cpn_last_error()


def test_gemm_simt_against_torch_fp32():
    from coponerf_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    for (M, N, K, relu) in [(300, 832, 848, 1), (257, 416, 832, 0), (1000, 128, 16, 1), (64, 128, 416, 0)]:
        A = torch.randn(M, K, device="cuda")
        Wm = torch.randn(N, K, device="cuda") / K ** 0.5
        bias = torch.randn(N, device="cuda")
        C = torch.empty(M, N, device="cuda")
        wt = Wm.t().contiguous()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(lib.cpn_gemm_simt(p(A), K, p(wt), p(bias), p(C), N, M, N, K, relu, st), "cpn_gemm_simt")
        ref = torch.nn.functional.linear(A.double(), Wm.double(), bias.double())
        if relu:
            ref = ref.relu()
        assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 2e-6, (M, N, K)


def test_pair_prologue_against_oracle():
    from oracle import render_oracle
    from coponerf_b200 import synth
    from cases import cuda_model, to_device
    inp = synth.make_input(64, 64, 8, seed=3, pose_set="oblique", batch=2)
    z, rel, flow = synth.make_features(64, 64, seed=3, batch=2)
    flow = tuple(f * 2 for f in flow)  # large enough flows to hit both branches of the masks
    up2, mask = render_oracle.pair_prologue(flow, inp["context"]["rgb"].shape[-2])
    m = cuda_model()
    dev = torch.device("cuda:0")
    st = m.engine().prepare_pair(to_device(inp, dev), to_device(z, dev), rel.to(dev), to_device(flow, dev), 64, 64, True)
    torch.cuda.synchronize()
    scale = 256 / inp["context"]["rgb"].shape[-2]
    assert rel_err((st.up_flow2 * scale).cpu().numpy(), up2.numpy()) < 1e-6
    got = st.mask_padded2.cpu().bool()
    # a mask bit may differ only where the cycle error sits on the threshold
    assert (got != mask).float().mean() < 1e-3
    assert 0.02 < mask.float().mean() < 0.98


TC_LAYERS = [(0, "query_encode_latent", 1), (1, "query_encode_latent_2", 0), (2, "latent_value", 0), (3, "key_map", 1),
             (4, "key_map_2", 0), (5, "query_embed_2", 0), (6, "query_repeat_embed_2", 0)]


def _tc_setup(name):
    from coponerf_b200 import _lib, synth
    from cases import cuda_model
    lib = _lib.load()
    eng = cuda_model().engine()
    sd = synth.render_state_dict(0)
    Wm = sd[name + ".weight"].reshape(sd[name + ".weight"].shape[0], -1).cuda()
    return lib, eng, Wm, sd[name + ".bias"].cuda()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("scheme", [0, 4], ids=["f16+f8", "f16x3"])
@pytest.mark.parametrize("layer,name,relu", TC_LAYERS)
def test_gemm_tc_against_torch_fp64(layer, name, relu, scheme):
    """tcgen05 GEMM of one packed layer vs fp64. Three fp16 MMAs per product sit at fp32 level (4e-6); the default
    fp16 + two e4m3 correction MMAs at about 2e-5 -- both far below a single fp16/TF32 pass (5e-4)."""
    from coponerf_b200 import _lib
    lib, eng, Wm, bias = _tc_setup(name)
    N, K = Wm.shape
    lda = 848 if K == 835 else K
    torch.manual_seed(layer)
    for M in (1, 128, 300, 1000, 1024):
        A = torch.zeros(M, lda, device="cuda")
        A[:, :K] = torch.randn(M, K, device="cuda") * 3
        C = torch.full((M, N), float("nan"), device="cuda")
        _lib.check(lib.cpn_gemm_tc(_p(eng.weights), layer, _p(A), lda, _p(C), N, M, relu, scheme, 1, 1, _st()), "cpn_gemm_tc")
        ref = torch.nn.functional.linear(A[:, :K].double(), Wm.double(), bias.double())
        if relu:
            ref = ref.relu()
        e = rel_err(C.cpu().numpy(), ref.cpu().numpy())
        print(f"gemm_tc {name} scheme={scheme} M={M}: rel err {e:.2e}")
        assert e < (1e-5 if scheme else 6e-5), (name, M, e)


@pytest.mark.parametrize("R", [300, 1024], ids=["ragged", "whole-pairs"])
@pytest.mark.parametrize("scheme", [0, 4, 12, 16, 4096, 4096 + 4],
                         ids=["f16+f8", "f16x3", "f16x3/cluster-multicast", "f16+f8/cta-pairs", "f16+f8/sub-tile-pipelined",
                              "f16x3/sub-tile-pipelined"])
def test_gemm_tc_operand_image_chain(scheme, R):
    """fp32 -> [GEMM1] -> hi/lo operand image -> [GEMM2, two branches side by side] -> image -> [latent_value] -> fp32,
    the way cpn_render_rays chains the encoder layers, against fp64."""
    from coponerf_b200 import _lib
    lib, eng, W1, b1 = _tc_setup("query_encode_latent")
    _, _, W2, b2 = _tc_setup("query_encode_latent_2")
    _, _, WV, bV = _tc_setup("latent_value")
    torch.manual_seed(7)
    Rp = (R + 127) // 128 * 128
    x = torch.randn(2, R, 835, device="cuda")  # [branch][row]
    A = torch.zeros(2 * Rp, 848, device="cuda")
    rows = torch.arange(R, device="cuda")
    for br in range(2):
        A[(rows // 128) * 256 + br * 128 + rows % 128, :835] = x[br]
    chunk = _lib.ACT_CHUNK_BYTES
    H1 = torch.empty(2 * Rp // 128 * 26 * chunk, dtype=torch.uint8, device="cuda")
    E = torch.empty(Rp // 128 * 26 * chunk, dtype=torch.uint8, device="cuda")
    V = torch.full((R, 416), float("nan"), device="cuda")
    w = _p(eng.weights)
    _lib.check(lib.cpn_gemm_tc(w, 0, _p(A), 848, _p(H1), 0, 2 * Rp, 1, _lib.TC_OUT_IMAGE | scheme, 1, 26, _st()), "gemm1")
    _lib.check(lib.cpn_gemm_tc(w, 1, _p(H1), 0, _p(E), 0, 2 * Rp, 0, _lib.TC_A_IMAGE | _lib.TC_OUT_IMAGE | scheme, 2, 26, _st()), "gemm2")
    _lib.check(lib.cpn_gemm_tc(w, 2, _p(E), 0, _p(V), 416, R, 0, _lib.TC_A_IMAGE | scheme, 1, 1, _st()), "gemmV")
    lin = torch.nn.functional.linear
    h = lin(x.double(), W1.double(), b1.double()).relu()
    e = lin(h, W2.double(), b2.double())                      # (2, R, 416)
    ref = lin(torch.cat((e[0], e[1]), dim=-1), WV.double(), bV.double())
    err = rel_err(V.cpu().numpy(), ref.cpu().numpy())
    print(f"gemm_tc chain scheme={scheme}: rel err {err:.2e}")
    assert err < (2e-5 if scheme & 4 else 1e-4), err


def test_config4_shape_512_s128_properties():
    """BASELINE config 4 shape (512x512 features, 128 samples per ray): a 4096-ray slice of the image through the
    same kernels; compared with the CPU oracle on a subset and checked for chunk invariance."""
    args = dict(H=512, W=512, n_rays=4096, S=128, seed=5, val=True, pose="mild")
    out = run_cuda(chunk_rays=1024, **args)
    assert out["rgb"].shape == (1, 1, 4096, 3) and out["at_wt"].shape == (2, 4096, 128)
    assert torch.isfinite(out["rgb"]).all()
    w = out["at_wt"].view(1, 2, 4096, 128)
    assert torch.allclose(w.sum(dim=(1, 3)), torch.ones(1, 4096), atol=1e-5)
    again = run_cuda(chunk_rays=700, **args)
    for k in ("rgb", "at_wt_max", "pixel_val", "depth_ray"):
        assert torch.equal(out[k], again[k]), k
    # oracle on the first 256 rays (the same uv order: make_case draws the permutation before slicing)
    sub = run_cuda(ray_slice=slice(0, 256), **args)
    inp, z, rel, flow = __import__("cases").make_case(512, 512, 4096, 5, "mild")
    inp["query"]["uv"] = inp["query"]["uv"][:, :, :256].contiguous()
    inp["query"]["rgb"] = inp["query"]["rgb"][:, :, :256].contiguous()
    from coponerf_b200 import synth
    from oracle import render_oracle
    ref = render_oracle.render_forward(synth.render_state_dict(0), inp, z, rel, flow, 512, 512, 128, True)
    err = (sub["rgb"] - ref["rgb"]).abs().amax(dim=-1)[0, 0] / ref["rgb"].abs().max()
    assert float(err.median()) < 1e-5 and float(err.kthvalue(int(0.9 * 256)).values) < 1e-4
    assert torch.equal(sub["rgb"], out["rgb"][:, :, :256])


def test_gemm_tc_cb16_and_rowdot_epilogues():
    """query_embed_2 written column-blocked (CB16) and key_map_2 fused with the per-row dot product of the attention
    logits (models/CoPoNeRF.py:446,450) against fp64."""
    from coponerf_b200 import _lib
    lib, eng, Wq, bq = _tc_setup("query_embed_2")
    _, _, Wk, bk = _tc_setup("key_map_2")
    torch.manual_seed(3)
    M = 700
    Mp = (M + 127) // 128 * 128
    x = torch.randn(M, 128, device="cuda")
    k1 = torch.randn(M, 128, device="cuda").relu()
    qcb = torch.zeros(Mp * 128, device="cuda")
    w = _p(eng.weights)
    _lib.check(lib.cpn_gemm_tc(w, 5, _p(x), 128, _p(qcb), 0, M, 0, _lib.TC_OUT_CB16, 1, 1, _st()), "qe cb16")
    qe = torch.nn.functional.linear(x.double(), Wq.double(), bq.double())
    got = qcb.view(Mp // 128, 8, 128, 16).permute(0, 2, 1, 3).reshape(Mp, 128)[:M]
    assert rel_err(got.cpu().numpy(), qe.cpu().numpy()) < 6e-5
    out = torch.full((M,), float("nan"), device="cuda")
    _lib.check(lib.cpn_gemm_tc_rowdot(w, 4, _p(k1), 128, _p(qcb), _p(out), M, 0, 0, 11.31, _st()), "rowdot")
    kk = torch.nn.functional.linear(k1.double(), Wk.double(), bk.double())
    ref = (kk * qe).sum(-1) / 11.31
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("case", CASES)
def test_unfolded_encoder_tail_matches_reference_golden(case):
    """CPN_FLAG_NO_FOLD: query_encode_latent_2, latent_value and key_map as three GEMMs (the default folds the
    activation-free query_encode_latent_2 into the other two at pack time)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val, flags=4)
    check_against(out, g, case + "/no-fold")


@pytest.mark.parametrize("R", [300, 1024], ids=["ragged", "whole-pairs"])
@pytest.mark.parametrize("scheme", [0, 4], ids=["f16+f8", "f16x3"])
def test_gemm_tc_folded_value_and_key_layers(scheme, R):
    """fp32 -> [GEMM1] -> H1 image -> folded layers 7 (latent_value o query_encode_latent_2) and 8 (key_map o
    query_encode_latent_2) reading the H1 image as a K = 1664 operand, against the unfolded chain in fp64."""
    from coponerf_b200 import _lib
    lib, eng, W1, b1 = _tc_setup("query_encode_latent")
    _, _, W2, b2 = _tc_setup("query_encode_latent_2")
    _, _, WV, bV = _tc_setup("latent_value")
    _, _, WK, bK = _tc_setup("key_map")
    torch.manual_seed(11)
    Rp = (R + 127) // 128 * 128
    x = torch.randn(2, R, 835, device="cuda")  # [branch][row]
    A = torch.zeros(2 * Rp, 848, device="cuda")
    rows = torch.arange(R, device="cuda")
    for br in range(2):
        A[(rows // 128) * 256 + br * 128 + rows % 128, :835] = x[br]
    chunk = _lib.ACT_CHUNK_BYTES
    H1 = torch.empty(2 * Rp // 128 * 26 * chunk, dtype=torch.uint8, device="cuda")
    K1 = torch.empty(Rp // 128 * 4 * chunk, dtype=torch.uint8, device="cuda")
    V = torch.full((R, 416), float("nan"), device="cuda")
    Kk = torch.full((R, 128), float("nan"), device="cuda")
    w = _p(eng.weights)
    _lib.check(lib.cpn_gemm_tc(w, 0, _p(A), 848, _p(H1), 0, 2 * Rp, 1, _lib.TC_OUT_IMAGE | scheme, 1, 26, _st()), "gemm1")
    _lib.check(lib.cpn_gemm_tc(w, 7, _p(H1), 0, _p(V), 416, R, 0, _lib.TC_A_IMAGE | scheme, 1, 1, _st()), "gemmVF")
    _lib.check(lib.cpn_gemm_tc(w, 8, _p(H1), 0, _p(K1), 0, R, 1, _lib.TC_A_IMAGE | _lib.TC_OUT_IMAGE | scheme, 1, 4, _st()), "gemmKF")
    _lib.check(lib.cpn_gemm_tc(w, 4, _p(K1), 0, _p(Kk), 128, R, 0, _lib.TC_A_IMAGE | scheme, 1, 1, _st()), "gemmK2")
    lin = torch.nn.functional.linear
    h = lin(x.double(), W1.double(), b1.double()).relu()
    e = lin(h, W2.double(), b2.double())                      # (2, R, 416)
    cat = torch.cat((e[0], e[1]), dim=-1)
    ref_v = lin(cat, WV.double(), bV.double())
    _, _, WK2, bK2 = _tc_setup("key_map_2")
    ref_k = lin(lin(cat, WK.double(), bK.double()).relu(), WK2.double(), bK2.double())
    ev, ek = rel_err(V.cpu().numpy(), ref_v.cpu().numpy()), rel_err(Kk.cpu().numpy(), ref_k.cpu().numpy())
    print(f"folded layers scheme={scheme}: V rel err {ev:.2e}, key rel err {ek:.2e}")
    assert ev < (2e-5 if scheme == 4 else 1e-4) and ek < (2e-5 if scheme == 4 else 1e-4), (ev, ek)


@pytest.mark.parametrize("case", CASES)
def test_early_value_path_matches_reference_golden(case):
    """CPN_FLAG_EARLY_V: V formed per sample by one GEMM and read by the attention kernels (the default reads out the
    hidden layer and applies the folded latent_value once per ray)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val, flags=8)
    check_against(out, g, case + "/early-v")


@pytest.mark.parametrize("case", CASES[:2])
def test_late_readout_with_f16x3_images(case):
    """The late readout reading [fp16 hi | fp16 lo] operand images (CPN_FLAG_F16X3)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val, flags=2)
    check_against(out, g, case + "/late-f16x3")


@pytest.mark.parametrize("case", CASES)
def test_three_layer_logit_path_matches_reference_golden(case):
    """CPN_FLAG_NO_BILINEAR: key_map_2, query_embed_2 and query_repeat_embed_2 as three 128 x 128 layers (the default
    evaluates both attention logits as bilinear forms of the hidden vectors, one layer per round)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val, flags=16)
    check_against(out, g, case + "/no-bilinear")


@pytest.mark.parametrize("case", CASES)
def test_per_ray_chain_path_matches_reference_golden(case):
    """CPN_FLAG_NO_GFOLD: round-1 readout, latent_value per ray, encode_latent and the z half of query_repeat_embed as
    separate steps (the default folds them into a 128-wide per-row term next to the key hidden layer and reads the hidden
    layer out once with the weights w2 + 2 w1)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    out = run_cuda(H, W, n_rays, S, seed, val, flags=32)
    check_against(out, g, case + "/no-gfold")


@pytest.mark.parametrize("scheme", [0, 4, 1024, 4096, 8192],
                         ids=["f16+f8", "f16x3", "f16+f8/compact-image", "f16+f8/sub-tile-pipelined", "f16+f8/weight-stationary"])
def test_gemm_tc_key_and_round2_bias_layer(scheme):
    """Layer 10 = [key_map ; G] o query_encode_latent_2 over the hidden image (cpn_gemm_tc_kg): the key tile leaves as the
    round-1 logit, the G tile as fp32 rows G h + g0, against the layer-by-layer chain of CoPoNeRF.py:393-408,463-472 in
    fp64 (G h + g0 = query_repeat_embed's z_embed columns applied to encode_latent(latent_value(.)))."""
    from coponerf_b200 import _lib
    lib, eng, W1, b1 = _tc_setup("query_encode_latent")
    _, _, W2, b2 = _tc_setup("query_encode_latent_2")
    _, _, WV, bV = _tc_setup("latent_value")
    _, _, WK, bK = _tc_setup("key_map")
    _, _, WE, bE = _tc_setup("encode_latent")
    _, _, WQR, bQR = _tc_setup("query_repeat_embed")
    torch.manual_seed(13)
    R = 700
    Rp = (R + 127) // 128 * 128
    x = torch.randn(2, R, 835, device="cuda")
    A = torch.zeros(2 * Rp, 848, device="cuda")
    rows = torch.arange(R, device="cuda")
    for br in range(2):
        A[(rows // 128) * 256 + br * 128 + rows % 128, :835] = x[br]
    chunk = _lib.ACT_CHUNK_BYTES
    H1 = torch.empty(2 * Rp // 128 * 26 * chunk, dtype=torch.uint8, device="cuda")
    w = _p(eng.weights)
    if scheme == _lib.TC_A_IMAGE3:
        # the compact hidden image (12 KB blocks [fp16 head | remainder plane], no value plane): layer 10 derives the value
        # plane in shared memory. Built here by dropping the plane from the full image; cpn_render_rays writes it directly
        # (CPN_TC_OUT_IMAGE3, covered by test_compact_hidden_image_is_bit_identical_to_the_full_one)
        _lib.check(lib.cpn_gemm_tc(w, 0, _p(A), 848, _p(H1), 0, 2 * Rp, 1, _lib.TC_OUT_IMAGE, 1, 26, _st()), "gemm1")
        H1c = H1.view(-1, chunk)[:, :_lib.ACT_CHUNK3_BYTES].contiguous().view(-1)      # drop the value plane of every block
        scheme_kg = _lib.TC_A_IMAGE3
    else:
        _lib.check(lib.cpn_gemm_tc(w, 0, _p(A), 848, _p(H1), 0, 2 * Rp, 1, _lib.TC_OUT_IMAGE | scheme, 1, 26, _st()), "gemm1")
        H1c, scheme_kg = H1, scheme
    # dotv: CB16 with 16 blocks per row tile (as the bilinear-logit layer writes it), blocks 0-7 used
    dv = torch.randn(Rp, 256, device="cuda")
    dv_cb = dv.view(Rp // 128, 128, 16, 16).permute(0, 2, 1, 3).contiguous()
    rowadd = torch.randn(Rp, device="cuda")
    lg = torch.full((Rp,), float("nan"), device="cuda")
    gh = torch.full((Rp, 128), float("nan"), device="cuda")
    _lib.check(lib.cpn_gemm_tc_kg(w, _p(H1c), _p(dv_cb), 16, _p(rowadd), 11.31, _p(lg), _p(gh), R, scheme_kg, _st()), "kg")
    if scheme == _lib.TC_A_IMAGE3:      # bit-identical to the consumer of the full image
        lg2 = torch.full((Rp,), float("nan"), device="cuda")
        gh2 = torch.full((Rp, 128), float("nan"), device="cuda")
        _lib.check(lib.cpn_gemm_tc_kg(w, _p(H1), _p(dv_cb), 16, _p(rowadd), 11.31, _p(lg2), _p(gh2), R, 0, _st()), "kg full")
        assert torch.equal(lg[:R], lg2[:R]) and torch.equal(gh, gh2)
        scheme = 0
    if scheme in (_lib.TC_PERSIST2, _lib.TC_WS):      # the experiment kernels issue the same MMAs per accumulator: same bits
        lg2 = torch.full((Rp,), float("nan"), device="cuda")
        gh2 = torch.full((Rp, 128), float("nan"), device="cuda")
        _lib.check(lib.cpn_gemm_tc_kg(w, _p(H1), _p(dv_cb), 16, _p(rowadd), 11.31, _p(lg2), _p(gh2), R, 0, _st()), "kg default")
        assert torch.equal(lg[:R], lg2[:R]) and torch.equal(gh, gh2)
        scheme = 0
    lin = torch.nn.functional.linear
    h = lin(x.double(), W1.double(), b1.double()).relu()
    e = lin(h, W2.double(), b2.double())
    cat = torch.cat((e[0], e[1]), dim=-1)
    k1 = lin(cat, WK.double(), bK.double()).relu()
    ref_lg = ((k1 * dv[:R, :128].double()).sum(-1) + rowadd[:R].double()) / 11.31
    v = lin(cat, WV.double(), bV.double())
    ref_gh = lin(lin(v, WE.double(), bE.double()), WQR[:, :128].double(), bQR.double())
    gh_rows = gh.view(Rp // 128, 8, 128, 16).permute(0, 2, 1, 3).reshape(Rp, 128)      # column-blocked -> rows
    e1, e2 = rel_err(lg[:R].cpu().numpy(), ref_lg.cpu().numpy()), rel_err(gh_rows[:R].cpu().numpy(), ref_gh.cpu().numpy())
    print(f"layer 10 scheme={scheme}: logit rel err {e1:.2e}, G h + g0 rel err {e2:.2e}")
    assert e1 < (2e-5 if scheme == 4 else 1e-4) and e2 < (2e-5 if scheme == 4 else 1e-4), (e1, e2)


@pytest.mark.parametrize("scheme", [0, 4], ids=["f16+f8", "f16x3"])
def test_gemm_tc_activation_scale_robustness(scheme):
    """The split-precision schemes must hold fp32-level accuracy over the dynamic range a trained checkpoint can produce, not
    only for O(1) activations: whole input scaled by 1e-3 and 1e3, and one matrix whose rows span 1e4 in magnitude. The error
    of a row is judged against that row's own output scale."""
    from coponerf_b200 import _lib
    lib, eng, Wm, bias = _tc_setup("query_encode_latent_2")
    N, K = Wm.shape
    torch.manual_seed(21)
    M = 512
    base = torch.randn(M, K, device="cuda")
    ramp = torch.logspace(-2, 2, M, device="cuda")[:, None]
    worst = {}
    for tag, A in (("x1", base), ("x1e-3", base * 1e-3), ("x1e3", base * 1e3), ("rows 1e-2..1e2", base * ramp),
                   ("relu x0.03", base.relu() * 0.03)):
        A = A.contiguous()
        C = torch.full((M, N), float("nan"), device="cuda")
        _lib.check(lib.cpn_gemm_tc(_p(eng.weights), 1, _p(A), K, _p(C), N, M, 0, scheme, 1, 1, _st()), "cpn_gemm_tc")
        ref = A.double() @ Wm.double().t()          # without the bias: it would mask the error of small rows
        got = C.double() - bias.double()
        err = (got - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)
        worst[tag] = float(err.max())
        print(f"gemm_tc scheme={scheme} {tag}: worst row rel err {worst[tag]:.2e}, median {float(err.median()):.2e}")
    for tag, e in worst.items():
        assert e < 6e-5, (tag, e, worst)      # measured / emulated: 2.5e-5 (f16+f8), 3e-5 at x1e-3 (f16x3: fp16 subnormal remainders)
    # beyond fp16's range the head saturates (tc::head2): one activation of 1e6 costs its own row its accuracy, but nothing
    # becomes inf - inf = NaN, and the other rows are untouched
    A = base.clone()
    A[7, 11] = 1e6
    C = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.cpn_gemm_tc(_p(eng.weights), 1, _p(A), K, _p(C), N, M, 0, scheme, 1, 1, _st()), "cpn_gemm_tc")
    assert torch.isfinite(C).all()
    ref = A.double() @ Wm.double().t() + bias.double()
    keep = torch.ones(M, dtype=torch.bool, device="cuda")
    keep[7] = False
    err = (C.double() - ref)[keep].abs().amax(dim=1) / ref[keep].abs().amax(dim=1)
    assert float(err.max()) < 6e-5, float(err.max())


def test_full_image_oblique_pose_against_oracle():
    """BASELINE config 2 size with the oblique pose (about a third of the rays miss both views, a tenth exactly one): the
    whole 65 536-ray image is rendered, and 4096 evenly strided rays of it are compared with the CPU oracle ray by ray under
    the same gates as the goldens (the oracle's own 1-ulp pose sensitivity widens the gate only where it is unstable)."""
    idx = torch.arange(0, 65536, 16)
    out = run_cuda(256, 256, None, 64, seed=2, val=True, pose="oblique")
    ref = run_oracle(256, 256, None, 64, seed=2, val=True, pose="oblique", chunk=512, with_sens=True, ray_idx=idx)
    v = ref["valid_mask"].numpy()
    assert 0.3 < v.mean() < 0.9                      # the degenerate-ray paths are exercised at full size
    cat_dim = {"pixel_val": 1, "at_wt": 1, "at_wt_max": 1, "coords": 1, "mask_c2": 1, "matchability_cycle_mask": 1, "rgb": 2}
    sub = {}
    for k, t in out.items():
        if isinstance(t, torch.Tensor) and k in ref and t.dim() > cat_dim.get(k, 1) and t.shape[cat_dim.get(k, 1)] == 65536:
            sub[k] = t.index_select(cat_dim.get(k, 1), idx)
        else:
            sub[k] = t
    tol = {k: v for k, v in GPU_TOL.items() if k != "C2_pts_to_C1"}
    check_against(sub, ref, "full-image-oblique", tol=tol, min_stable=0.4)
    # C2_pts_to_C1 = T_to_C2_pts + flow looked up at trunc(T_to_C2_pts) (utils.py:52-69): a ray whose projected point sits on
    # a pixel boundary reads a neighbouring flow texel; everywhere else the 1e-2 gate holds
    a, b = sub["C2_pts_to_C1"].numpy().astype(np.float64), ref["C2_pts_to_C1"].numpy().astype(np.float64)
    c2 = ref["T_to_C2_pts"].numpy().astype(np.float64)
    near = (np.abs(c2 - np.round(c2)) <= 2e-3 * np.maximum(1.0, np.abs(c2))).any(axis=-1)
    err = np.abs(a - b).max(axis=-1) / max(np.abs(b).max(), 1e-30)
    assert (err[~near] <= 1e-2).all(), float(err[~near].max())
    assert (~near).mean() > 0.1          # (points far outside the image count as 'near': the flow lookup clamps them)


@pytest.mark.parametrize("case", CASES[:3])
def test_compact_hidden_image_is_bit_identical_to_the_full_one(case):
    """Default path: the hidden-layer image is written without its value plane (3 bytes per element) and layer 10 derives
    e5m2(head) in shared memory; CPN_FLAG_FULL_H1 writes the plane. Same conversion of the same fp16 values: same bits."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, case + ".npz")))
    H, W, n_rays, S, seed, val = [int(v) for v in g["meta"]]
    a = run_cuda(H, W, n_rays, S, seed, val)
    b = run_cuda(H, W, n_rays, S, seed, val, flags=64)
    for k in EXACT_KEYS:
        assert torch.equal(a[k], b[k]), k
    check_against(b, g, case + "/full-h1")
