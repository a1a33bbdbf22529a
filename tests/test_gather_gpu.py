"""The epipolar feature gather alone (cpn_gather_rows: the fp32 kernel and the operand-image kernel) against
F.grid_sample, the call it replaces (models/CoPoNeRF.py:312 'border', :370 'zeros'; bilinear, align_corners=False),
including coordinates outside [-1, 1], exactly on the border, and huge ones."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

LEVELS = [(256, 16), (256, 32), (256, 64), (64, 256)]      # (channels, size) of the four maps at 256 x 256


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _setup(B, nr, S, seed, levels=LEVELS):
    from coponerf_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    maps = [torch.randn(2 * B, c, n, n, generator=g).cuda() for c, n in levels]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    a = _lib.RenderArgs()
    a.B, a.S = B, S
    keep = []
    for l, m in enumerate(maps):
        n_img, c, h, w = m.shape
        cl = torch.empty((n_img, h, w, c), device="cuda")
        _lib.check(lib.cpn_pack_features(_p(m), _p(cl), n_img, c, h, w, st), "cpn_pack_features")
        keep.append(cl)
        a.feat[l] = cl.data_ptr()
        a.feat_h[l], a.feat_w[l], a.feat_c[l] = h, w, c
    R = B * nr * 2 * S
    # coordinates: mostly inside, some outside [-1, 1] (border clamps / zeros pads), some exactly on texel centres and
    # image edges, a few far away
    xy = torch.rand(R, 4, generator=g) * 2.6 - 1.3
    xy[::17] = torch.tensor([-1.0, 1.0, 1.0, -1.0])
    xy[5::29] = torch.tensor([1.0 - 1.0 / 16, -1.0 + 1.0 / 16, 0.0, 0.0])
    xy[3::41] = torch.tensor([7.5, -3.0, -1e4, 2e4])
    rowaux = torch.zeros(R, 8)
    rowaux[:, :4] = xy
    return lib, a, maps, keep, rowaux.cuda(), st, R


def _reference(maps, rowaux, B, nr, S):
    """(R, 2 branches, 832): primary = view v with 'border', secondary = view 1 - v with 'zeros'."""
    R = rowaux.shape[0]
    grid = rowaux[:, :4].reshape(B, nr, 2, S, 4).permute(0, 2, 1, 3, 4).reshape(B * 2, nr, S, 4)   # (b v) n s
    outs = []
    for br, pad in ((0, "border"), (1, "zeros")):
        cols = []
        for m in maps:
            src = m if br == 0 else m.view(B, 2, *m.shape[1:]).flip(1).reshape(m.shape)
            cols.append(F.grid_sample(src.double(), grid[..., 2 * br:2 * br + 2].double(), mode="bilinear", padding_mode=pad,
                                      align_corners=False))
        o = torch.cat(cols, dim=1)                                                        # (b v) c n s
        outs.append(o.view(B, 2, 832, nr, S).permute(0, 3, 1, 4, 2).reshape(R, 832))      # row = ((b n) v) s
    return torch.stack(outs, dim=1)


@pytest.mark.parametrize("B,nr,S", [(1, 5, 64), (2, 3, 32)])
def test_fp32_gather_matches_grid_sample(B, nr, S):
    from coponerf_b200 import _lib
    lib, a, maps, keep, rowaux, st, R = _setup(B, nr, S, seed=0)
    Rp = (R + 127) // 128 * 128
    out = torch.zeros(2 * Rp, 848, device="cuda")
    _lib.check(lib.cpn_gather_rows(ctypes.byref(a), nr, _p(rowaux), _p(out), 0, None, st), "cpn_gather_rows")
    ref = _reference(maps, rowaux, B, nr, S)
    rows = torch.arange(R, device="cuda")
    for br in range(2):
        got = out[(rows // 128) * 256 + br * 128 + rows % 128, :832]
        err = float((got.double() - ref[:, br]).abs().max() / ref[:, br].abs().max())
        print(f"fp32 gather branch {br}: rel err {err:.2e}")
        assert err <= 1e-6, (br, err)
    zero_rows = (rowaux[:, 2].abs() > 1.5) | (rowaux[:, 3].abs() > 1.5)
    assert zero_rows.any() and (out[(rows // 128) * 256 + 128 + rows % 128][zero_rows][:, :832] == 0).all()   # 'zeros' padding


def _decode_image(img, Rp, kchunks, f8):
    """Operand image -> (tiles * 128 rows, kchunks * 32) fp64 values (hi + correction) and, for f8, the e4m3 value plane."""
    blocks = img.view(-1, kchunks, 16384)
    nt = blocks.shape[0]
    hi = blocks[:, :, :8192].contiguous().view(torch.float16).view(nt, kchunks, 4, 128, 8).double()
    val = hi.permute(0, 3, 1, 2, 4).reshape(nt * 128, kchunks * 32)
    if f8:
        # tc_common.cuh: remainder plane = e5m2((x - hi) * 2^10), value plane = e5m2(hi)
        lo = blocks[:, :, 8192:12288].contiguous().view(torch.float8_e5m2).float().view(nt, kchunks, 2, 128, 16).double() / 1024.0
        x8 = blocks[:, :, 12288:].contiguous().view(torch.float8_e5m2).float().view(nt, kchunks, 2, 128, 16).double()
        return val + lo.permute(0, 3, 1, 2, 4).reshape(nt * 128, kchunks * 32), x8.permute(0, 3, 1, 2, 4).reshape(nt * 128, kchunks * 32)
    lo = blocks[:, :, 8192:].contiguous().view(torch.float16).view(nt, kchunks, 4, 128, 8).double()
    return val + lo.permute(0, 3, 1, 2, 4).reshape(nt * 128, kchunks * 32), None


@pytest.mark.parametrize("form", [2, 1], ids=["f16+f8", "f16x3"])
@pytest.mark.parametrize("B,nr,S", [(1, 5, 64), (2, 3, 32)])
def test_operand_image_gather_matches_grid_sample(B, nr, S, form):
    from coponerf_b200 import _lib
    lib, a, maps, keep, rowaux, st, R = _setup(B, nr, S, seed=1)
    Rp = (R + 127) // 128 * 128
    img = torch.zeros(2 * Rp // 128 * 27 * 16384, dtype=torch.uint8, device="cuda")
    taps = torch.empty(lib.cpn_gather_rows_taps_bytes(R), dtype=torch.uint8, device="cuda")
    _lib.check(lib.cpn_gather_rows(ctypes.byref(a), nr, _p(rowaux), _p(img), form, _p(taps), st), "cpn_gather_rows")
    val, x8 = _decode_image(img, Rp, 27, form == 2)
    ref = _reference(maps, rowaux, B, nr, S)
    rows = torch.arange(R, device="cuda")
    for br in range(2):
        got = val[(rows // 128) * 256 + br * 128 + rows % 128, :832]
        d = (got - ref[:, br]).abs()
        # f16+f8: |x - hi| <= 2^-11 |x| kept to 3 significant bits (e5m2) -> 2^-14 |x|;
        # f16x3: two fp16 -> 2^-22 |x|; the blend itself is fp32 (1e-7)
        tol = (2.0 ** -13.9 if form == 2 else 2.0 ** -21) * ref[:, br].abs() + 1e-6
        assert (d <= tol).all(), (br, float((d / (ref[:, br].abs() + 1e-3)).max()))
        if x8 is not None:        # the e5m2 copy of the value: 3 significant bits of x
            g8 = x8[(rows // 128) * 256 + br * 128 + rows % 128, :832]
            assert ((g8 - ref[:, br]).abs() <= 2.0 ** -2.9 * ref[:, br].abs() + 2e-5).all()


@pytest.mark.parametrize("B,nr,S", [(1, 5, 64), (2, 3, 32)])
def test_compact_operand_image_equals_full_image_without_value_plane(B, nr, S):
    """form 3 (12 KB blocks, the encoder GEMM derives the value plane on chip) holds the same fp16 heads and remainder
       bytes as form 2, block by block (the sample kernel writes k-chunk 26 of both forms; here it is left zero in both)."""
    from coponerf_b200 import _lib
    lib, a, maps, keep, rowaux, st, R = _setup(B, nr, S, seed=3)
    Rp = (R + 127) // 128 * 128
    nblk = 2 * Rp // 128 * 27
    full = torch.zeros(nblk * 16384, dtype=torch.uint8, device="cuda")
    compact = torch.zeros(nblk * 12288, dtype=torch.uint8, device="cuda")
    taps = torch.empty(lib.cpn_gather_rows_taps_bytes(R), dtype=torch.uint8, device="cuda")
    _lib.check(lib.cpn_gather_rows(ctypes.byref(a), nr, _p(rowaux), _p(full), 2, _p(taps), st), "cpn_gather_rows")
    _lib.check(lib.cpn_gather_rows(ctypes.byref(a), nr, _p(rowaux), _p(compact), 3, _p(taps), st), "cpn_gather_rows")
    torch.cuda.synchronize()
    assert torch.equal(full.view(nblk, 16384)[:, :12288], compact.view(nblk, 12288))
    assert full.view(nblk, 16384)[:, :8192].any()
