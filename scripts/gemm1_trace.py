"""Per-tile phase trace of a tensor-core GEMM at bench size (cpn_gemm_tc_trace): where an SM's time goes.

    python scripts/gemm1_trace.py [rows] [layer: 0 | 8 | 10] [persist: 1 | 0 | 2 | 3] [compact: 0 | 1]

CPN_TC_DBG_SKIP (bit set, trace runs only): 1 no image stores, 2 no split, 4 no TMEM loads, 8 no MMAs, 16 no activation copies,
32 no weight copies -- the attribution runs behind profiles/r2_gemm1_mainloop_attribution.log.
"""
import ctypes, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import cuda_model
from coponerf_b200 import _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 524288       # encoder rows (layer 0); layers 8 / 10 see M / 2 sample rows
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 0
persist = int(sys.argv[3]) if len(sys.argv) > 3 else 1     # 1 persistent, 0 one tile per CTA, 2 cta_group::2 pairs (one tile per pair), 3 persistent pairs
lib = _lib.load(); eng = cuda_model().engine()
tiles = M // 128
kin = {0: 27, 8: 52, 10: 52}[layer]
A = torch.zeros((tiles // (2 if layer else 1)) * kin * 16384, dtype=torch.uint8, device="cuda")
Av = A.view(-1, 16384)
Av[:, :8192] = (torch.rand(Av.shape[0], 4096, device="cuda") * 4 - 2).half().view(torch.uint8)
C = torch.empty(tiles * 26 * 16384, dtype=torch.uint8, device="cuda")
dv = torch.randn(M // 2 * 256, device="cuda")
gh = torch.empty(M // 2 * 128, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr()); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rows = M if layer == 0 else M // 2
ntn = 4 if layer == 0 else 1
ntiles = ntn * (rows // 256)
extra = {1: 0, 0: _lib.TC_NO_PERSIST, 2: _lib.TC_PAIR, 3: _lib.TC_PPAIR}[persist]
compact = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if compact:
    extra |= _lib.TC_A_IMAGE3 | (_lib.TC_OUT_IMAGE3 if layer == 0 else 0)
p2 = persist == 1 and os.environ.get("CPN_TC_PERSIST2", "0") not in ("", "0") and os.environ.get("CPN_TC_EPI_WARPS", "16") != "24"
nslots = ntiles * (2 if p2 else 1)          # the sub-tile pipelined kernel stamps per (tile, sub-tile)
buf = torch.zeros(nslots * 8, dtype=torch.int64, device="cuda")
def run():
    if layer == 0:
        _lib.check(lib.cpn_gemm_tc(p(eng.weights), 0, p(A), 0, p(C), 0, M, 1, _lib.TC_A_IMAGE | _lib.TC_OUT_IMAGE | extra, 1, 26, st), "gemm_tc")
    elif layer == 10:
        _lib.check(lib.cpn_gemm_tc_kg(p(eng.weights), p(A), p(dv), 16, None, 11.31, p(C), p(gh), rows, extra, st), "gemm_tc_kg")
    else:
        _lib.check(lib.cpn_gemm_tc(p(eng.weights), layer, p(A), 0, p(C), 128, rows, 1, _lib.TC_A_IMAGE | extra, 1, 1, st), "gemm_tc")
run(); run(); torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); run(); ev[1].record(); torch.cuda.synchronize()
lib.cpn_gemm_tc_trace(p(buf), nslots)
run(); torch.cuda.synchronize()
lib.cpn_gemm_tc_trace(None, 0)
t = buf.cpu().numpy().reshape(nslots, 8).astype(np.float64)
st_ = lambda v: {"mean": float(v.mean() / 1e3), "p10": float(np.quantile(v, .1) / 1e3), "p50": float(np.median(v) / 1e3), "p90": float(np.quantile(v, .9) / 1e3)}
out = {"layer": layer, "rows": rows, "tiles": ntiles, "persistent": bool(persist), "untraced_launch_ms": ev[0].elapsed_time(ev[1])}
if persist == 3:
    nct = 148
    per = (ntiles // 2) // 74                  # complete rounds of 512-row pair tiles per cluster
    tt = t[:per * nct].reshape(per, nct, 8)
    lead, peer = tt[:, 0::2], tt[:, 1::2]
    out["kernel_span_us"] = float((t[:, 4].max() - lead[0, :, 0].min()) / 1e3)
    out["ring_wait_at_tile_start_us"] = st_(lead[:, :, 1] - lead[:, :, 0])
    out["mainloop_issue_us"] = st_(lead[:, :, 2] - lead[:, :, 1])
    out["accum_ready_after_last_issue_leader_us"] = st_(lead[:, :, 3] - lead[:, :, 2])
    out["accum_ready_after_last_issue_peer_us"] = st_(peer[:, :, 3] - lead[:, :, 2])
    out["drain_leader_us"] = st_(lead[:, :, 4] - lead[:, :, 3])
    out["drain_peer_us"] = st_(peer[:, :, 4] - peer[:, :, 3])
    out["tile_period_us"] = st_(lead[1:, :, 0] - lead[:-1, :, 0])
    out["restart_after_last_drain_us"] = st_(lead[1:, :, 0] - np.maximum(lead[:-1, :, 4], peer[:-1, :, 4]))
elif p2:
    ncta = min(148, ntiles)
    per = ntiles // ncta                      # complete rounds of tiles
    tt = t[:per * 2 * ncta].reshape(per, 2, ncta, 8)    # [tile round][sub-tile][CTA]
    seq = tt.transpose(0, 2, 1, 3).transpose(1, 0, 2, 3).reshape(ncta, per * 2, 8)   # per CTA: sub-tiles in issue order
    out["kernel_span_us"] = float((t[:, 4].max() - tt[0, 0, :, 0].min()) / 1e3)
    out["ring_wait_at_subtile_start_us"] = st_(seq[:, :, 1] - seq[:, :, 0])
    out["mainloop_issue_us"] = st_(seq[:, :, 2] - seq[:, :, 1])
    out["accum_ready_after_last_issue_us"] = st_(seq[:, :, 3] - seq[:, :, 2])
    out["drain_us"] = st_(seq[:, :, 4] - seq[:, :, 3])
    out["subtile_period_us"] = st_(seq[:, 1:, 0] - seq[:, :-1, 0])
    out["tile_period_us"] = st_(seq[:, 2:, 0] - seq[:, :-2, 0])
    out["mma_idle_between_subtiles_us"] = st_(seq[:, 1:, 0] - seq[:, :-1, 2])
    out["wait_for_accumulator_slot_us"] = st_(np.maximum(0, seq[:, 2:, 0] - np.maximum(seq[:, :-2, 4], seq[:, 1:-1, 2])))
elif persist == 1:
    ncta = min(148, ntiles)
    per = ntiles // ncta                      # complete rounds of tiles
    tt = t[:per * ncta].reshape(per, ncta, 8)
    out["kernel_span_us"] = float((t[:, 4].max() - tt[0, :, 0].min()) / 1e3)
    out["ring_wait_at_tile_start_us"] = st_(tt[:, :, 1] - tt[:, :, 0])
    out["mainloop_issue_us"] = st_(tt[:, :, 2] - tt[:, :, 1])
    out["accum_ready_after_last_issue_us"] = st_(tt[:, :, 3] - tt[:, :, 2])
    out["drain_us"] = st_(tt[:, :, 4] - tt[:, :, 3])
    out["tile_period_us"] = st_(tt[1:, :, 0] - tt[:-1, :, 0])
    out["mma_idle_between_tiles_us"] = st_(tt[1:, :, 0] - tt[:-1, :, 2])
else:
    t0 = t[:, 0].min()
    out["kernel_span_us"] = float((t[:, 6].max() - t0) / 1e3)
    for k, (a, b) in {"setup": (0, 1), "first_stage_wait": (1, 2), "mainloop_issue": (2, 3), "accum_ready_after_last_issue": (3, 4),
                      "drain": (4, 5), "exit": (5, 6), "cta_total": (0, 6)}.items():
        out[k + "_us"] = st_(t[:, b] - t[:, a])
    sm = t[:, 7].astype(int)
    gaps = []
    for s in np.unique(sm):
        a = t[sm == s]; a = a[np.argsort(a[:, 0])]
        gaps += list(a[1:, 0] - a[:-1, 6])
    out["gap_between_ctas_us"] = st_(np.array(gaps))
print(json.dumps(out, indent=1))
