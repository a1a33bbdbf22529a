"""Thin host wrappers of three cost-aggregation entry points (models/aggregation.py, models/conv4d.py) used by the operator
tests and benches: the closing stage of UFC.forward() (`ufc_tail`: aggregation.py:527,539,549-561), one Encoder4D block
(`conv4d_block`: conv4d.py:149-153) and LinearAttention (`linear_attention`: aggregation.py:84-117). The whole
UFC.forward runs through coponerf_b200/ufc_native.py over coponerf_b200/ufc_ops.py::CudaOps."""
import ctypes

import torch

from . import _lib


_LIN = {}


def _linspace(n, dev):
    """linspace(-1, 1, n), built on the host like the reference's and cached per device (graph-capture safe)."""
    key = (n, str(dev))
    if key not in _LIN:
        _LIN[key] = torch.linspace(-1, 1, n).to(dev)
    return _LIN[key]


def ufc_tail(src_feats, trg_feats, sizes=(16, 32, 64), out=64):
    """src_feats / trg_feats: three CUDA fp32 token tensors (B, sizes[l]^2, C).

    Returns ((flow, flow_flip, flow_t_to_s, flow_s_to_t), c) with the shapes UFC.forward returns:
    flows (B, 2, out, out), c (B, 1, out, out, out, out).
    """
    lib = _lib.load()
    dev = src_feats[0].device
    if dev.type != "cuda":
        raise _lib.CpnError("ufc_tail runs on CUDA only (no CPU fallback)")
    B, _, C = src_feats[0].shape
    src = [t.detach().to(torch.float32).contiguous() for t in src_feats]
    trg = [t.detach().to(torch.float32).contiguous() for t in trg_feats]
    for l, n in enumerate(sizes):
        if tuple(src[l].shape) != (B, n * n, C) or tuple(trg[l].shape) != (B, n * n, C):
            raise ValueError(f"level {l}: expected (B, {n * n}, {C}) token features")
    with torch.cuda.device(dev):
        f32 = dict(dtype=torch.float32, device=dev)
        c = torch.empty((B, 1, out, out, out, out), **f32)
        flows = [torch.empty((B, 2, out, out), **f32) for _ in range(4)]
        lin = _linspace(out, dev)
        a = _lib.UfcTailArgs()
        a.B, a.C, a.out = B, C, out
        for l in range(3):
            a.sizes[l] = sizes[l]
            a.src[l] = src[l].data_ptr()
            a.trg[l] = trg[l].data_ptr()
        a.lin, a.c = lin.data_ptr(), c.data_ptr()
        a.flow, a.flow_flip, a.flow_t_to_s, a.flow_s_to_t = (f.data_ptr() for f in flows)
        nbytes = lib.cpn_ufc_tail_workspace_bytes(B, C, out, a.sizes)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
        _lib.check(lib.cpn_ufc_tail(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "cpn_ufc_tail")
    return tuple(flows), c


def conv4d_block(x, wq, bq, ws, bs, gamma=None, beta=None, stride=1, pad=1):
    """One Encoder4D block (models/conv4d.py:149-153): Conv4d -> GroupNorm(1 group) -> ReLU on CUDA; without
    gamma / beta the plain Conv4d. x (B, Ci, Hq, Hq, Hs, Hs) -> (B, Co, oq, oq, os, os)."""
    lib = _lib.load()
    dev = x.device
    if dev.type != "cuda":
        raise _lib.CpnError("conv4d_block runs on CUDA only (no CPU fallback)")
    f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    x, wq, bq, ws, bs = f(x), f(wq), f(bq), f(ws), f(bs)
    B, Ci, Hq, Wq, Hs, Ws = x.shape
    Co, _, k, _ = wq.shape
    if Hq != Wq or Hs != Ws or tuple(ws.shape) != tuple(wq.shape) or wq.shape[1] != Ci:
        raise ValueError("conv4d_block: square query / support axes and matching (Co, Ci, k, k) weights expected")
    oq, os_ = (Hq + 2 * pad - k) // stride + 1, (Hs + 2 * pad - k) // stride + 1
    with torch.cuda.device(dev):
        y = torch.empty((B, Co, oq, oq, os_, os_), dtype=torch.float32, device=dev)
        a = _lib.Conv4dArgs()
        a.B, a.Ci, a.Co, a.Hq, a.Hs, a.k, a.stride, a.pad = B, Ci, Co, Hq, Hs, k, stride, pad
        a.norm_relu = int(gamma is not None)
        keep = [x, wq, bq, ws, bs]
        a.x, a.wq, a.bq, a.ws, a.bs, a.y = (t.data_ptr() for t in (x, wq, bq, ws, bs, y))
        if gamma is not None:
            g, bt = f(gamma), f(beta)
            keep += [g, bt]
            a.gamma, a.beta = g.data_ptr(), bt.data_ptr()
        nbytes = lib.cpn_conv4d_workspace_bytes(B, Hq, Hs, k, stride, pad)
        wsb = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        a.workspace, a.workspace_bytes = wsb.data_ptr(), nbytes
        _lib.check(lib.cpn_conv4d(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "cpn_conv4d")
    return y


def linear_attention(queries, keys, values):
    """LinearAttention.forward (models/aggregation.py:84-117) on CUDA: (N, L, H, 32), (N, S, H, 32), (N, S, H, Dv)."""
    lib = _lib.load()
    dev = queries.device
    if dev.type != "cuda":
        raise _lib.CpnError("linear_attention runs on CUDA only (no CPU fallback)")
    f = lambda t: t.detach().to(torch.float32).contiguous()
    q, k, v = f(queries), f(keys), f(values)
    N, L, H, D = q.shape
    S, Dv = k.shape[1], v.shape[-1]
    with torch.cuda.device(dev):
        out = torch.empty((N, L, H, Dv), dtype=torch.float32, device=dev)
        nbytes = lib.cpn_linear_attention_workspace_bytes(N, H, Dv)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(lib.cpn_linear_attention(p(q), p(k), p(v), N, L, S, H, D, Dv, p(out), p(ws), nbytes,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "cpn_linear_attention")
    return out
