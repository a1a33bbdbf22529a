"""Host side of the per-pair cost-aggregation kernels (models/aggregation.py). Built so far: the closing stage of
UFC.forward() -- correlation of the refined features of the three levels, 4-D upsampling, their mean `c`, the two
soft-argmax flow fields and the conversion to pixel flow (aggregation.py:527,539,549-561) -- as `ufc_tail`."""
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


# ---------------------------------------------------------------------------------------------------------------
# UFC.forward with the native closing stage. The coarse-to-fine refinement (proj_feat, embedding, the five UFCLayer
# blocks: aggregation.py:509-549) is not native yet and is delegated to the attached reference module's own
# submodules; what is replaced is everything after the last UFCLayer.
def _tokens_to_map(x, n):
    return x.transpose(1, 2).reshape(x.shape[0], x.shape[2], n, n)


def _correlation(src_tok, trg_tok, n, eps=1e-5):
    """aggregation.py:70-74 on token features."""
    s, t = _tokens_to_map(src_tok, n), _tokens_to_map(trg_tok, n)
    s = s / (s.norm(dim=1, p=2, keepdim=True) + eps)
    t = t / (t.norm(dim=1, p=2, keepdim=True) + eps)
    return torch.einsum("bchw,bcxy->bhwxy", s, t)[:, None]


def _upsample_tokens(x, n_out):
    """aggregation.py:58-63 (interpolate2d_token)."""
    n = int(round(x.shape[1] ** 0.5))
    y = torch.nn.functional.interpolate(_tokens_to_map(x, n), size=(n_out, n_out), mode="bilinear", align_corners=True)
    return y.flatten(2).transpose(1, 2)


def _encoder4d_forward(enc, block_fn):
    """forward() for a reference Encoder4D module (conv4d.py:156-163) that runs every
    Conv4d -> GroupNorm -> ReLU block through `block_fn` (conv4d_block on CUDA)."""
    def forward(x):
        for blk in enc.conv4d:
            c4, gn = blk[0], blk[1]
            x = block_fn(x, c4.query_conv.weight, c4.query_conv.bias, c4.supp_conv.weight, c4.supp_conv.bias,
                         gn.weight, gn.bias, c4.stride[0], c4.padding[0])
        return x
    return forward


class _patched_modules:
    """Context manager: for the duration of a call, route every Encoder4D inside `root` through `block_fn` and
    every LinearAttention through `attention_fn`."""

    def __init__(self, root, block_fn, attention_fn):
        self.enc = [m for m in root.modules() if type(m).__name__ == "Encoder4D"]
        self.att = [m for m in root.modules() if type(m).__name__ == "LinearAttention"]
        self.block_fn, self.attention_fn = block_fn, attention_fn

    def __enter__(self):
        for m in self.enc:
            m.forward = _encoder4d_forward(m, self.block_fn)
        for m in self.att:
            m.forward = lambda q, k, v, q_mask=None, kv_mask=None, _f=self.attention_fn: _f(q, k, v)

    def __exit__(self, *exc):
        for m in self.enc + self.att:
            del m.forward          # back to the class's own forward
        return False


def ufc_forward(fca, feat, nview, tail=None, conv_block=None, attention=None):
    """Drop-in for UFC.forward(feat, nview) (aggregation.py:509-562) of the attached reference module `fca`.

    Returns (feat_list, (flow, flow_flip, flow_t_to_s, flow_s_to_t), c) like the reference. Native so far: every
    Encoder4D block (63 Conv4d + GroupNorm + ReLU per pair, `conv_block`, default conv4d_block), every
    LinearAttention (20 per pair, `attention`, default linear_attention) and the closing stage (`tail`, default
    ufc_tail). Tests pass the CPU oracles for all three to check the orchestration without a GPU.
    """
    tail = tail or ufc_tail
    with _patched_modules(fca, conv_block or conv4d_block, attention or linear_attention):
        return _ufc_forward(fca, feat, nview, tail)


def _ufc_forward(fca, feat, nview, tail):
    B = feat[0].shape[0]
    sizes = [f.shape[-1] for f in feat]

    def side(i, v):
        x = feat[i].view(B // nview, nview, -1, sizes[i], sizes[i])[:, v]
        return fca.proj_feat[i](x.flatten(2).transpose(1, 2))

    src = [side(i, 0) for i in range(3)]
    trg = [side(i, 1) for i in range(3)]
    feat_list, refined = [], []
    corr, s, t = None, None, None
    for lvl in range(3):
        raw = fca.embedding[lvl](_correlation(src[lvl], trg[lvl], sizes[lvl]))
        corr = raw if lvl == 0 else corr + raw
        s = src[lvl] if lvl == 0 else _upsample_tokens(s, sizes[lvl]) + src[lvl]
        t = trg[lvl] if lvl == 0 else _upsample_tokens(t, sizes[lvl]) + trg[lvl]
        corr, s, t = fca.layers[lvl](corr, s, t)
        both = torch.stack((s, t), dim=1).flatten(0, 1)
        feat_list.append(_tokens_to_map(both, sizes[lvl]))
        refined.append((s, t))
    flows, c = tail([r[0] for r in refined], [r[1] for r in refined], tuple(sizes), sizes[-1])
    return feat_list, flows, c
