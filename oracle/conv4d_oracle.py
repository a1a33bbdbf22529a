"""CPU oracle for the centre-pivot 4-D convolution block  --  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/models/conv4d.py:7-32 (MaxPool4d), :57-135 (Conv4d) and :138-163 (Encoder4D:
Conv4d -> GroupNorm(1 group) -> ReLU) in plain PyTorch, without einops. Pinned against the reference modules by
tests/golden/make_goldens_ufc.py -> tests/golden/conv4d_*.npz. Nothing in coponerf_b200/ may import this file.
"""
import torch
import torch.nn.functional as F


def _pool(x, s, dims):
    """MaxPool4d(kernel = stride = s, ceil_mode=True) over the support (last two) or query (dims 2, 3) axes."""
    B, L, hq, wq, hs, ws = x.shape
    if dims == "support":
        y = F.max_pool2d(x.reshape(B * L * hq * wq, 1, hs, ws), s, s, 0, ceil_mode=True)
        return y.reshape(B, L, hq, wq, y.shape[-2], y.shape[-1])
    y = x.permute(0, 1, 4, 5, 2, 3).reshape(B * L * hs * ws, 1, hq, wq)
    y = F.max_pool2d(y, s, s, 0, ceil_mode=True)
    return y.reshape(B, L, hs, ws, y.shape[-2], y.shape[-1]).permute(0, 1, 4, 5, 2, 3)


def conv4d(x, wq, bq, ws_, bs, stride, pad):
    """conv4d.py:108-135: query_conv over (Hq, Wq) of the support-pooled input + supp_conv over (Hs, Ws) of the
    query-pooled input. x (B, Ci, Hq, Wq, Hs, Ws); weights (Co, Ci, k, k)."""
    B, Ci, hq, wq_, hs, ws = x.shape
    xq = _pool(x, stride, "support") if stride > 1 else x
    xs = _pool(x, stride, "query") if stride > 1 else x
    hs2, ws2 = xq.shape[-2:]
    a = xq.permute(0, 4, 5, 1, 2, 3).reshape(B * hs2 * ws2, Ci, hq, wq_)
    a = F.conv2d(a, wq, bq, stride=stride, padding=pad)
    Co, hq2, wq2 = a.shape[1:]
    a = a.reshape(B, hs2, ws2, Co, hq2, wq2).permute(0, 3, 4, 5, 1, 2)
    hq3, wq3 = xs.shape[2:4]
    b = xs.permute(0, 2, 3, 1, 4, 5).reshape(B * hq3 * wq3, Ci, hs, ws)
    b = F.conv2d(b, ws_, bs, stride=stride, padding=pad)
    b = b.reshape(B, hq3, wq3, Co, b.shape[-2], b.shape[-1]).permute(0, 3, 1, 2, 4, 5)
    return a + b


def encoder4d_layer(x, p, stride, pad, eps=1e-5):
    """One [Conv4d -> GroupNorm(1, Co) -> ReLU] block (conv4d.py:149-153). p: dict with wq, bq, ws, bs, gamma, beta."""
    y = conv4d(x, p["wq"], p["bq"], p["ws"], p["bs"], stride, pad)
    y = F.group_norm(y, 1, p["gamma"], p["beta"], eps)
    return F.relu(y)
