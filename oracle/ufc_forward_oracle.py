"""CPU restatement of the whole cost aggregation UFC.forward()  --  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
CPU legs). Nothing in coponerf_b200/ may import this file, and this file imports nothing from coponerf_b200/.

Written in the reference's own formulation (rearrange -> 2-D op -> rearrange, three materialised correlation volumes,
separable 4-D upsampling), as a function of the module's state_dict:
  UFCLayer.forward_attention / forward_cross / forward    models/aggregation.py:269-356
  UFC.forward                                             models/aggregation.py:509-562
  Mlp with DWConv                                         models/aggregation.py:18-29,184-189
  correlation / correlation_token / interpolate2d_token   models/aggregation.py:58-80
The 4-D convolution blocks, linear attention and the closing stage come from the other oracle files.

Pinned by tests/golden/pair_256.npz (outputs of the unmodified reference's get_z) through oracle/pair_oracle.py, and in the
build container directly against the reference module (tests/test_ufc_orchestration_cpu.py).
"""
import torch
import torch.nn.functional as F

from . import conv4d_oracle, ufc_oracle

NHEAD, DIM = 8, 32
CONV = ((3, 1, 1), (3, 2, 1), (5, 4, 2))      # (kernel, stride, padding) of embedding / feat_to_corr per level
LAYERS = (2, 2, 1)


def _encoder4d(sd, prefix, x, stride, pad):
    i = 0
    while f"{prefix}.conv4d.{i}.0.query_conv.weight" in sd:
        b = f"{prefix}.conv4d.{i}"
        p = dict(wq=sd[b + ".0.query_conv.weight"], bq=sd[b + ".0.query_conv.bias"], ws=sd[b + ".0.supp_conv.weight"],
                 bs=sd[b + ".0.supp_conv.bias"], gamma=sd[b + ".1.weight"], beta=sd[b + ".1.bias"])
        x = conv4d_oracle.encoder4d_layer(x.contiguous(), p, stride, pad)
        i += 1
    return x


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _mlp(sd, name, x, n):
    """nn.Sequential(Linear, DWConv, GELU, Linear), aggregation.py:184-189."""
    B, L, _ = x.shape
    h = _lin(sd, name + ".0", x)
    C = h.shape[-1]
    h = F.conv2d(h.transpose(1, 2).reshape(B, C, n, n), sd[name + ".1.dwconv.weight"], sd[name + ".1.dwconv.bias"], 1, 1, 1, C)
    h = F.gelu(h.flatten(2).transpose(1, 2))
    return _lin(sd, name + ".3", h)


def _corr_to_maps(corr, n):
    """'B H Hs Ws Ht Wt -> B (H Ht Wt) Hs Ws', bilinear (align_corners) to n x n."""
    B, H, hs, ws, ht, wt = corr.shape
    x = corr.permute(0, 1, 4, 5, 2, 3).reshape(B, H * ht * wt, hs, ws)
    return F.interpolate(x, size=(n, n), mode="bilinear", align_corners=True)


def _forward_attention(sd, p, corr, feat, n):
    B, H, hs, ws, ht, wt = corr.shape
    feat_r = feat
    feat = _ln(sd, p + ".norm1", feat)
    cf = torch.cat((_corr_to_maps(corr, n).flatten(2).transpose(1, 2), feat), dim=-1)
    pos = sd[p + ".pos_embed"]
    query = _lin(sd, p + ".q_proj", cf).view(B, -1, NHEAD, DIM) + pos
    key = _lin(sd, p + ".k_proj", cf).view(B, -1, NHEAD, DIM) + pos
    value_feat = _lin(sd, p + ".v_proj", feat).view(B, -1, NHEAD, DIM)
    value_corr = _corr_to_maps(_encoder4d(sd, p + ".v_proj_corr", corr, 1, 1), n)          # B (H Ht Wt) n n
    value_corr = value_corr.reshape(B, H, ht * wt, n * n).permute(0, 3, 1, 2)               # B (Hs Ws) H (Ht Wt)
    msg_feat = ufc_oracle.linear_attention(query, key, value_feat).reshape(B, -1, NHEAD * DIM)
    msg_corr = ufc_oracle.linear_attention(query, key, value_corr)                           # B (n n) H (Ht Wt)
    msg_corr = msg_corr.permute(0, 2, 3, 1).reshape(B, H * ht * wt, n, n)
    msg_corr = F.interpolate(msg_corr, size=(hs, ws), mode="bilinear", align_corners=True)
    msg_corr = msg_corr.reshape(B, H, ht, wt, hs, ws).permute(0, 1, 4, 5, 2, 3)
    msg_feat = feat_r + msg_feat
    msg_corr = corr + msg_corr
    msg_feat = msg_feat + _mlp(sd, p + ".mlp", _ln(sd, p + ".norm2", msg_feat), n)
    msg_corr = msg_corr + _encoder4d(sd, p + ".mlp_corr", msg_corr, 1, 1)
    return msg_corr, msg_feat


def _forward_cross(sd, p, corr, src_feat, trg_feat, n):
    B, H, hs, ws, ht, wt = corr.shape
    c2 = corr.reshape(B, H, hs * ws, ht * wt)

    def pooled(x, m):                      # 'B (H W) C -> B C H W', mean over (n/m) x (n/m) blocks, back to tokens
        C = x.shape[-1]
        y = x.transpose(1, 2).reshape(B, C, n, n)
        if n != m:
            y = F.avg_pool2d(y, n // m)
        return y.flatten(2).transpose(1, 2)

    def spread(x, m):                      # einops repeat 'B C H W -> B C (H P1) (W P2)'
        C = x.shape[-1]
        y = x.transpose(1, 2).reshape(B, C, m, m)
        if n != m:
            y = y.repeat_interleave(n // m, dim=2).repeat_interleave(n // m, dim=3)
        return y.flatten(2).transpose(1, 2)

    trg = _lin(sd, p + ".v_cross", _ln(sd, p + ".norm_cross1", pooled(trg_feat, ht))).view(B, -1, NHEAD, DIM)
    src = _lin(sd, p + ".v_cross", _ln(sd, p + ".norm_cross1", pooled(src_feat, hs))).view(B, -1, NHEAD, DIM)
    src_attn = torch.einsum("bhst,bthc->bshc", c2.softmax(-1), trg).reshape(B, -1, NHEAD * DIM)
    trg_attn = torch.einsum("bhst,bshc->bthc", c2.softmax(-2), src).reshape(B, -1, NHEAD * DIM)
    src_feat = src_feat + spread(src_attn, hs)
    trg_feat = trg_feat + spread(trg_attn, ht)
    src_feat = src_feat + _mlp(sd, p + ".mlp_cross", _ln(sd, p + ".norm_cross2", src_feat), n)
    trg_feat = trg_feat + _mlp(sd, p + ".mlp_cross", _ln(sd, p + ".norm_cross2", trg_feat), n)
    return src_feat, trg_feat


def _layer(sd, p, lvl, corr, src, trg, n):
    k, s, pad = CONV[lvl]
    corr_src, src_r = _forward_attention(sd, p, corr, src, n)
    corr_trg, trg_r = _forward_attention(sd, p, corr.permute(0, 1, 4, 5, 2, 3), trg, n)
    corr_r = corr_src + corr_trg.permute(0, 1, 4, 5, 2, 3)
    corr_r = corr_r + _encoder4d(sd, p + ".feat_to_corr1", ufc_oracle.correlation_token(src_r, trg_r, n), s, pad)
    corr_r = corr_r + _encoder4d(sd, p + ".mlp_refine_corr", corr_r, 1, 1)
    src_r, trg_r = _forward_cross(sd, p, corr_r, src_r, trg_r, n)
    corr_r = corr_r + _encoder4d(sd, p + ".feat_to_corr2", ufc_oracle.correlation_token(src_r, trg_r, n), s, pad)
    corr_r = corr_r + _encoder4d(sd, p + ".mlp_refine_corr2", corr_r, 1, 1)
    return corr_r, src_r, trg_r


def _upsample_tokens(x, n_out):
    """interpolate2d_token, aggregation.py:58-63."""
    B, L, C = x.shape
    n = int(round(L ** 0.5))
    y = F.interpolate(x.transpose(1, 2).reshape(B, C, n, n), size=(n_out, n_out), mode="bilinear", align_corners=True)
    return y.flatten(2).transpose(1, 2)


@torch.no_grad()
def ufc_forward(sd, feat, nview=2):
    """sd: state_dict of the UFC module (keys without the 'feature_cost_aggregation.' prefix); feat: the encoder pyramid
    [(2B, 512, 16, 16), (2B, 256, 32, 32), (2B, 128, 64, 64)]. Returns (feat_list, flows, c) like UFC.forward."""
    B2 = feat[0].shape[0]
    sizes = [f.shape[-1] for f in feat]

    def side(i, v):
        x = feat[i].view(B2 // nview, nview, -1, sizes[i], sizes[i])[:, v].flatten(2).transpose(1, 2)
        return F.relu(_lin(sd, f"proj_feat.{i}.0", x))

    src = [side(i, 0) for i in range(3)]
    trg = [side(i, 1) for i in range(3)]
    feat_list, refined = [], []
    corr = s = t = None
    for lvl, n in enumerate(sizes):
        k, stride, pad = CONV[lvl]
        raw = _encoder4d(sd, f"embedding.{lvl}", ufc_oracle.correlation_token(src[lvl], trg[lvl], n), stride, pad)
        corr = raw if lvl == 0 else corr + raw
        s = src[lvl] if lvl == 0 else _upsample_tokens(s, n) + src[lvl]
        t = trg[lvl] if lvl == 0 else _upsample_tokens(t, n) + trg[lvl]
        for j in range(LAYERS[lvl]):
            corr, s, t = _layer(sd, f"layers.{lvl}.{j}", lvl, corr, s, t, n)
        both = torch.stack((s, t), dim=1).flatten(0, 1)
        feat_list.append(both.transpose(1, 2).reshape(both.shape[0], both.shape[2], n, n))
        refined.append((s, t))
    flows, c = ufc_oracle.ufc_tail([r[0] for r in refined], [r[1] for r in refined], tuple(sizes), sizes[-1])
    return feat_list, flows, c
