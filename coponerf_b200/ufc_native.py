"""UFC.forward (models/aggregation.py:509-562) re-stated as a function of the module's state_dict over a small set
of operators, so that the whole per-pair cost aggregation runs without the reference's Python modules.

`ops` supplies the operators. The product backend is `coponerf_b200.ufc_ops.CudaOps` (sm_100a kernels behind the
C-ABI); tests check this orchestration on CPU against the unmodified reference with the PyTorch restatement of the
same operators (oracle/ufc_ops_torch.py). Tensors are (B, L, C) tokens and (B, H, hs, ws, ht, wt) correlation volumes.
"""
import torch

# Encoder4D geometry (kernel, stride, padding) of embedding[l] / feat_to_corr{1,2} per level (aggregation.py:363-470)
LEVEL_CONV = ((3, 1, 1), (3, 2, 1), (5, 4, 2))
LAYER_NUMS = (2, 2, 1)
NHEAD, HEAD_DIM = 8, 32


def _enc_params(sd, prefix):
    """Blocks of an Encoder4D at `prefix` (conv4d.i.0 = Conv4d, conv4d.i.1 = GroupNorm)."""
    blocks, i = [], 0
    while f"{prefix}.conv4d.{i}.0.query_conv.weight" in sd:
        b = f"{prefix}.conv4d.{i}"
        blocks.append(dict(wq=sd[b + ".0.query_conv.weight"], bq=sd[b + ".0.query_conv.bias"],
                           ws=sd[b + ".0.supp_conv.weight"], bs=sd[b + ".0.supp_conv.bias"],
                           gamma=sd[b + ".1.weight"], beta=sd[b + ".1.bias"]))
        i += 1
    return blocks


def _cat_cached(sd, key, names):
    """Concatenation of parameters along dim 0, made once per parameter view `sd` (rebuilt when the weights change), so that
    the operator set sees one stable tensor and packs it once."""
    if key not in sd:
        sd[key] = torch.cat([sd[n] for n in names], dim=0)
    return sd[key]


def _mlp(ops, sd, p, x, n):
    """Linear -> DWConv 3x3 -> GELU -> Linear (aggregation.py:184-189)."""
    h = ops.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
    h = ops.dwconv_gelu(h, sd[p + ".1.dwconv.weight"], sd[p + ".1.dwconv.bias"], n)
    return ops.linear(h, sd[p + ".3.weight"], sd[p + ".3.bias"])


def _forward_attention(ops, sd, p, corr, feat, n):
    """UFCLayer.forward_attention (aggregation.py:269-310)."""
    B, L, C = feat.shape
    featn = ops.layernorm(feat, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    cf = torch.cat((ops.corr_to_tokens(corr, n), featn), dim=-1)
    qk = ops.linear(cf, _cat_cached(sd, f"_{p}.qk_proj.weight", (p + ".q_proj.weight", p + ".k_proj.weight")),
                    _cat_cached(sd, f"_{p}.qk_proj.bias", (p + ".q_proj.bias", p + ".k_proj.bias")))
    pos = sd[p + ".pos_embed"]
    query = qk[..., :C].reshape(B, L, NHEAD, HEAD_DIM) + pos
    key = qk[..., C:].reshape(B, L, NHEAD, HEAD_DIM) + pos
    value_feat = ops.linear(featn, sd[p + ".v_proj.weight"], sd[p + ".v_proj.bias"]).reshape(B, L, NHEAD, HEAD_DIM)
    value_corr = ops.encoder4d(corr, _enc_params(sd, p + ".v_proj_corr"), 1, 1)
    value_corr = ops.corr_to_tokens(value_corr, n).reshape(B, L, NHEAD, -1)
    msg_feat = ops.linear_attention(query, key, value_feat).reshape(B, L, C)
    msg_corr = ops.linear_attention(query, key, value_corr)
    msg_corr = ops.tokens_to_corr(msg_corr.reshape(B, L, -1), n, NHEAD, corr.shape[2])
    msg_feat = feat + msg_feat
    msg_corr = corr + msg_corr
    msg_feat = msg_feat + _mlp(ops, sd, p + ".mlp", ops.layernorm(msg_feat, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"]), n)
    msg_corr = msg_corr + ops.encoder4d(msg_corr, _enc_params(sd, p + ".mlp_corr"), 1, 1)
    return msg_corr, msg_feat


def _forward_cross(ops, sd, p, corr, src, trg, n):
    """UFCLayer.forward_cross (aggregation.py:312-340)."""
    B, L, C = src.shape
    hs = corr.shape[2]
    pool = n // hs

    def values(x):
        xr = ops.avgpool_tokens(x, n, pool)
        xr = ops.layernorm(xr, sd[p + ".norm_cross1.weight"], sd[p + ".norm_cross1.bias"])
        return ops.linear(xr, sd[p + ".v_cross.weight"], sd[p + ".v_cross.bias"]).reshape(B, -1, NHEAD, HEAD_DIM)

    src_attn, trg_attn = ops.cross_attention(corr, values(src), values(trg))
    src = src + ops.repeat_tokens(src_attn, hs, pool)
    trg = trg + ops.repeat_tokens(trg_attn, hs, pool)
    nw, nb = sd[p + ".norm_cross2.weight"], sd[p + ".norm_cross2.bias"]
    src = src + _mlp(ops, sd, p + ".mlp_cross", ops.layernorm(src, nw, nb), n)
    trg = trg + _mlp(ops, sd, p + ".mlp_cross", ops.layernorm(trg, nw, nb), n)
    return src, trg


def _layer(ops, sd, p, level, corr, src, trg, n):
    """UFCLayer.forward (aggregation.py:342-356)."""
    k, s, pad = LEVEL_CONV[level]
    corr_src, src_r = _forward_attention(ops, sd, p, corr, src, n)
    corr_trg, trg_r = _forward_attention(ops, sd, p, ops.transpose4d(corr), trg, n)
    corr_r = corr_src + ops.transpose4d(corr_trg)
    corr_r = corr_r + ops.encoder4d(ops.correlation(src_r, trg_r, n), _enc_params(sd, p + ".feat_to_corr1"), s, pad)
    corr_r = corr_r + ops.encoder4d(corr_r, _enc_params(sd, p + ".mlp_refine_corr"), 1, 1)
    src_r, trg_r = _forward_cross(ops, sd, p, corr_r, src_r, trg_r, n)
    corr_r = corr_r + ops.encoder4d(ops.correlation(src_r, trg_r, n), _enc_params(sd, p + ".feat_to_corr2"), s, pad)
    corr_r = corr_r + ops.encoder4d(corr_r, _enc_params(sd, p + ".mlp_refine_corr2"), 1, 1)
    return corr_r, src_r, trg_r


@torch.no_grad()
def ufc_forward(sd, feat, nview, ops):
    """sd: state_dict of the reference UFC module (keys without the 'feature_cost_aggregation.' prefix).
    feat: [(2B, 512, 16, 16), (2B, 256, 32, 32), (2B, 128, 64, 64)]. Returns what UFC.forward returns."""
    B2 = feat[0].shape[0]
    sizes = [f.shape[-1] for f in feat]
    launches0 = getattr(ops, "launches", 0)

    def side(i, v):
        x = feat[i].reshape(B2 // nview, nview, -1, sizes[i] * sizes[i])[:, v].transpose(1, 2)   # 'B C H W -> B (H W) C'
        return ops.linear(x, sd[f"proj_feat.{i}.0.weight"], sd[f"proj_feat.{i}.0.bias"], act="relu")

    src = [side(i, 0) for i in range(3)]
    trg = [side(i, 1) for i in range(3)]
    feat_list, refined = [], []
    corr = s = t = None
    for lvl in range(3):
        n = sizes[lvl]
        k, stride, pad = LEVEL_CONV[lvl]
        raw = ops.encoder4d(ops.correlation(src[lvl], trg[lvl], n), _enc_params(sd, f"embedding.{lvl}"), stride, pad)
        corr = raw if lvl == 0 else corr + raw
        s = src[lvl] if lvl == 0 else ops.upsample_tokens(s, n) + src[lvl]
        t = trg[lvl] if lvl == 0 else ops.upsample_tokens(t, n) + trg[lvl]
        for j in range(LAYER_NUMS[lvl]):
            corr, s, t = _layer(ops, sd, f"layers.{lvl}.{j}", lvl, corr, s, t, n)
        both = torch.stack((s, t), dim=1).flatten(0, 1)                        # (2B, L, C)
        feat_list.append(both.transpose(1, 2).reshape(both.shape[0], both.shape[2], n, n))
        refined.append((s, t))
    flows, c = ops.tail([r[0] for r in refined], [r[1] for r in refined], tuple(sizes), sizes[-1])
    if hasattr(ops, "launches"):
        ops.launches_per_forward = ops.launches - launches0
    return feat_list, flows, c
