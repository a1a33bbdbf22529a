"""PyTorch restatement of the operator set of coponerf_b200/ufc_native.py  --  TEST INFRASTRUCTURE ONLY.

Each function follows the lines of /root/reference/models/aggregation.py / conv4d.py it cites. Used (a) to check the
state_dict-driven orchestration of ufc_native.ufc_forward against the unmodified reference UFC on CPU and (b) as the
per-operator reference for the CUDA operators. Nothing in coponerf_b200/ may import this file.
"""
import torch
import torch.nn.functional as F

from . import conv4d_oracle, ufc_oracle


class TorchOps:
    def layernorm(self, x, w, b):                       # nn.LayerNorm(d_model), aggregation.py:257-258
        return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)

    def linear(self, x, w, b, act=None):
        y = F.linear(x, w, b)
        return F.relu(y) if act == "relu" else F.gelu(y) if act == "gelu" else y

    # ---- operators of coponerf_b200/pose_native.py (models/backbone.py:279-330, models/CoPoNeRF.py:33-52,106-128)
    def matmul(self, a, b):
        return a @ b

    def matmul_tn(self, a, b, bias=None, act=None):
        y = a.transpose(0, 1) @ b
        if bias is not None:
            y = y + bias
        return F.relu(y) if act == "relu" else F.gelu(y) if act == "gelu" else y

    def linear_skinny(self, x, w, b, act=None):
        return self.linear(x, w, b, act)

    def dual_softmax(self, c):                          # backbone.py:290-291
        return c.softmax(dim=-1) * c.softmax(dim=-2)

    def pose_head(self, h0, sd):                        # CoPoNeRF.py:33-52,106-128,198-204; h0 = relu(pose_regressor[0](x))
        h = F.relu(F.linear(h0, sd["pose_regressor.2.weight"], sd["pose_regressor.2.bias"]))
        h = F.relu(F.linear(h, sd["pose_regressor.4.weight"], sd["pose_regressor.4.bias"]))[:, :128]

        def head(name):
            y = F.relu(h)
            y = F.relu(F.linear(y, sd[name + ".1.weight"], sd[name + ".1.bias"]))
            y = F.relu(F.linear(y, sd[name + ".3.weight"], sd[name + ".3.bias"]))
            return F.linear(y, sd[name + ".5.weight"], sd[name + ".5.bias"])
        d6, tr = head("rotation_regressor"), head("translation_regressor")
        a1, a2 = d6[..., :3], d6[..., 3:]
        b1 = F.normalize(a1, dim=-1)
        b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
        R = torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)
        bottom = torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(h0.shape[0], 1, -1).to(h0.device)
        return torch.cat((torch.cat((R, tr.unsqueeze(-1)), dim=-1), bottom), dim=1)

    def corr_to_tokens(self, corr, n):                  # aggregation.py:283-285
        B, H, hs, ws, ht, wt = corr.shape
        x = corr.permute(0, 1, 4, 5, 2, 3).reshape(B, H * ht * wt, hs, ws)
        x = F.interpolate(x, size=(n, n), mode="bilinear", align_corners=True)
        return x.flatten(2).transpose(1, 2)

    def tokens_to_corr(self, tok, n, H, hs):            # aggregation.py:298-300
        B, L, CH = tok.shape
        q = int(round((CH // H) ** 0.5))
        x = tok.reshape(B, n, n, H, q, q).permute(0, 3, 4, 5, 1, 2).reshape(B, CH, n, n)
        x = F.interpolate(x, size=(hs, hs), mode="bilinear", align_corners=True)
        return x.reshape(B, H, q, q, hs, hs).permute(0, 1, 4, 5, 2, 3)

    def transpose4d(self, corr):                        # aggregation.py:344
        return corr.permute(0, 1, 4, 5, 2, 3)

    def encoder4d(self, x, blocks, stride, pad):        # conv4d.py:138-163
        for p in blocks:
            x = conv4d_oracle.encoder4d_layer(x.contiguous(), p, stride, pad)
        return x

    def linear_attention(self, q, k, v):                # aggregation.py:84-117
        return ufc_oracle.linear_attention(q, k, v)

    def dwconv_gelu(self, x, w, b, n):                  # DWConv + nn.GELU, aggregation.py:18-29,186-187
        B, L, C = x.shape
        y = F.conv2d(x.transpose(1, 2).reshape(B, C, n, n), w, b, stride=1, padding=1, groups=C)
        return F.gelu(y.flatten(2).transpose(1, 2))

    def correlation(self, src, trg, n):                 # aggregation.py:70-80
        return ufc_oracle.correlation_token(src, trg, n)

    def avgpool_tokens(self, x, n, pool):               # aggregation.py:316-317 (einops reduce 'mean')
        B, L, C = x.shape
        if pool == 1:
            return x
        y = F.avg_pool2d(x.transpose(1, 2).reshape(B, C, n, n), pool)
        return y.flatten(2).transpose(1, 2)

    def repeat_tokens(self, x, hs, pool):               # aggregation.py:327-332 (einops repeat)
        B, L, C = x.shape
        if pool == 1:
            return x
        y = x.transpose(1, 2).reshape(B, C, hs, hs)
        y = y.repeat_interleave(pool, dim=2).repeat_interleave(pool, dim=3)
        return y.flatten(2).transpose(1, 2)

    def cross_attention(self, corr, src_v, trg_v):      # aggregation.py:314,324-325
        B, H = corr.shape[:2]
        c2 = corr.reshape(B, H, corr.shape[2] * corr.shape[3], -1)
        src_attn = torch.einsum("bhst,bthc->bshc", c2.softmax(-1), trg_v).reshape(B, c2.shape[2], -1)
        trg_attn = torch.einsum("bhst,bshc->bthc", c2.softmax(-2), src_v).reshape(B, c2.shape[3], -1)
        return src_attn, trg_attn

    def upsample_tokens(self, x, n_out):                # interpolate2d_token, aggregation.py:58-63
        B, L, C = x.shape
        n = int(round(L ** 0.5))
        y = F.interpolate(x.transpose(1, 2).reshape(B, C, n, n), size=(n_out, n_out), mode="bilinear", align_corners=True)
        return y.flatten(2).transpose(1, 2)

    def tail(self, src, trg, sizes, out):               # aggregation.py:551-562
        return ufc_oracle.ufc_tail(src, trg, sizes, out)
