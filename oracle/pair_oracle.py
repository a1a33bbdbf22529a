"""CPU restatement of CoPoNeRF.get_z()  --  TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs). Nothing in coponerf_b200/ may import this file.

Follows the reference's own formulation line by line (not the re-associated one of coponerf_b200/pose_native.py):
  get_z                      models/CoPoNeRF.py:159-206
  SpatialEncoder.forward     models/backbone.py:65-104 (torchvision resnet34, use_first_pool=False)
  normalize_imagenet         utils_training/utils.py:247-257
  get_positional_encodings   models/backbone.py:209-278 (the Python double loop is kept: it is what the reference
                             spends ~0.3 s per call on; `fast_pos=True` swaps in one batched matmul, same values)
  CrossAttention.forward     models/backbone.py:279-330 (both dual softmaxes, (v^T A) v association)
  CrossBlock.forward         models/backbone.py:400-420
  pose head, r6d2mat         models/CoPoNeRF.py:33-52,106-128,194-204
The cost aggregation (models/aggregation.py:509-562) is oracle/ufc_forward_oracle.py, a restatement in the reference's own
formulation that tests/test_ufc_orchestration_cpu.py pins to the unmodified reference module. Nothing here imports
coponerf_b200/.

Pinned by tests/golden/pair_256.npz (outputs of the unmodified reference's get_z, tests/golden/make_goldens_pair.py).
"""
import torch
import torch.nn.functional as F

from . import ufc_forward_oracle


def _encoder(sd, x):
    import torchvision
    net = torchvision.models.resnet34(weights=None)
    net.fc = torch.nn.Sequential()
    net.avgpool = torch.nn.Sequential()
    net.load_state_dict({k[len("encoder.model."):]: v for k, v in sd.items() if k.startswith("encoder.model.")})
    net.eval()
    x = net.relu(net.bn1(net.conv1(x)))
    latents = [x]
    x = net.layer1(x)
    latents.append(x)
    x = net.layer2(x)
    latents.append(x)
    x = net.layer3(x)
    latents.append(x)
    x = net.layer4(x)
    latents.append(x)
    return latents[::-1][:3]


def positional_encodings(B, N, intr, fast_pos=False):
    """backbone.py:209-278."""
    h, w = 48, 64
    if N == 64 * 64:
        h, w = 64, 64
    elif N != 48 * 64:
        raise AssertionError("unexpected resolution for positional encoding")
    positional = torch.ones([B, N, 6])
    ys = torch.linspace(-1, 1, steps=h)
    xs = torch.linspace(-1, 1, steps=w)
    p3 = ys.unsqueeze(0).repeat(B, w)
    p4 = xs.repeat_interleave(h).unsqueeze(0).repeat(B, 1)
    fx, fy, cx, cy = intr
    hpix, wpix = cy * 2, cx * 2
    K = torch.zeros([B, 3, 3])
    K[:, 0, 0] = ((fx / wpix) * 2).squeeze()
    K[:, 1, 1] = ((fy / hpix) * 2).squeeze()
    K[:, 0, 2] = ((cx / wpix) * 2 - 1).squeeze()
    K[:, 1, 2] = ((cy / hpix) * 2 - 1).squeeze()
    K[:, 2, 2] = 1
    Kinv = torch.inverse(K)
    if fast_pos:
        kk, jj = torch.meshgrid(torch.arange(w), torch.arange(h), indexing="ij")
        kk, jj = kk.reshape(-1), jj.reshape(-1)
        wv = Kinv @ torch.stack((xs[kk], ys[jj], torch.ones(kk.numel())))
        p3[:, kk * w + jj] = wv[:, 1] / wv[:, 2]
        p4[:, kk * w + jj] = wv[:, 0] / wv[:, 2]
    else:
        for j in range(h):
            for k in range(w):
                w1, w2, w3 = torch.split(Kinv @ torch.tensor([xs[k], ys[j], 1]), 1, dim=1)
                p3[:, int(k * w + j)] = w2.squeeze() / w3.squeeze()
                p4[:, int(k * w + j)] = w1.squeeze() / w3.squeeze()
    positional[:, :, :5] = torch.stack([p3 * p3, p4 * p4, p3 * p4, p3, p4], dim=2)
    return positional


def cross_block(sd, x, corr, intr, fast_pos=False):
    """CrossBlock.forward (backbone.py:400-420): x (2B, L, 256) -> (2B, 262, 256)."""
    p = "cross_attention."
    b_s, h_w, nf = x.shape
    x = x.reshape([-1, 2, h_w, nf])
    ln = lambda t, n: F.layer_norm(t, (nf,), sd[p + n + ".weight"], sd[p + n + ".bias"], 1e-5)
    x1, x2 = ln(x[:, 0], "norm1"), ln(x[:, 1], "norm1")
    B, N, C = x1.shape
    attn_1 = corr.squeeze(1).flatten(-2, -1).flatten(1, 2)
    attn_2 = attn_1.transpose(-2, -1)
    af1 = attn_1.softmax(dim=-1) * attn_1.softmax(dim=-2)
    af2 = attn_2.softmax(dim=-1) * attn_2.softmax(dim=-2)
    positional = positional_encodings(B, N, intr, fast_pos)
    v1 = torch.cat([x1, positional], dim=2)
    v2 = torch.cat([x2, positional], dim=2)
    f1 = (v1.transpose(-2, -1) @ af1) @ v1
    f2 = (v2.transpose(-2, -1) @ af2) @ v2
    f1 = f1.reshape(B, C + 6, C + 6).transpose(-2, -1)
    f2 = f2.reshape(B, C + 6, C + 6).transpose(-2, -1)
    proj = lambda t: F.linear(t, sd[p + "cross_attn.proj_fundamental.weight"], sd[p + "cross_attn.proj_fundamental.bias"])
    f2, f1 = proj(f2), proj(f1)
    fund = torch.cat([f2.unsqueeze(1), f1.unsqueeze(1)], dim=1).reshape(b_s, -1, nf)
    h = F.linear(ln(fund, "norm2"), sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.linear(F.gelu(h), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return ln(fund + h, "norm")


def pose_head(h0, sd):
    """CoPoNeRF.py:33-52,106-128,198-204 after the first Linear + ReLU: h0 = relu(pose_regressor[0](pose_feat))."""
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
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(h0.shape[0], 1, -1)
    return torch.cat((torch.cat((R, tr.unsqueeze(-1)), dim=-1), bottom), dim=1)


@torch.no_grad()
def get_z(sd, inp, fast_pos=False, timings=None):
    """sd: full model state_dict (CPU tensors). Returns (z list, rel_pose, flows) like CoPoNeRF.get_z."""
    import time
    t0 = time.perf_counter()
    rgb = inp["context"]["rgb"]
    B, n_ctxt, H, W, _ = rgb.shape
    x = torch.flatten(rgb, 0, 1).permute(0, -1, 1, 2)
    x = ((x + 1) / 2.).clone()
    x[:, 0] = (x[:, 0] - 0.485) / 0.229
    x[:, 1] = (x[:, 1] - 0.456) / 0.224
    x[:, 2] = (x[:, 2] - 0.406) / 0.225
    z = _encoder(sd, x)
    z_conv = F.conv2d(x, sd["conv_map.weight"], sd["conv_map.bias"], padding=3)
    t1 = time.perf_counter()
    pre = "feature_cost_aggregation."
    ufc_sd = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    feats, flows, c = ufc_forward_oracle.ufc_forward(ufc_sd, z, 2)
    t2 = time.perf_counter()
    k = inp["context"]["intrinsics"].clone()
    k[:, :, :2, :] = k[:, :, :2, :] / H
    intr = [k[:, 0, 0, 0].reshape(B, 1), k[:, 0, 1, 1].reshape(B, 1), k[:, 0, 0, 2].reshape(B, 1), k[:, 0, 1, 2].reshape(B, 1)]
    feat = cross_block(sd, feats[-1].flatten(-2, -1).transpose(-1, -2), c, intr, fast_pos).reshape([B, -1])
    h0 = F.relu(F.linear(feat, sd["pose_regressor.0.weight"], sd["pose_regressor.0.bias"]))
    rel_pose = pose_head(h0, sd)
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(encoder=t1 - t0, ufc=t2 - t1, pose=t3 - t2)
    return feats + [z_conv], rel_pose, flows
