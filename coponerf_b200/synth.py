"""Seeded synthetic inputs and weights for the render path.

Everything is drawn from numpy's PCG64 so the same seed yields the same
bytes in this container, on the GPU box and inside the golden-vector script.
Shapes follow the reference dataloader's output dict
(data/realestate10k_dataio.py:442-451) and the render-path parameters of
models/CoPoNeRF.py:71-104 / models/lightfield.py:87-116.
"""
import math

import numpy as np
import torch

# name -> shape of every parameter the per-ray stage reads (models/CoPoNeRF.py:71-104).
RENDER_PARAM_SHAPES = {
    "query_encode_latent.weight": (832, 835, 1, 1),
    "query_encode_latent.bias": (832,),
    "query_encode_latent_2.weight": (416, 832, 1, 1),
    "query_encode_latent_2.bias": (416,),
    "latent_value.weight": (416, 832, 1, 1),
    "latent_value.bias": (416,),
    "key_map.weight": (128, 832, 1, 1),
    "key_map.bias": (128,),
    "key_map_2.weight": (128, 128, 1, 1),
    "key_map_2.bias": (128,),
    "query_embed.weight": (128, 16, 1, 1),
    "query_embed.bias": (128,),
    "query_embed_2.weight": (128, 128, 1, 1),
    "query_embed_2.bias": (128,),
    "query_repeat_embed.weight": (128, 144, 1, 1),
    "query_repeat_embed.bias": (128,),
    "query_repeat_embed_2.weight": (128, 128, 1, 1),
    "query_repeat_embed_2.bias": (128,),
    "encode_latent.weight": (128, 416, 1),
    "encode_latent.bias": (128,),
    "phi.lin_in.weight": (128, 18),
    "phi.lin_in.bias": (128,),
    "phi.lin_z.0.weight": (128, 832),
    "phi.lin_z.0.bias": (128,),
    "phi.lin_z.1.weight": (128, 832),
    "phi.lin_z.1.bias": (128,),
    "phi.lin_z.2.weight": (128, 832),
    "phi.lin_z.2.bias": (128,),
    "phi.blocks.0.fc_0.weight": (128, 128),
    "phi.blocks.0.fc_0.bias": (128,),
    "phi.blocks.0.fc_1.weight": (128, 128),
    "phi.blocks.0.fc_1.bias": (128,),
    "phi.blocks.1.fc_0.weight": (128, 128),
    "phi.blocks.1.fc_0.bias": (128,),
    "phi.blocks.1.fc_1.weight": (128, 128),
    "phi.blocks.1.fc_1.bias": (128,),
    "phi.blocks.2.fc_0.weight": (128, 128),
    "phi.blocks.2.fc_0.bias": (128,),
    "phi.blocks.2.fc_1.weight": (128, 128),
    "phi.blocks.2.fc_1.bias": (128,),
    "phi.lin_out.weight": (3, 128),
    "phi.lin_out.bias": (3,),
}

# Layers whose gain is raised so the 128-way softmax is peaked rather than
# uniform; a uniform softmax makes the exported argmax (at_wt_max) a coin toss.
_GAIN = {"key_map_2.weight": 6.0, "query_embed_2.weight": 6.0, "query_repeat_embed_2.weight": 6.0}


def render_state_dict(seed=0):
    """Render-path weights, U(-g/sqrt(fan_in), g/sqrt(fan_in)); biases U(-0.1, 0.1).

    The reference zero-initialises phi.blocks.*.fc_1 and every phi bias
    (models/lightfield.py:35-38,88-93); they are randomised here so the whole
    decoder is exercised (SURVEY.md section 8(c) pitfall i).
    """
    rng = np.random.default_rng(1000 + seed)
    sd = {}
    for name, shape in RENDER_PARAM_SHAPES.items():
        if name.endswith(".bias"):
            a = rng.uniform(-0.1, 0.1, size=shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = _GAIN.get(name, 1.7) / math.sqrt(fan_in)
            a = rng.uniform(-bound, bound, size=shape)
        sd[name] = torch.from_numpy(a.astype(np.float32))
    return sd


def _pose(tx, yaw, ty=0.0, tz=0.0):
    c, s = math.cos(yaw), math.sin(yaw)
    m = np.eye(4, dtype=np.float64)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    m[:3, 3] = (tx, ty, tz)
    return m.astype(np.float32)


POSE_SETS = {
    # every target ray has a valid epipolar segment in both context views
    "frontal": dict(ctx=((-0.15, 0.05), (0.15, -0.05)), qry=(0.0, 0.0)),
    # 54 % valid in both / 9 % in one / 37 % in none
    "oblique": dict(ctx=((-0.15, 0.05), (0.15, -0.05)), qry=(0.6, 0.4)),
    "mild": dict(ctx=((-0.15, 0.05), (0.15, -0.05)), qry=(0.3, 0.25)),
}


def smooth_image(rng, H, W):
    """Sum of 8 random sinusoids per channel, in [-1, 1]."""
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    img = np.zeros((H, W, 3), dtype=np.float64)
    for c in range(3):
        for _ in range(8):
            fx, fy = rng.uniform(-6, 6, size=2)
            ph = rng.uniform(0, 2 * math.pi)
            img[..., c] += rng.uniform(0.2, 1.0) * np.sin(2 * math.pi * (fx * xx + fy * yy) + ph)
    img /= np.abs(img).max()
    return img.astype(np.float32)


def make_input(H=256, W=256, n_rays=None, seed=1, pose_set="frontal", batch=1, focal=0.9):
    """The dict CoPoNeRF.forward() takes (models/CoPoNeRF.py:208-216)."""
    rng = np.random.default_rng(2000 + seed)
    ps = POSE_SETS[pose_set]
    K = np.eye(4, dtype=np.float32)
    K[0, 0] = K[1, 1] = focal * W
    K[0, 2], K[1, 2] = W / 2, H / 2
    ctx_c2w = np.stack([_pose(*ps["ctx"][0]), _pose(*ps["ctx"][1])])[None].repeat(batch, 0)
    qry_c2w = _pose(*ps["qry"])[None, None].repeat(batch, 0)
    rgb = np.stack([np.stack([smooth_image(rng, H, W) for _ in range(2)]) for _ in range(batch)])
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    uv_full = np.stack([xs.reshape(-1), ys.reshape(-1)], -1).astype(np.float32)
    if n_rays is None or n_rays >= H * W:
        uv = np.broadcast_to(uv_full, (batch, 1, H * W, 2)).copy()
    else:
        uv = np.stack([uv_full[rng.permutation(H * W)[:n_rays]] for _ in range(batch)])[:, None]
    N = uv.shape[2]
    t = torch.from_numpy
    return {
        "context": {
            "rgb": t(rgb),
            "cam2world": t(ctx_c2w.copy()),
            "intrinsics": t(K[None, None].repeat(batch, 0).repeat(2, 1)),
        },
        "query": {
            "uv": t(uv),
            "cam2world": t(qry_c2w.copy()),
            "intrinsics": t(K[None, None].repeat(batch, 0)),
            "rgb": t(rng.uniform(-1, 1, size=(batch, 1, N, 3)).astype(np.float32)),
        },
    }


def make_features(H=256, W=256, seed=1, batch=1):
    """Stand-ins for get_z()'s outputs at any resolution.

    Returns (z list of 4 maps, rel_pose (B,4,4), flow tuple of 4).
    """
    rng = np.random.default_rng(3000 + seed)
    n = 2 * batch
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    z = [
        f32(rng.standard_normal((n, 256, H // 16, W // 16))),
        f32(rng.standard_normal((n, 256, H // 8, W // 8))),
        f32(rng.standard_normal((n, 256, H // 4, W // 4))),
        f32(rng.standard_normal((n, 64, H, W))),
    ]
    fh, fw = H // 4, W // 4
    flow = (
        f32(rng.standard_normal((batch, 2, fh, fw)) * 2),
        f32(rng.standard_normal((batch, 2, fh, fw)) * 2),
        f32(rng.uniform(-1, 1, size=(batch, 2, fh, fw))),
        f32(rng.uniform(-1, 1, size=(batch, 2, fh, fw))),
    )
    # estimated context-1 -> context-2 pose: close to the true one, not equal to it
    rel = np.stack([_pose(0.3, -0.1, 0.01 * b, -0.02) for b in range(batch)])
    return z, f32(rel), flow


def ufc_tail_features(sizes=(16, 32, 64), batch=1, seed=12, C=256):
    """Refined source / target token features for the closing stage of UFC (three (B, n*n, C) tensors each).

    Smooth random maps plus noise; the target is a shifted copy of the source, so neighbouring pixels correlate and
    the soft-argmax flow has structure instead of being uniform.
    """
    import torch.nn.functional as F
    rng = np.random.default_rng(4000 + seed)
    src, trg = [], []
    for n in sizes:
        base = torch.from_numpy(rng.standard_normal((batch, 8, 8, C))).permute(0, 3, 1, 2).float()
        up = F.interpolate(base, size=(n, n), mode="bicubic", align_corners=True)
        shift = torch.roll(up, shifts=(max(1, n // 8), -max(1, n // 16)), dims=(2, 3))
        noise = lambda: torch.from_numpy(rng.standard_normal((batch, C, n, n))).float() * 0.3
        src.append((up + noise()).flatten(2).transpose(1, 2).contiguous())
        trg.append((shift + noise()).flatten(2).transpose(1, 2).contiguous())
    return src, trg


# (name, B, [channels...], k, stride, pad, input size) of the Encoder4D blocks UFC uses (models/aggregation.py:209-318,436-480)
CONV4D_CASES = {
    "conv4d_embed16": (2, (1, 8), 3, 1, 1, 16),        # embedding[0] / feat_to_corr of level 0
    "conv4d_mlp16": (1, (8, 32, 8), 3, 1, 1, 16),      # mlp_corr / mlp_refine_corr: two blocks
    "conv4d_embed32": (1, (1, 8), 3, 2, 1, 32),        # embedding[1]: stride 2 with MaxPool4d(2)
    "conv4d_embed64": (1, (1, 8), 5, 4, 2, 64),        # embedding[2]: kernel 5, stride 4 with MaxPool4d(4)
}


def conv4d_case(name):
    """Seeded input volume and Encoder4D parameters for one CONV4D_CASES entry (numpy PCG64: same bytes everywhere)."""
    B, chans, k, stride, pad, n = CONV4D_CASES[name]
    rng = np.random.default_rng(5000 + sorted(CONV4D_CASES).index(name))
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    x = f32(rng.standard_normal((B, chans[0], n, n, n, n), dtype=np.float32))
    layers = []
    for ci, co in zip(chans[:-1], chans[1:]):
        bound = 1.5 / math.sqrt(ci * k * k)
        layers.append(dict(wq=f32(rng.uniform(-bound, bound, (co, ci, k, k))), bq=f32(rng.uniform(-0.2, 0.2, co)),
                           ws=f32(rng.uniform(-bound, bound, (co, ci, k, k))), bs=f32(rng.uniform(-0.2, 0.2, co)),
                           gamma=f32(rng.uniform(0.5, 1.5, co)), beta=f32(rng.uniform(-0.3, 0.3, co))))
    return x, layers, stride, pad


# (N, L = S, Dv) of the LinearAttention calls in UFCLayer.forward_attention (models/aggregation.py:296-297)
LINATT_CASES = {"linatt_256_feat": (2, 256, 32), "linatt_256_corr": (2, 256, 256), "linatt_1024_corr": (1, 1024, 256),
                "linatt_4096_feat": (1, 4096, 32)}


def linatt_case(name, H=8, D=32):
    N, L, Dv = LINATT_CASES[name]
    rng = np.random.default_rng(6000 + sorted(LINATT_CASES).index(name))
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return (f32(rng.standard_normal((N, L, H, D))), f32(rng.standard_normal((N, L, H, D))),
            f32(rng.standard_normal((N, L, H, Dv))))


def ufc_param_shapes(sizes=(16, 32, 64)):
    """name -> shape of every parameter of the reference UFC module (models/aggregation.py:358-490), without the
    'feature_cost_aggregation.' prefix. 439 entries. `sizes` are the feature-map sizes of the three levels; the
    correlation volumes live at sizes[0] (16 at 256x256, the only geometry the reference itself supports; 32 for the
    512x512 restatement of BASELINE config 4, where q_proj / k_proj take 256 + 8 * 32^2 inputs)."""
    sh = {}
    corr = sizes[0]

    def enc(prefix, chans, k):
        for i, (ci, co) in enumerate(zip(chans[:-1], chans[1:])):
            for br in ("query_conv", "supp_conv"):
                sh[f"{prefix}.conv4d.{i}.0.{br}.weight"] = (co, ci, k, k)
                sh[f"{prefix}.conv4d.{i}.0.{br}.bias"] = (co,)
            sh[f"{prefix}.conv4d.{i}.1.weight"] = (co,)
            sh[f"{prefix}.conv4d.{i}.1.bias"] = (co,)

    def lin(prefix, o, i):
        sh[prefix + ".weight"] = (o, i)
        sh[prefix + ".bias"] = (o,)

    for lvl, (nlayers, n, k) in enumerate(((2, sizes[0], 3), (2, sizes[1], 3), (1, sizes[2], 5))):
        for j in range(nlayers):
            p = f"layers.{lvl}.{j}"
            sh[p + ".pos_embed"] = (1, n * n, 1, 32)
            lin(p + ".q_proj", 256, 256 + 8 * corr * corr)
            lin(p + ".k_proj", 256, 256 + 8 * corr * corr)
            lin(p + ".v_proj", 256, 256)
            enc(p + ".v_proj_corr", (8, 8), 3)
            for m in ("mlp", "mlp_cross"):
                lin(f"{p}.{m}.0", 1024, 256)
                sh[f"{p}.{m}.1.dwconv.weight"] = (1024, 1, 3, 3)
                sh[f"{p}.{m}.1.dwconv.bias"] = (1024,)
                lin(f"{p}.{m}.3", 256, 1024)
            for m in ("mlp_corr", "mlp_refine_corr", "mlp_refine_corr2"):
                enc(f"{p}.{m}", (8, 32, 8), 3)
            enc(p + ".feat_to_corr1", (1, 8), k)
            enc(p + ".feat_to_corr2", (1, 8), k)
            for m in ("norm1", "norm2", "norm_cross1", "norm_cross2"):
                sh[f"{p}.{m}.weight"] = (256,)
                sh[f"{p}.{m}.bias"] = (256,)
            lin(p + ".v_cross", 256, 256)
    for lvl, k in enumerate((3, 3, 5)):
        enc(f"embedding.{lvl}", (1, 8), k)
    for lvl, cin in enumerate((512, 256, 128)):
        lin(f"proj_feat.{lvl}.0", 256, cin)
    return sh


def ufc_state_dict(seed=0, sizes=(16, 32, 64)):
    """Seeded random parameters with the reference UFC's names and shapes (numpy PCG64)."""
    rng = np.random.default_rng(7000 + seed)
    sd = {}
    for name, shape in ufc_param_shapes(sizes).items():
        if name.endswith("pos_embed"):
            a = rng.normal(0, 0.02, shape)
        elif ".conv4d." in name and name.endswith(".1.weight") or ".norm" in name and name.endswith("weight"):
            a = rng.uniform(0.7, 1.3, shape)          # GroupNorm / LayerNorm scales
        elif name.endswith(".bias"):
            a = rng.uniform(-0.1, 0.1, shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.uniform(-1, 1, shape) * (1.2 / math.sqrt(fan_in))
        sd[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def ufc_inputs(seed=0, batch=1, sizes=(16, 32, 64)):
    """Encoder feature pyramid as UFC.forward receives it: [(2B,512,16,16), (2B,256,32,32), (2B,128,64,64)]."""
    rng = np.random.default_rng(7500 + seed)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return [f32(rng.standard_normal((2 * batch, c, n, n), dtype=np.float32)) for c, n in zip((512, 256, 128), sizes)]


# ---------------------------------------------------------------------------------------------------------------
# the rest of the model: image encoder, conv_map, pose features and pose head (models/CoPoNeRF.py:31-69)

# parameters the reference constructs but never reads in forward() (SURVEY.md section 8(a) notes)
UNUSED_PARAM_SHAPES = {
    "corr_embed.weight": (832, 4096, 1, 1), "corr_embed.bias": (832,),
    "latent_avg_query.weight": (128, 25, 1, 1), "latent_avg_query.bias": (128,),
    "latent_avg_query_2.weight": (128, 128, 1, 1), "latent_avg_query_2.bias": (128,),
    "latent_avg_key.weight": (128, 416, 1, 1), "latent_avg_key.bias": (128,),
    "latent_avg_key_2.weight": (128, 128, 1, 1), "latent_avg_key_2.bias": (128,),
    "latent_avg_repeat_query.weight": (128, 153, 1, 1), "latent_avg_repeat_query.bias": (128,),
    "latent_avg_repeat_query_2.weight": (128, 128, 1, 1), "latent_avg_repeat_query_2.bias": (128,),
}


def encoder_param_shapes():
    """'encoder.model.*' = torchvision resnet34 without fc (models/backbone.py:52-58): name -> (shape, dtype)."""
    import torchvision
    net = torchvision.models.resnet34(weights=None)
    return {"encoder.model." + k: (tuple(v.shape), v.dtype) for k, v in net.state_dict().items() if not k.startswith("fc.")}


def pair_state_dict(seed=0):
    """Seeded parameters of everything get_z() reads besides the cost aggregation (numpy PCG64): the ResNet-34
    encoder with non-trivial BatchNorm statistics, conv_map, CrossBlock and the three regressors. Output scales are
    chosen so the estimated pose is a moderate rotation / translation instead of saturating."""
    from .pair_stage import POSE_PARAM_SHAPES
    rng = np.random.default_rng(9000 + seed)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    sd = {}
    for name, (shape, dtype) in encoder_param_shapes().items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros(shape, dtype=dtype)
        elif name.endswith("running_mean"):
            sd[name] = f32(rng.normal(0, 0.1, shape))
        elif name.endswith("running_var"):
            sd[name] = f32(rng.uniform(0.75, 1.25, shape))
        elif len(shape) == 1:
            sd[name] = f32(rng.uniform(0.7, 1.3, shape) if name.endswith("weight") else rng.uniform(-0.1, 0.1, shape))
        else:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = f32(rng.uniform(-1, 1, shape) * (1.7 / math.sqrt(fan_in)))
    sd["conv_map.weight"] = f32(rng.uniform(-1, 1, (64, 3, 7, 7)) * (1.7 / math.sqrt(147)))
    sd["conv_map.bias"] = f32(rng.uniform(-0.1, 0.1, 64))
    for name, shape in POSE_PARAM_SHAPES.items():
        if ".norm" in name:
            a = rng.uniform(0.7, 1.3, shape) if name.endswith("weight") else rng.uniform(-0.1, 0.1, shape)
        elif name.endswith(".bias"):
            a = rng.uniform(-0.1, 0.1, shape)
        else:
            a = rng.uniform(-1, 1, shape, ) * (1.7 / math.sqrt(shape[1]))
        sd[name] = f32(a)
    return sd


def full_state_dict(seed=0):
    """All 744 state_dict entries of the reference model (models/CoPoNeRF.py:19-104), seeded."""
    sd = dict(render_state_dict(seed))
    sd.update({"feature_cost_aggregation." + k: v for k, v in ufc_state_dict(seed).items()})
    sd.update(pair_state_dict(seed))
    rng = np.random.default_rng(9500 + seed)
    for name, shape in UNUSED_PARAM_SHAPES.items():
        sd[name] = torch.from_numpy(rng.uniform(-0.05, 0.05, shape).astype(np.float32))
    return sd
