"""The per-pair stage behind CoPoNeRF.get_z() (models/CoPoNeRF.py:159-206), standalone: no reference code is
needed at run time, and a reference checkpoint loads key for key.

  image encoder     torchvision ResNet-34 without the first max-pool, conv_map 7x7 (models/backbone.py:10-104,
                    models/CoPoNeRF.py:66-69): stays PyTorch/cuDNN, as BASELINE.json asks ("models/backbone.py ...
                    stay"); run in strict fp32 (TF32 off) so the features equal the reference's fp32 ones
  cost aggregation  ufc_native.ufc_forward over the sm_100a operators (models/aggregation.py)
  pose features     pose_native: CrossBlock + regressors over the sm_100a operators (models/backbone.py:262-431)

Parameters live in containers that only reproduce the reference's state_dict names and shapes; none of the
reference's module code exists here.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pose_native, synth, ufc_native


class ParamTree(nn.Module):
    """Parameters registered under dotted state_dict names ('layers.0.1.q_proj.weight'): one nested container per
    name component. Initialisation is generic (inference framework: real values come from a checkpoint)."""

    def __init__(self, shapes=None):
        super().__init__()
        for name, shape in (shapes or {}).items():
            self._add(name.split("."), shape)

    def _add(self, parts, shape):
        if len(parts) == 1:
            self.register_parameter(parts[0], nn.Parameter(_init(parts[0], shape)))
            return
        if parts[0] not in self._modules:
            self.add_module(parts[0], ParamTree())
        self._modules[parts[0]]._add(parts[1:], shape)


def _init(leaf, shape):
    if len(shape) == 1:
        return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
    t = torch.empty(shape)
    if leaf == "pos_embed":
        return nn.init.trunc_normal_(t, std=0.02)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    bound = 1.0 / fan_in ** 0.5
    return t.uniform_(-bound, bound)


# name -> shape of the pose-half parameters (models/CoPoNeRF.py:33-52, models/backbone.py:262-277,383-398)
POSE_PARAM_SHAPES = {
    "cross_attention.norm1.weight": (256,), "cross_attention.norm1.bias": (256,),
    "cross_attention.cross_attn.qkv.weight": (768, 256),               # unused by forward (noess=False), kept for the state_dict
    "cross_attention.cross_attn.proj_fundamental.weight": (256, 262),
    "cross_attention.cross_attn.proj_fundamental.bias": (256,),
    "cross_attention.norm2.weight": (256,), "cross_attention.norm2.bias": (256,),
    "cross_attention.mlp.fc1.weight": (1024, 256), "cross_attention.mlp.fc1.bias": (1024,),
    "cross_attention.mlp.fc2.weight": (256, 1024), "cross_attention.mlp.fc2.bias": (256,),
    "cross_attention.norm.weight": (256,), "cross_attention.norm.bias": (256,),
    "pose_regressor.0.weight": (512, (16 * 16 + 6) * 256 * 2), "pose_regressor.0.bias": (512,),
    "pose_regressor.2.weight": (256, 512), "pose_regressor.2.bias": (256,),
    "pose_regressor.4.weight": (256, 256), "pose_regressor.4.bias": (256,),
    "rotation_regressor.1.weight": (64, 128), "rotation_regressor.1.bias": (64,),
    "rotation_regressor.3.weight": (32, 64), "rotation_regressor.3.bias": (32,),
    "rotation_regressor.5.weight": (6, 32), "rotation_regressor.5.bias": (6,),
    "translation_regressor.1.weight": (64, 128), "translation_regressor.1.bias": (64,),
    "translation_regressor.3.weight": (32, 64), "translation_regressor.3.bias": (32,),
    "translation_regressor.5.weight": (3, 32), "translation_regressor.5.bias": (3,),
}


class SpatialEncoder(nn.Module):
    """models/backbone.py:10-104 as CoPoNeRF configures it (resnet34, num_layers=5, use_first_pool=False): returns
    the three coarsest ResNet stages, coarsest first. Parameter names: 'model.<torchvision resnet34 names>'."""

    def __init__(self):
        super().__init__()
        import torchvision
        self.model = torchvision.models.resnet34(weights=None)
        self.model.fc = nn.Sequential()
        self.model.avgpool = nn.Sequential()

    def forward(self, x):
        m = self.model
        x = m.relu(m.bn1(m.conv1(x)))
        x = m.layer1(x)                       # no max-pool (use_first_pool=False)
        l2 = m.layer2(x)
        l3 = m.layer3(l2)
        l4 = m.layer4(l3)
        return [l4, l3, l2]


class PairStage(nn.Module):
    """Owns every parameter of get_z(): 'encoder.*', 'conv_map.*', 'feature_cost_aggregation.*', 'cross_attention.*',
    '{pose,rotation,translation}_regressor.*'. It is mixed into coponerf_b200.model.CoPoNeRF so the names carry no prefix."""

    @staticmethod
    def build_into(module):
        module.encoder = SpatialEncoder()
        module.conv_map = nn.Conv2d(3, 64, kernel_size=7, stride=1, padding=3)
        module.feature_cost_aggregation = ParamTree(synth.ufc_param_shapes())
        pose = {}
        for k, v in POSE_PARAM_SHAPES.items():
            pose.setdefault(k.split(".")[0], {})[k.split(".", 1)[1]] = v
        for top, shapes in pose.items():
            setattr(module, top, ParamTree(shapes))


_IMAGENET_MEAN = (0.485, 0.456, 0.406)
_IMAGENET_STD = (0.229, 0.224, 0.225)
_NORM_CONST = {}


def _norm_const(dev):
    key = str(dev)
    if key not in _NORM_CONST:      # built once per device, outside any graph capture
        _NORM_CONST[key] = (torch.tensor(_IMAGENET_MEAN, device=dev).view(1, 3, 1, 1),
                            torch.tensor(_IMAGENET_STD, device=dev).view(1, 3, 1, 1))
    return _NORM_CONST[key]


# the folded encoder runs in channels-last memory format (cuDNN picks its NHWC fp32 kernels: 7.33 vs 7.53 ms per pair, same
# values within the test gates); CPN_ENCODER_NHWC=0 keeps NCHW for A/B runs
_ENCODER_NHWC = os.environ.get("CPN_ENCODER_NHWC", "1") == "1"


def _fold_bn(conv, bn):
    """Eval-mode BatchNorm folded into the preceding bias-free convolution: w' = w g / sqrt(var + eps),
    b' = beta - mean g / sqrt(var + eps). Same function, one kernel instead of two."""
    scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    w = conv.weight.detach() * scale.view(-1, 1, 1, 1)
    b = bn.bias.detach() - bn.running_mean * scale
    w = w.contiguous(memory_format=torch.channels_last) if _ENCODER_NHWC else w.contiguous()
    return w, b.contiguous(), conv.stride, conv.padding


def fold_encoder(encoder):
    """SpatialEncoder parameters as a list of folded convolutions, in execution order (stem, then per BasicBlock
    conv1, conv2 and the optional 1x1 downsample)."""
    m = encoder.model
    plan = {"stem": _fold_bn(m.conv1, m.bn1), "layers": []}
    for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
        blocks = []
        for blk in layer:
            down = _fold_bn(blk.downsample[0], blk.downsample[1]) if blk.downsample is not None else None
            blocks.append((_fold_bn(blk.conv1, blk.bn1), _fold_bn(blk.conv2, blk.bn2), down))
        plan["layers"].append(blocks)
    return plan


def _conv(x, p, relu):
    y = F.conv2d(x, p[0], p[1], stride=p[2], padding=p[3])
    return F.relu_(y) if relu else y


def run_folded_encoder(plan, x):
    """torchvision BasicBlock ResNet-34 forward (no max-pool) over the folded convolutions -> [l4, l3, l2]."""
    if _ENCODER_NHWC:
        x = x.contiguous(memory_format=torch.channels_last)
    x = _conv(x, plan["stem"], True)
    outs = []
    for blocks in plan["layers"]:
        for c1, c2, down in blocks:
            idt = x if down is None else _conv(x, down, False)
            x = F.relu_(_conv(_conv(x, c1, True), c2, False).add_(idt))
        outs.append(x)
    return [outs[3], outs[2], outs[1]]


def encode_images(model, rgb, plan=None):
    """rgb (B, n_ctxt, H, W, 3) in [-1, 1] -> (ResNet pyramid [3 tensors], conv_map features). CoPoNeRF.py:171-184
    with utils.normalize_imagenet (utils_training/utils.py:247-257). With `plan` (fold_encoder) the BatchNorms are
    folded into their convolutions; strict fp32 (TF32 off) either way, cuDNN autotuned per shape."""
    x = torch.flatten(rgb, 0, 1).permute(0, 3, 1, 2).to(torch.float32)
    x = (x + 1) / 2.
    mean, std = _norm_const(x.device)
    x = (x - mean) / std
    enc = (lambda t: run_folded_encoder(plan, t)) if plan is not None else model.encoder
    if x.is_cuda:
        with torch.backends.cudnn.flags(enabled=True, benchmark=True, deterministic=False, allow_tf32=False):
            return enc(x), model.conv_map(x)
    return enc(x), model.conv_map(x)


def _sub_state(model, prefix):
    return {k[len(prefix) + 1:]: v for k, v in model.state_dict(keep_vars=True).items() if k.startswith(prefix + ".")}


def _state_cache(model):
    cache = model.__dict__.setdefault("_sd_cache", {})
    ver = model._weights_version() if hasattr(model, "_weights_version") else \
        tuple((p.data_ptr(), p._version) for p in list(model.parameters()) + list(model.buffers()))
    if cache.get("ver") != ver:
        cache.clear()
        ops = model.__dict__.get("_ufc_ops")
        if ops is not None:         # the transposed weight copies belong to the old parameter versions
            ops._wt.clear()
        cache["ver"] = ver
        cache["ufc"] = _sub_state(model, "feature_cost_aggregation")
        cache["encoder"] = fold_encoder(model.encoder)
        full = model.state_dict(keep_vars=True)
        cache["pose"] = {k: v for k, v in full.items()
                         if k.split(".")[0] in ("cross_attention", "pose_regressor", "rotation_regressor",
                                                "translation_regressor")}
    return cache


def _pair_body(model, cache, rgb, pos, ops):
    """Everything of get_z() that runs on the device, as one stream-ordered sequence with no host dependence
    (so it can be captured into a CUDA graph): encoder, cost aggregation, pose features, pose head."""
    z, z_conv = encode_images(model, rgb, cache["encoder"])
    feats, flows, c = ufc_native.ufc_forward(cache["ufc"], z, model.n_view, ops)
    tokens = feats[-1].flatten(-2, -1).transpose(-1, -2)            # (2B, L, 256)
    pose_feat = pose_native.cross_block(cache["pose"], "cross_attention", tokens, c, pos, ops)
    rel_pose = pose_native.pose_head(cache["pose"], pose_feat, ops)
    return feats + [z_conv], rel_pose, flows


class _PairGraph:
    """get_z()'s device work for one input geometry captured once and replayed: ~800 short kernels (the per-pair
    stage is launch-latency bound) become one graph launch. Outputs are cloned out of the graph's
    static buffers, so callers own what they get, as with the reference."""

    def __init__(self, model, cache, ops, shape, n_tokens, dev):
        self.static_rgb = torch.zeros(shape, dtype=torch.float32, device=dev)
        self.static_pos = torch.zeros((shape[0], n_tokens, 6), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):          # warm-up: weight-layout caches, scratch buffers, cuDNN handles
                _pair_body(model, cache, self.static_rgb, self.static_pos, ops)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = _pair_body(model, cache, self.static_rgb, self.static_pos, ops)

    def run(self, rgb, pos):
        self.static_rgb.copy_(rgb, non_blocking=True)
        self.static_pos.copy_(pos, non_blocking=True)
        self.graph.replay()
        z, rel_pose, flows = self.out
        return [t.clone() for t in z], rel_pose.clone(), tuple(t.clone() for t in flows)


@torch.no_grad()
def get_z(model, input, ops, use_graph=False):
    """models/CoPoNeRF.py:159-206. Returns (z list of 4 feature maps, rel_pose (B, 4, 4), flow tuple)."""
    ctx = input["context"]
    dev = next(model.parameters()).device
    rgb = ctx["rgb"]
    B, n_ctxt, H, W, _ = rgb.shape
    if n_ctxt != 2:
        raise ValueError("the per-pair stage is built for two context views")
    model.H, model.W = H, W
    if model.training:
        raise RuntimeError("coponerf_b200 is an inference path: call .eval() (BatchNorm must use running statistics)")
    cache = _state_cache(model)
    n_tokens = (H // 4) * (W // 4)                                  # the finest refined level (ResNet layer2)
    pos = pose_native.positional_encodings_for(ctx["intrinsics"], n_tokens, H, dev)
    if not (use_graph and dev.type == "cuda"):
        return _pair_body(model, cache, rgb.to(dev), pos, ops)
    key = (B, H, W, str(dev))
    g = cache.get("graph")
    if g is None or cache.get("graph_key") != key:
        cache["graph"] = cache["graph_key"] = None                  # free the old pool before capturing anew
        g = cache["graph"] = _PairGraph(model, cache, ops, (B, n_ctxt, H, W, 3), n_tokens, dev)
        cache["graph_key"] = key
    return g.run(rgb, pos)
