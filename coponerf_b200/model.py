"""Drop-in entry point: a CoPoNeRF module with the reference's constructor, state_dict layout for the
render path, get_z()/forward() signatures and output dictionary (models/CoPoNeRF.py:19-104,159-576),
whose per-ray stage runs on the sm_100a kernels of libcoponerf_b200.so.

get_z() is standalone too (coponerf_b200/pair_stage.py): the ResNet-34 image encoder stays PyTorch/cuDNN
(BASELINE.json: "models/backbone.py and the train/test drivers stay"), the cost aggregation and the pose
features / pose head run on the sm_100a operators. The module carries all 744 state_dict keys of the reference
model, so `load_state_dict(torch.load(ckpt)['model'], strict=False)` (test.py:143) works unchanged.
"""
import torch
import torch.nn as nn

from . import pair_stage
from .render import RenderEngine


class _ResnetBlockFC(nn.Module):
    """Parameter container with the names of models/lightfield.py:10-62."""

    def __init__(self, d):
        super().__init__()
        self.fc_0 = nn.Linear(d, d)
        self.fc_1 = nn.Linear(d, d)


class _ResnetFC(nn.Module):
    """Parameter container with the names of models/lightfield.py:65-129 (d_in 18, d_latent 832, 3 blocks)."""

    def __init__(self, d_in, d_latent, d_hidden, d_out=3, n_blocks=3):
        super().__init__()
        self.lin_in = nn.Linear(d_in, d_hidden)
        self.lin_out = nn.Linear(d_hidden, d_out)
        self.blocks = nn.ModuleList([_ResnetBlockFC(d_hidden) for _ in range(n_blocks)])
        self.lin_z = nn.ModuleList([nn.Linear(d_latent, d_hidden) for _ in range(n_blocks)])


def pose_inverse_4x4(mat):
    """utils_training/utils.py:111-138."""
    out = torch.zeros_like(mat)
    r_inv = mat[:, :3, :3].transpose(-1, -2)
    out[:, :3, :3] = r_inv
    out[:, :3, 3] = (-r_inv @ mat[:, :3, 3:])[..., 0]
    out[:, 3, 3] = 1
    return out


class CoPoNeRF(nn.Module):
    """models/CoPoNeRF.py:19. Only n_view == 2 is supported (the value both reference drivers pass)."""

    def __init__(self, n_view=1, npoints=64, num_hidden_units_phi=128, chunk_rays=2048, lanes=2):
        super().__init__()
        self.n_view = n_view
        self.npoints = npoints if npoints else 64
        latent = 256 * 3 + 64
        hidden = 128
        if num_hidden_units_phi != hidden:
            raise ValueError("the sm_100a render path is built for num_hidden_units_phi == 128")
        # per-pair stage: encoder, conv_map, feature_cost_aggregation, cross_attention, pose / rotation / translation
        # regressors (models/CoPoNeRF.py:31-69) as parameter containers with the reference's names
        pair_stage.PairStage.build_into(self)
        # render-path parameters, same names and shapes as models/CoPoNeRF.py:71-104
        self.query_encode_latent = nn.Conv2d(latent + 3, latent, 1)
        self.query_encode_latent_2 = nn.Conv2d(latent, latent // 2, 1)
        self.corr_embed = nn.Conv2d(4096, latent, 1)                       # unused by forward
        half = latent // 2
        self.latent_value = nn.Conv2d(half * n_view, half, 1)
        self.key_map = nn.Conv2d(half * n_view, hidden, 1)
        self.key_map_2 = nn.Conv2d(hidden, hidden, 1)
        self.query_embed = nn.Conv2d(16, hidden, 1)
        self.query_embed_2 = nn.Conv2d(hidden, hidden, 1)
        self.latent_avg_query = nn.Conv2d(9 + 16, hidden, 1)               # unused by forward
        self.latent_avg_query_2 = nn.Conv2d(hidden, hidden, 1)             # unused by forward
        self.latent_avg_key = nn.Conv2d(half, hidden, 1)                   # unused by forward
        self.latent_avg_key_2 = nn.Conv2d(hidden, hidden, 1)               # unused by forward
        self.query_repeat_embed = nn.Conv2d(16 + 128, hidden, 1)
        self.query_repeat_embed_2 = nn.Conv2d(hidden, hidden, 1)
        self.latent_avg_repeat_query = nn.Conv2d(9 + 16 + 128, hidden, 1)  # unused by forward
        self.latent_avg_repeat_query_2 = nn.Conv2d(hidden, hidden, 1)      # unused by forward
        self.encode_latent = nn.Conv1d(half, 128, 1)
        self.phi = _ResnetFC(n_view * 9, half * n_view, hidden)
        self.chunk_rays = chunk_rays
        self.lanes = lanes
        self.graph_get_z = True         # get_z(): replay the per-pair stage (~800 short kernels) from a CUDA graph: same
                                        # kernels, same bits; keeps the stage at its GPU time when the host is slow
        self._ufc_ops = None
        self.pixel_val_on_host = True   # the reference returns out['pixel_val'] as a CPU tensor (CoPoNeRF.py:490)
        self._engine = None
        self._engine_version = None
        self.H = self.W = None

    # ------------------------------------------------------------------ engine management
    def _param_list(self):
        """All parameters, cached: walking the module tree costs ~1 ms per call, the version check itself 0.2 ms.
        The cache is dropped by load_state_dict() and by .to() / .cuda(); call refresh_parameters() after replacing a
        Parameter OBJECT by hand (in-place updates are seen through the tensors' version counters)."""
        pl = self.__dict__.get("_plist")
        if pl is None:
            # buffers too: fold_encoder reads the BatchNorm running statistics (pair_stage.py)
            pl = self.__dict__["_plist"] = list(self.parameters()) + list(self.buffers())
        return pl

    def refresh_parameters(self):
        self.__dict__["_plist"] = None

    def _apply(self, fn, *args, **kwargs):
        self.refresh_parameters()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.refresh_parameters()
        return super().load_state_dict(*args, **kwargs)

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self._param_list())

    def engine(self):
        """RenderEngine holding the packed weights; repacked when parameters change or move."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("coponerf_b200.CoPoNeRF renders on CUDA only: call .cuda() first (no CPU fallback)")
        ver = self._weights_version()
        if self._engine is None or self._engine_version != ver or self._engine.device != dev:
            self._engine = RenderEngine(self.state_dict(), device=dev, chunk_rays=self.chunk_rays, lanes=self.lanes)
            self._engine_version = ver
        return self._engine

    # ------------------------------------------------------------------ reference API
    def get_z(self, input, val=False):
        """models/CoPoNeRF.py:159-206: image features, estimated relative pose, correspondence fields."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("coponerf_b200.CoPoNeRF.get_z() runs on CUDA only: call .cuda() first (no CPU fallback)")
        if self._ufc_ops is None:
            from .ufc_ops import CudaOps
            self._ufc_ops = CudaOps()
        with torch.cuda.device(dev):
            return pair_stage.get_z(self, input, self._ufc_ops, use_graph=self.graph_get_z)

    def render_pairs(self, inputs, val=False):
        """forward(input, val=val) for a sequence of pairs, as a generator: get_z() of pair k + 1 runs on a second stream while
        pair k renders (pipeline.py). Same outputs as calling forward() pair by pair."""
        from .pipeline import render_pairs
        return render_pairs(self, inputs, val=val)

    @torch.no_grad()
    def forward(self, input, z=None, rel_pose=None, val=False, flow=None, debug=False):
        """models/CoPoNeRF.py:208-576 (inference; the sm_100a path has no backward)."""
        if self.n_view != 2:
            raise ValueError("the sm_100a render path supports n_view == 2")
        ctx, qry = input["context"], input["query"]
        if z is None:
            z, rel_pose, flow = self.get_z(input)
        if self.H is None or self.W is None:
            raise RuntimeError("model.H / model.W are unset: get_z() sets them, or assign them by hand")
        eng = self.engine()
        st = eng.prepare_pair(input, z, rel_pose, flow, self.H, self.W, val)
        B = ctx["rgb"].shape[0]
        uv = qry["uv"]
        N = uv.shape[2]
        o = eng.render_rays(st, uv.reshape(B, N, 2), self.npoints)
        out = {"flow": flow, "z": z, "uv": uv, "coords": o["coords"]}
        for k in ("rgb", "valid_mask", "depth_ray", "at_wt", "at_wt_max", "T_to_C1_pts", "T_to_C2_pts",
                  "C2_pts_to_C1"):
            out[k] = o[k]
        out["at_wts"] = [o["at_wt"]]
        if self.pixel_val_on_host:   # pinned staging (torch's caching host allocator) instead of a pageable .cpu()
            pv = torch.empty(o["pixel_val"].shape, dtype=torch.float32, pin_memory=True)
            pv.copy_(o["pixel_val"], non_blocking=True)
            torch.cuda.current_stream(o["pixel_val"].device).synchronize()   # the stream of the model's device, not of the current one
            out["pixel_val"] = pv
        else:
            out["pixel_val"] = o["pixel_val"]
        out["mask_c2"] = o["mask_c2"].view(torch.bool)
        out["matchability_cycle_mask"] = o["matchability_cycle_mask"].view(torch.bool)
        c = st.consts
        out["rel_pose"] = rel_pose
        out["rel_pose_flip"] = c[:, 219:235].reshape(B, 4, 4)
        out["gt_rel_pose"] = c[:, 235:251].reshape(B, 4, 4)
        out["gt_rel_pose_flip"] = c[:, 251:267].reshape(B, 4, 4)
        return out
