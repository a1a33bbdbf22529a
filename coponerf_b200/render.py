"""Host side of the per-ray render stage: owns the packed weights and per-pair state on one GPU and
drives the C-ABI (include/coponerf_b200.h). PyTorch is used for device memory and streams only.

Mirrors what models/CoPoNeRF.py:208-576 does around the math: it takes the same input dict, z list,
rel_pose and flow tuple and returns the same output dict.
"""
import ctypes
from collections import OrderedDict

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class PairState:
    """Everything that is identical for every chunk of rays of one batch of stereo pairs."""
    __slots__ = ("feat", "feat_shape", "consts", "up_flow2", "mask_padded2", "B", "H", "W", "flow_h", "val")


class RenderEngine:
    """Packed render-path weights + workspace on one CUDA device."""

    def __init__(self, state_dict, device=None, chunk_rays=2048, lanes=2):
        if not torch.cuda.is_available():
            raise _lib.CpnError("coponerf_b200 needs a CUDA device: the render path has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.chunk_rays = int(chunk_rays)
        self.lanes = int(lanes)   # chunks in flight on the library's internal streams
        self._workspace = None
        self._interval = {}
        self._pair_cache = OrderedDict()
        self._feat_cache = OrderedDict()
        self.flags = 0
        self.load_weights(state_dict)

    # ------------------------------------------------------------------ weights
    def load_weights(self, state_dict):
        """cpn_pack_weights: state_dict tensors -> kernel layouts (done once per checkpoint)."""
        lib = self.lib
        parts = []
        for i, name in enumerate(_lib.weight_names()):
            if name not in state_dict:
                raise KeyError(f"state_dict is missing render-path key {name!r}")
            t = state_dict[name].detach().reshape(-1).to(torch.float32)
            if t.numel() != lib.cpn_weight_numel(i):
                raise ValueError(f"{name}: expected {lib.cpn_weight_numel(i)} elements, got {t.numel()}")
            parts.append(t.cpu())
        raw = torch.cat(parts).to(self.device)
        assert raw.numel() == lib.cpn_raw_weights_floats()
        self.weights = torch.empty(lib.cpn_packed_weights_bytes(), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.cpn_pack_weights(_ptr(raw), _ptr(self.weights), _stream()), "cpn_pack_weights")
            torch.cuda.current_stream().synchronize()  # `raw` is freed on return

    # ------------------------------------------------------------------ per-pair state
    def pack_features(self, z):
        """NCHW feature maps -> channels-last copies the gather kernel reads (cached per z list)."""
        key = tuple((t.data_ptr(), tuple(t.shape), t._version) for t in z)
        hit = self._feat_cache.get(key)
        if hit is not None:
            return hit
        feats = []
        for t in z:
            t = _f32c(t, self.device)
            n, c, h, w = t.shape
            out = torch.empty((n, h, w, c), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.cpn_pack_features(_ptr(t), _ptr(out), n, c, h, w, _stream()), "cpn_pack_features")
            feats.append(out)
        self._feat_cache[key] = (feats, list(z))  # keep z alive so data_ptr keys stay unique
        while len(self._feat_cache) > 4:
            self._feat_cache.popitem(last=False)
        return self._feat_cache[key]

    def prepare_pair(self, inp, z, rel_pose, flow, H, W, val):
        ctx, qry = inp["context"], inp["query"]
        B = ctx["rgb"].shape[0]
        if len(z) != _lib.N_LEVELS:
            raise ValueError(f"expected {_lib.N_LEVELS} feature maps, got {len(z)}")
        dev = self.device
        st = PairState()
        st.B, st.H, st.W, st.val = B, int(H), int(W), bool(val)
        with torch.cuda.device(dev):
            st.feat, _ = self.pack_features(z)
            st.feat_shape = [tuple(f.shape) for f in st.feat]
            for f in st.feat:
                if f.shape[0] != 2 * B:
                    raise ValueError("feature maps must have 2*B images (two context views per pair)")
            c2w = _f32c(ctx["cam2world"], dev)
            Kc = _f32c(ctx["intrinsics"], dev)
            qc2w = _f32c(qry["cam2world"], dev).reshape(B, 4, 4)
            Kq = _f32c(qry["intrinsics"], dev).reshape(B, 4, 4)
            rel = _f32c(rel_pose, dev)
            if c2w.shape != (B, 2, 4, 4) or Kc.shape != (B, 2, 4, 4) or rel.shape != (B, 4, 4):
                raise ValueError("context cam2world/intrinsics must be (B,2,4,4) and rel_pose (B,4,4)")
            st.consts = torch.empty((B, _lib.PAIR_CONSTS_FLOATS), dtype=torch.float32, device=dev)
            _lib.check(self.lib.cpn_pair_setup(_ptr(c2w), _ptr(Kc), _ptr(qc2w), _ptr(Kq), _ptr(rel), B, st.H,
                                               int(st.val), _ptr(st.consts), _stream()), "cpn_pair_setup")
            f0, f1 = _f32c(flow[0], dev), _f32c(flow[1], dev)
            fh, fw = f1.shape[-2:]
            st.flow_h = int(fh)
            st.up_flow2 = torch.empty((B, 2, 256, 256), dtype=torch.float32, device=dev)
            st.mask_padded2 = torch.empty((B, 256, 256), dtype=torch.uint8, device=dev)
            rgb_w = int(ctx["rgb"].shape[-2])
            _lib.check(self.lib.cpn_pair_prologue(_ptr(f0), _ptr(f1), B, int(fh), int(fw), rgb_w, _ptr(st.up_flow2),
                                                  _ptr(st.mask_padded2), _stream()), "cpn_pair_prologue")
        return st

    # ------------------------------------------------------------------ rays
    def _get_workspace(self, nbytes):
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace

    def _get_interval(self, S):
        if S not in self._interval:
            self._interval[S] = torch.linspace(0, 1, S).to(self.device)  # built on the host like the reference's
        return self._interval[S]

    def render_rays(self, st, uv, S, out=None):
        """uv (B, N, 2) device fp32 -> dict of device tensors in the reference's shapes."""
        dev = self.device
        B = st.B
        uv = _f32c(uv, dev).reshape(B, -1, 2)
        N = uv.shape[1]
        chunk = max(1, min(self.chunk_rays, N)) if N > 0 else 1
        with torch.cuda.device(dev):
            f32 = dict(dtype=torch.float32, device=dev)
            o = out if out is not None else {
                "rgb": torch.empty((B, 1, N, 3), **f32),
                "valid_mask": torch.empty((B, N, 1), **f32),
                "depth_ray": torch.empty((B, N, 1), **f32),
                "at_wt": torch.empty((2 * B, N, S), **f32),
                "at_wt_max": torch.empty((2 * B, N, 1), dtype=torch.int64, device=dev),
                "pixel_val": torch.empty((2 * B, N, S, 2), **f32),
                "coords": torch.empty((2 * B, N, 9), **f32),
                "T_to_C1_pts": torch.empty((B, N, 2), **f32),
                "T_to_C2_pts": torch.empty((B, N, 2), **f32),
                "C2_pts_to_C1": torch.empty((B, N, 2), **f32),
                "mask_c2": torch.empty((B, N), dtype=torch.uint8, device=dev),
                "matchability_cycle_mask": torch.empty((B, N), dtype=torch.uint8, device=dev),
            }
            if N == 0:
                return o
            lanes = max(1, min(self.lanes, 4, -(-N // chunk)))
            ws_bytes = self.lib.cpn_render_workspace_bytes_for(B, N, chunk, S, lanes, self.flags)
            ws = self._get_workspace(ws_bytes)
            a = _lib.RenderArgs()
            a.B, a.N, a.S, a.H, a.W = B, N, S, st.H, st.W
            a.flow_h, a.chunk_rays, a.flags, a.lanes = st.flow_h, chunk, self.flags, lanes
            for l, f in enumerate(st.feat):
                a.feat[l] = f.data_ptr()
                a.feat_h[l], a.feat_w[l], a.feat_c[l] = f.shape[1], f.shape[2], f.shape[3]
            a.pair_consts = st.consts.data_ptr()
            a.uv = uv.data_ptr()
            a.interval = self._get_interval(S).data_ptr()
            a.weights = self.weights.data_ptr()
            a.up_flow2 = st.up_flow2.data_ptr()
            a.mask_padded2 = st.mask_padded2.data_ptr()
            for k in ("rgb", "valid_mask", "depth_ray", "at_wt", "at_wt_max", "pixel_val", "coords", "T_to_C1_pts",
                      "T_to_C2_pts", "C2_pts_to_C1", "mask_c2", "matchability_cycle_mask"):
                setattr(a, k, o[k].data_ptr())
            a.workspace = ws.data_ptr()
            a.workspace_bytes = ws.numel()
            _lib.check(self.lib.cpn_render_rays(ctypes.byref(a), _stream()), "cpn_render_rays")
            self.last_launch_count = self.lib.cpn_render_launch_count(ctypes.byref(a))
        return o
