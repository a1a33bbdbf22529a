"""ctypes binding of libcoponerf_b200.so (the C-ABI declared in include/coponerf_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcoponerf_b200.so")

N_LEVELS = 4
FEAT_DIM = 832
LATENT = 416
HIDDEN = 128
PAIR_CONSTS_FLOATS = 320
FLAG_SIMT_ONLY = 1
FLAG_F16X3 = 2
FLAG_NO_FOLD = 4
FLAG_EARLY_V = 8
FLAG_NO_BILINEAR = 16
FLAG_NO_GFOLD = 32
FLAG_FULL_H1 = 64
TC_F16X3 = 4
TC_CLUSTER = 8
TC_PAIR = 16
TC_OUT_CB16 = 32
TC_OUT_KG = 128
TC_NO_PERSIST = 256
TC_PPAIR = 512
TC_A_IMAGE3 = 1024
TC_OUT_IMAGE3 = 2048
TC_PERSIST2 = 4096
TC_WS = 8192
ACT_CHUNK3_BYTES = 12288
TC_A_IMAGE = 1
TC_OUT_IMAGE = 2
ACT_CHUNK_BYTES = 16384
ABI_VERSION = 200        # cpn_version() of the library this binding was written against

c_float_p = ctypes.POINTER(ctypes.c_float)


class RenderArgs(ctypes.Structure):
    """cpn_render_args (include/coponerf_b200.h)."""
    _fields_ = [
        ("B", ctypes.c_int32), ("N", ctypes.c_int32), ("S", ctypes.c_int32),
        ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("flow_h", ctypes.c_int32),
        ("chunk_rays", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("lanes", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("feat", ctypes.c_void_p * N_LEVELS),
        ("feat_h", ctypes.c_int32 * N_LEVELS), ("feat_w", ctypes.c_int32 * N_LEVELS),
        ("feat_c", ctypes.c_int32 * N_LEVELS),
        ("pair_consts", ctypes.c_void_p), ("uv", ctypes.c_void_p), ("interval", ctypes.c_void_p),
        ("weights", ctypes.c_void_p), ("up_flow2", ctypes.c_void_p), ("mask_padded2", ctypes.c_void_p),
        ("rgb", ctypes.c_void_p), ("valid_mask", ctypes.c_void_p), ("depth_ray", ctypes.c_void_p),
        ("at_wt", ctypes.c_void_p), ("at_wt_max", ctypes.c_void_p), ("pixel_val", ctypes.c_void_p),
        ("coords", ctypes.c_void_p), ("T_to_C1_pts", ctypes.c_void_p), ("T_to_C2_pts", ctypes.c_void_p),
        ("C2_pts_to_C1", ctypes.c_void_p), ("mask_c2", ctypes.c_void_p),
        ("matchability_cycle_mask", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


class UfcTailArgs(ctypes.Structure):
    """cpn_ufc_tail_args (include/coponerf_b200.h)."""
    _fields_ = [("B", ctypes.c_int32), ("C", ctypes.c_int32), ("out", ctypes.c_int32), ("sizes", ctypes.c_int32 * 3),
                ("src", ctypes.c_void_p * 3), ("trg", ctypes.c_void_p * 3), ("lin", ctypes.c_void_p),
                ("c", ctypes.c_void_p), ("flow", ctypes.c_void_p), ("flow_flip", ctypes.c_void_p),
                ("flow_t_to_s", ctypes.c_void_p), ("flow_s_to_t", ctypes.c_void_p),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t)]


class Conv4dArgs(ctypes.Structure):
    """cpn_conv4d_args (include/coponerf_b200.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "Ci", "Co", "Hq", "Hs", "k", "stride", "pad", "norm_relu", "reserved")] + \
               [(n, ctypes.c_void_p) for n in ("x", "wq", "bq", "ws", "bs", "gamma", "beta", "y", "workspace")] + \
               [("workspace_bytes", ctypes.c_size_t)]


class PoseHeadArgs(ctypes.Structure):
    """cpn_pose_head_args (include/coponerf_b200.h)."""
    _fields_ = [("B", ctypes.c_int32), ("reserved", ctypes.c_int32)] + \
               [(n, ctypes.c_void_p) for n in ("h0", "w2", "b2", "w4", "b4", "rw1", "rb1", "rw3", "rb3", "rw5", "rb5",
                                               "tw1", "tb1", "tw3", "tb3", "tw5", "tb5", "rel_pose")]


# symbol -> (restype, argtypes); every entry point include/coponerf_b200.h declares
SIGNATURES = {
    "cpn_version": (ctypes.c_int, []),
    "cpn_shutdown": (ctypes.c_int, []),
    "cpn_last_error": (ctypes.c_char_p, []),
    "cpn_sizeof_render_args": (ctypes.c_size_t, []),
    "cpn_prof_begin": (ctypes.c_int, [ctypes.c_int]),
    "cpn_prof_end": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
    "cpn_packed_weights_bytes": (ctypes.c_size_t, []),
    "cpn_raw_weights_floats": (ctypes.c_size_t, []),
    "cpn_n_weight_tensors": (ctypes.c_int, []),
    "cpn_weight_name": (ctypes.c_char_p, [ctypes.c_int]),
    "cpn_weight_numel": (ctypes.c_size_t, [ctypes.c_int]),
    "cpn_pack_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "cpn_pack_features": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p]),
    "cpn_pair_setup": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p]),
    "cpn_pair_prologue": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 +
                          [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "cpn_render_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "cpn_render_workspace_bytes_for": (ctypes.c_size_t, [ctypes.c_int] * 6),
    "cpn_render_rays": (ctypes.c_int, [ctypes.POINTER(RenderArgs), ctypes.c_void_p]),
    "cpn_render_launch_count": (ctypes.c_int, [ctypes.POINTER(RenderArgs)]),
    "cpn_gemm_tc_rowdot": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_void_p]),
    "cpn_linear_tc_packed_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "cpn_linear_tc_pack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "cpn_linear_tc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "cpn_gather_rows_taps_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "cpn_gather_rows": (ctypes.c_int, [ctypes.POINTER(RenderArgs), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    "cpn_gemm_tc_trace": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "cpn_gemm_tc_kg": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p]),
    "cpn_ufc_tail_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                       ctypes.POINTER(ctypes.c_int32)]),
    "cpn_ufc_tail": (ctypes.c_int, [ctypes.POINTER(UfcTailArgs), ctypes.c_void_p]),
    "cpn_conv4d_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 6),
    "cpn_conv4d": (ctypes.c_int, [ctypes.POINTER(Conv4dArgs), ctypes.c_void_p]),
    "cpn_linear_attention_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 3),
    "cpn_linear_attention": (ctypes.c_int, [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 +
                             [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "cpn_layernorm": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 2 + [ctypes.c_void_p]),
    "cpn_corr_to_tokens": (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 7 + [ctypes.c_void_p]),
    "cpn_tokens_to_corr": (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 5 + [ctypes.c_void_p]),
    "cpn_transpose_pq": (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 3 + [ctypes.c_void_p]),
    "cpn_dwconv_gelu": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3 + [ctypes.c_void_p]),
    "cpn_resample_tokens": (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 5 + [ctypes.c_void_p]),
    "cpn_cross_attention": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int] * 5 + [ctypes.c_void_p]),
    "cpn_correlation_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 3),
    "cpn_correlation": (ctypes.c_int, [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_size_t,
                                                                                   ctypes.c_void_p]),
    "cpn_dual_softmax_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 2),
    "cpn_dual_softmax": (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 2 + [ctypes.c_void_p, ctypes.c_size_t,
                                                                                    ctypes.c_void_p]),
    "cpn_gemm_tn_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 3),
    "cpn_gemm_tn": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_size_t,
                                                                           ctypes.c_void_p]),
    "cpn_linear_skinny_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 3),
    "cpn_linear_skinny": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_size_t,
                                                                                    ctypes.c_void_p]),
    "cpn_pose_head": (ctypes.c_int, [ctypes.POINTER(PoseHeadArgs), ctypes.c_void_p]),
    "cpn_gemm_simt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p]),
    "cpn_gemm_simt_splitk_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 3),
    "cpn_gemm_simt_splitk": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "cpn_gemm_tc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p]),
}

_lib = None


class CpnError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and os.environ.get("CPN_NO_AUTOBUILD") != "1":
        # the library is built in-tree by __graft_entry__.build(); if a checkout arrives without it and nvcc is there,
        # build it now (about a minute) instead of failing every caller
        try:
            from .build import build_library_locked
            print(f"coponerf_b200: {LIB_PATH} is missing, building it with nvcc ...", flush=True)
            build_library_locked()      # inter-process lock + atomic rename: safe under torchrun
        except Exception as e:      # no nvcc, compile error: fall through to the loud failure below
            print(f"coponerf_b200: automatic build failed: {e}", flush=True)
    if not os.path.exists(LIB_PATH):
        raise CpnError(f"{LIB_PATH} is missing: run `python -m coponerf_b200.build` (or __graft_entry__.build()); "
                       "there is no CPU or PyTorch fallback for the render path")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.cpn_sizeof_render_args() != ctypes.sizeof(RenderArgs) or lib.cpn_version() != ABI_VERSION:
        raise CpnError(f"{LIB_PATH} is stale (ABI version {lib.cpn_version()}, this binding expects {ABI_VERSION}): "
                       "rebuild with `python -m coponerf_b200.build`")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().cpn_last_error()
        raise CpnError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


def weight_names():
    lib = load()
    return [lib.cpn_weight_name(i).decode() for i in range(lib.cpn_n_weight_tensors())]
