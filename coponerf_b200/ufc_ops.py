"""CUDA operator set for coponerf_b200.ufc_native.ufc_forward: every operator is one C-ABI call into
libcoponerf_b200.so (include/coponerf_b200.h). PyTorch only provides device memory, views and elementwise adds."""
import ctypes
import os

import torch

from . import _lib
from .ufc import conv4d_block, linear_attention, ufc_tail


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _c(t):
    return t.detach().to(torch.float32).contiguous()


_ACT = {None: 0, "relu": 1, "gelu": 2}


class CudaOps:
    def __init__(self):
        self.lib = _lib.load()
        self._ws = None    # scratch for the split reductions (stream-ordered reuse)
        self._ws_old = []
        self._wt = {}      # weight tensor -> transposed [K][N] copy for the GEMM (made once per weight)
        # token Linears on the tcgen05 kernel (three fp16 MMAs per product: 1e-5 of fp64 at K = 2304); CPN_UFC_SIMT_LINEAR=1 keeps every Linear on the fp32 CUDA-core GEMM for A/B runs
        self.tc_linear = os.environ.get("CPN_UFC_SIMT_LINEAR", "0") != "1"
        self.tc_min_rows = int(os.environ.get("CPN_UFC_TC_MIN_ROWS", "2048"))
        self.launches = 0  # kernels launched so far (bench.py's gpu_launches)
        self.launches_per_forward = 0

    def _check_dev(self, t):
        if t.device.type != "cuda":
            raise _lib.CpnError("CudaOps run on CUDA tensors only (no CPU fallback)")

    def _transposed(self, w):
        key = (w.data_ptr(), tuple(w.shape), w._version)
        hit = self._wt.get(key)
        if hit is None:
            w = _c(w)
            N, K = w.shape
            wt = torch.empty((K, N), dtype=torch.float32, device=w.device)
            _lib.check(self.lib.cpn_transpose_pq(_p(w), _p(wt), 1, N, K, _st()), "cpn_transpose_pq")
            hit = self._wt[key] = (wt, w)
        return hit[0]

    # ------------------------------------------------------------------ dense layers
    def layernorm(self, x, w, b):
        x = _c(x)
        self._check_dev(x)
        y = torch.empty_like(x)
        self.launches += 1
        _lib.check(self.lib.cpn_layernorm(_p(x), _p(_c(w)), _p(_c(b)), _p(y), x.numel() // x.shape[-1], x.shape[-1], _st()),
                   "cpn_layernorm")
        return y

    def _packed_tc(self, w):
        """W (N, K) -> tensor-core tiles, once per weight tensor (cpn_linear_tc_pack)."""
        key = ("tc", w.data_ptr(), tuple(w.shape), w._version)
        hit = self._wt.get(key)
        if hit is None:
            w = _c(w)
            N, K = w.shape
            packed = torch.empty(self.lib.cpn_linear_tc_packed_bytes(N, K), dtype=torch.uint8, device=w.device)
            _lib.check(self.lib.cpn_linear_tc_pack(_p(w), N, K, _p(packed), _st()), "cpn_linear_tc_pack")
            hit = self._wt[key] = (packed, w)
        return hit[0]

    def linear(self, x, w, b, act=None):
        x = _c(x)
        self._check_dev(x)
        N, K = w.shape
        M = x.numel() // K
        # from 2048 rows on (the 64 x 64 level, every level at 512 x 512): below that the tensor-core kernel has too few CTAs
        # (one 256-row tile each, serial over K) and the split-K CUDA-core GEMM over all SMs is as fast (measured)
        if self.tc_linear and N % 128 == 0 and K % 8 == 0 and M >= self.tc_min_rows:
            y = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
            self.launches += 1
            _lib.check(self.lib.cpn_linear_tc(_p(self._packed_tc(w)), N, K, _p(x), K, _p(_c(b)) if b is not None else None,
                                              _p(y), N, M, _ACT[act], _lib.TC_F16X3, _st()), "cpn_linear_tc")
            return y
        wt, bias = self._transposed(w), _c(b)
        y = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
        nbytes = self.lib.cpn_gemm_simt_splitk_workspace_bytes(M, N, K) if M * N <= 148 * 64 * 64 else 0
        if nbytes:     # few output tiles, long K: split the k range over CTAs (fixed-order reduction)
            ws = self._scratch(nbytes, x.device)
            self.launches += 2
            _lib.check(self.lib.cpn_gemm_simt_splitk(_p(x), K, _p(wt), _p(bias), _p(y), N, M, N, K, _ACT[act], _p(ws),
                                                     ws.numel(), _st()), "cpn_gemm_simt_splitk")
            return y
        self.launches += 1
        _lib.check(self.lib.cpn_gemm_simt(_p(x), K, _p(wt), _p(bias), _p(y), N, M, N, K, _ACT[act], _st()),
                   "cpn_gemm_simt")
        return y

    def matmul(self, a, b):
        """(M, K) @ (K, N), fp32, both row-major."""
        a, b = _c(a), _c(b)
        self._check_dev(a)
        M, K = a.shape
        N = b.shape[1]
        y = torch.empty((M, N), dtype=torch.float32, device=a.device)
        self.launches += 1
        _lib.check(self.lib.cpn_gemm_simt(_p(a), K, _p(b), None, _p(y), N, M, N, K, 0, _st()), "cpn_gemm_simt")
        return y

    def _scratch(self, nbytes, dev):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            if self._ws is not None:
                self._ws_old.append(self._ws)   # a captured graph may still point at it: never hand it back
            self._ws = torch.empty(max(nbytes, 8 << 20), dtype=torch.uint8, device=dev)
        return self._ws

    def matmul_tn(self, a, b, bias=None, act=None):
        """a^T @ b (+ bias, activation) for a (L, Mi), b (L, Nj); rows may be strided views of wider matrices."""
        self._check_dev(a)
        if a.dtype != torch.float32 or b.dtype != torch.float32 or a.stride(1) != 1 or b.stride(1) != 1:
            a, b = _c(a), _c(b)
        L, Mi = a.shape
        Nj = b.shape[1]
        y = torch.empty((Mi, Nj), dtype=torch.float32, device=a.device)
        nbytes = self.lib.cpn_gemm_tn_workspace_bytes(Mi, Nj, L)
        ws = self._scratch(nbytes, a.device)
        self.launches += 2
        _lib.check(self.lib.cpn_gemm_tn(_p(a), a.stride(0), _p(b), b.stride(0), _p(_c(bias)) if bias is not None else None,
                                        _p(y), Nj, Mi, Nj, L, _ACT[act], _p(ws), ws.numel(), _st()), "cpn_gemm_tn")
        return y

    def linear_skinny(self, x, w, b, act=None):
        """act(x W^T + b) for a few rows and a very long K, W as stored (N, K)."""
        x, w, b = _c(x), _c(w), _c(b)
        self._check_dev(x)
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        for m0 in range(0, M, 8):
            m = min(8, M - m0)
            nbytes = self.lib.cpn_linear_skinny_workspace_bytes(m, N, K)
            ws = self._scratch(nbytes, x.device)
            self.launches += 2
            _lib.check(self.lib.cpn_linear_skinny(_p(x[m0:m0 + m]), _p(w), _p(b), _p(y[m0:m0 + m]), m, N, K, _ACT[act],
                                                  _p(ws), ws.numel(), _st()), "cpn_linear_skinny")
        return y

    # ------------------------------------------------------------------ pose features (models/backbone.py)
    def dual_softmax(self, c):
        c = _c(c)
        self._check_dev(c)
        B, L, _ = c.shape
        P = torch.empty_like(c)
        nbytes = self.lib.cpn_dual_softmax_workspace_bytes(B, L)
        ws = self._scratch(nbytes, c.device)
        self.launches += 4
        _lib.check(self.lib.cpn_dual_softmax(_p(c), _p(P), B, L, _p(ws), ws.numel(), _st()), "cpn_dual_softmax")
        return P

    def pose_head(self, h0, sd):
        h0 = _c(h0)
        self._check_dev(h0)
        B = h0.shape[0]
        out = torch.empty((B, 4, 4), dtype=torch.float32, device=h0.device)
        a = _lib.PoseHeadArgs()
        a.B = B
        a.h0, a.rel_pose = h0.data_ptr(), out.data_ptr()
        keep = []
        for field, key in (("w2", "pose_regressor.2.weight"), ("b2", "pose_regressor.2.bias"),
                           ("w4", "pose_regressor.4.weight"), ("b4", "pose_regressor.4.bias"),
                           ("rw1", "rotation_regressor.1.weight"), ("rb1", "rotation_regressor.1.bias"),
                           ("rw3", "rotation_regressor.3.weight"), ("rb3", "rotation_regressor.3.bias"),
                           ("rw5", "rotation_regressor.5.weight"), ("rb5", "rotation_regressor.5.bias"),
                           ("tw1", "translation_regressor.1.weight"), ("tb1", "translation_regressor.1.bias"),
                           ("tw3", "translation_regressor.3.weight"), ("tb3", "translation_regressor.3.bias"),
                           ("tw5", "translation_regressor.5.weight"), ("tb5", "translation_regressor.5.bias")):
            t = _c(sd[key])
            keep.append(t)
            setattr(a, field, t.data_ptr())
        self.launches += 1
        _lib.check(self.lib.cpn_pose_head(ctypes.byref(a), _st()), "cpn_pose_head")
        return out

    # ------------------------------------------------------------------ correlation volume <-> tokens
    def corr_to_tokens(self, corr, n):
        corr = _c(corr)
        self._check_dev(corr)
        B, H, hs, _, q, _ = corr.shape
        tok = torch.empty((B, n * n, H * q * q), dtype=torch.float32, device=corr.device)
        self.launches += 1
        _lib.check(self.lib.cpn_corr_to_tokens(_p(corr), _p(tok), B, H, hs, q, n, H * q * q, 0, _st()), "cpn_corr_to_tokens")
        return tok

    def tokens_to_corr(self, tok, n, H, hs):
        tok = _c(tok)
        self._check_dev(tok)
        B, L, CH = tok.shape
        q = int(round((CH // H) ** 0.5))
        corr = torch.empty((B, H, hs, hs, q, q), dtype=torch.float32, device=tok.device)
        self.launches += 1
        _lib.check(self.lib.cpn_tokens_to_corr(_p(tok), _p(corr), B, H, hs, q, n, _st()), "cpn_tokens_to_corr")
        return corr

    def transpose4d(self, corr):
        corr = _c(corr)
        self._check_dev(corr)
        B, H, hs, ws, ht, wt = corr.shape
        out = torch.empty((B, H, ht, wt, hs, ws), dtype=torch.float32, device=corr.device)
        self.launches += 1
        _lib.check(self.lib.cpn_transpose_pq(_p(corr), _p(out), B * H, hs * ws, ht * wt, _st()), "cpn_transpose_pq")
        return out

    def encoder4d(self, x, blocks, stride, pad):
        for p in blocks:
            self.launches += 2      # conv4d_kernel + gn_relu_kernel
            x = conv4d_block(x, p["wq"], p["bq"], p["ws"], p["bs"], p["gamma"], p["beta"], stride, pad)
        return x

    def correlation(self, src, trg, n):
        src, trg = _c(src), _c(trg)
        self._check_dev(src)
        B, L, C = src.shape
        out = torch.empty((B, 1, n, n, n, n), dtype=torch.float32, device=src.device)
        nbytes = self.lib.cpn_correlation_workspace_bytes(B, L, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=src.device)
        self.launches += 3 + B      # two normalisations, one transpose, one GEMM per pair
        _lib.check(self.lib.cpn_correlation(_p(src), _p(trg), _p(out), B, L, C, _p(ws), nbytes, _st()), "cpn_correlation")
        return out

    # ------------------------------------------------------------------ attention
    def linear_attention(self, q, k, v):
        self.launches += 2 + (2 if k.shape[1] > 128 else 0)     # kv, (two partial reductions), out
        return linear_attention(q, k, v)

    def cross_attention(self, corr, src_v, trg_v):
        corr, src_v, trg_v = _c(corr), _c(src_v), _c(trg_v)
        self._check_dev(corr)
        B, H = corr.shape[:2]
        S, T = corr.shape[2] * corr.shape[3], corr.shape[4] * corr.shape[5]
        D = src_v.shape[-1]
        src_attn = torch.empty((B, S, H * D), dtype=torch.float32, device=corr.device)
        trg_attn = torch.empty((B, T, H * D), dtype=torch.float32, device=corr.device)
        self.launches += 2
        _lib.check(self.lib.cpn_cross_attention(_p(corr), _p(src_v), _p(trg_v), _p(src_attn), _p(trg_attn), B, H, S, T, D,
                                                _st()), "cpn_cross_attention")
        return src_attn, trg_attn

    # ------------------------------------------------------------------ token maps
    def dwconv_gelu(self, x, w, b, n):
        x = _c(x)
        self._check_dev(x)
        B, L, C = x.shape
        y = torch.empty_like(x)
        self.launches += 1
        _lib.check(self.lib.cpn_dwconv_gelu(_p(x), _p(_c(w)), _p(_c(b)), _p(y), B, n, C, _st()), "cpn_dwconv_gelu")
        return y

    def _resample(self, x, n, arg, m, mode):
        x = _c(x)
        self._check_dev(x)
        B, L, C = x.shape
        y = torch.empty((B, m * m, C), dtype=torch.float32, device=x.device)
        self.launches += 1
        _lib.check(self.lib.cpn_resample_tokens(_p(x), _p(y), B, n, arg, C, mode, _st()), "cpn_resample_tokens")
        return y

    def upsample_tokens(self, x, n_out):
        n = int(round(x.shape[1] ** 0.5))
        return self._resample(x, n, n_out, n_out, 0)

    def avgpool_tokens(self, x, n, pool):
        return x if pool == 1 else self._resample(x, n, pool, n // pool, 1)

    def repeat_tokens(self, x, hs, pool):
        return x if pool == 1 else self._resample(x, hs, pool, hs * pool, 2)

    def tail(self, src, trg, sizes, out):
        self.launches += 12 + src[0].shape[0] + 2     # 6 normalisations, 6 upsample-packs, one GEMM per pair, 2 soft-argmax
        return ufc_tail(src, trg, sizes, out)
