"""CUDA operator set for coponerf_b200.ufc_native.ufc_forward: every operator is one C-ABI call into
libcoponerf_b200.so (include/coponerf_b200.h). PyTorch only provides device memory, views and elementwise adds."""
import ctypes

import torch

from . import _lib
from .ufc import conv4d_block, linear_attention, ufc_tail


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _c(t):
    return t.detach().to(torch.float32).contiguous()


class CudaOps:
    def __init__(self):
        self.lib = _lib.load()
        self._wt = {}      # weight tensor -> transposed [K][N] copy for the GEMM (made once per weight)
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

    def linear(self, x, w, b, act=None):
        x = _c(x)
        self._check_dev(x)
        N, K = w.shape
        M = x.numel() // K
        wt, bias = self._transposed(w), _c(b)
        y = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
        self.launches += 1
        _lib.check(self.lib.cpn_gemm_simt(_p(x), K, _p(wt), _p(bias), _p(y), N, M, N, K, int(act == "relu"), _st()),
                   "cpn_gemm_simt")
        return y

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
