// Small operators of the cost aggregation (models/aggregation.py) that sit between its GEMMs, 4-D convolutions and
// attention kernels. Tokens are (B, L = n*n, C) row-major, correlation volumes (B, H, hs, ws, ht, wt).
// Each replaces a chain of einops rearranges + F.interpolate / nn.LayerNorm / DWConv / softmax calls without
// materialising the rearranged copies (44 % of the reference's UFC time is aten::copy_ from those rearranges).
#include <math.h>
#include <stdlib.h>
#include "cpn_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// bilinear source coordinates, align_corners=True (ATen area_pixel_compute_source_index)
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_ac(int dst, int in, int out) {
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  const float src = scale * (float)dst;
  Lerp r;
  r.i0 = (int)src;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

// ---- nn.LayerNorm over the last dim (eps 1e-5), one warp per token ------------------------------------------------
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                 float* __restrict__ y, int tokens, int C) {
  const int tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  const float* xr = x + (size_t)tok * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum_f(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = xr[c] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum_f(v) / (float)C + 1e-5f);
  for (int c = lane; c < C; c += 32) y[(size_t)tok * C + c] = (xr[c] - mean) * rstd * g[c] + b[c];
}

// ---- corr (B, H, hs, hs, q, q) -> tokens (B, n*n, H*q*q): 'B H Hs Ws Ht Wt -> B (H Ht Wt) Hs Ws', bilinear
// (align_corners) over (Hs, Ws) to n x n, 'B C Hs Ws -> B (Hs Ws) C'  (aggregation.py:283-285,291-293).
// One CTA per token; threads along the H*q*q channels (coalesced writes; reads of 4 source positions, each a
// contiguous q*q run per head).
__global__ void corr_to_tokens_kernel(const float* __restrict__ corr, float* __restrict__ tok, int H, int hs, int q, int n,
                                      int ld, int col0) {
  const int p = blockIdx.x, b = blockIdx.y, y = p / n, x = p % n, qq = q * q, CH = H * qq;
  const Lerp ly = lerp_ac(y, hs, n), lx = lerp_ac(x, hs, n);
  const float* cb = corr + (size_t)b * H * hs * hs * qq;
  float* out = tok + ((size_t)b * n * n + p) * ld + col0;
  for (int c = threadIdx.x; c < CH; c += blockDim.x) {
    const int h = c / qq, t = c % qq;
    const float* ch = cb + (size_t)h * hs * hs * qq + t;
    const float v00 = ch[(size_t)(ly.i0 * hs + lx.i0) * qq], v01 = ch[(size_t)(ly.i0 * hs + lx.i1) * qq];
    const float v10 = ch[(size_t)(ly.i1 * hs + lx.i0) * qq], v11 = ch[(size_t)(ly.i1 * hs + lx.i1) * qq];
    out[c] = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
  }
}

// ---- tokens (B, n*n, H*q*q) -> corr (B, H, hs, hs, q, q): 'B (Hs Ws) H (Ht Wt) -> B (H Ht Wt) Hs Ws', bilinear to
// hs x hs, back to 'B H Hs Ws Ht Wt' (aggregation.py:298-300). One CTA per (output position, b).
__global__ void tokens_to_corr_kernel(const float* __restrict__ tok, float* __restrict__ corr, int H, int hs, int q, int n) {
  const int p = blockIdx.x, b = blockIdx.y, y = p / hs, x = p % hs, qq = q * q, CH = H * qq;
  const Lerp ly = lerp_ac(y, n, hs), lx = lerp_ac(x, n, hs);
  const float* tb = tok + (size_t)b * n * n * CH;
  const float* r00 = tb + (size_t)(ly.i0 * n + lx.i0) * CH;
  const float* r01 = tb + (size_t)(ly.i0 * n + lx.i1) * CH;
  const float* r10 = tb + (size_t)(ly.i1 * n + lx.i0) * CH;
  const float* r11 = tb + (size_t)(ly.i1 * n + lx.i1) * CH;
  float* cb = corr + (size_t)b * H * hs * hs * qq;
  for (int c = threadIdx.x; c < CH; c += blockDim.x) {
    const int h = c / qq, t = c % qq;
    cb[((size_t)h * hs * hs + p) * qq + t] = ly.l0 * (lx.l0 * r00[c] + lx.l1 * r01[c]) + ly.l1 * (lx.l0 * r10[c] + lx.l1 * r11[c]);
  }
}

// ---- 'B H Hs Ws Ht Wt -> B H Ht Wt Hs Ws': batched (P x Q) -> (Q x P) transpose through shared memory -------------
__global__ void transpose_pq_kernel(const float* __restrict__ in, float* __restrict__ out, int P, int Q) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * P * Q;
  const int p0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, q = q0 + threadIdx.x;
    if (p < P && q < Q) tile[i][threadIdx.x] = in[base + (size_t)p * Q + q];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int q = q0 + i, p = p0 + threadIdx.x;
    if (p < P && q < Q) out[base + (size_t)q * P + p] = tile[threadIdx.x][i];
  }
}

// ---- depthwise 3x3 conv (pad 1) on the n x n token map + bias + exact GELU (aggregation.py:18-29,186-187) ----------
__global__ void dwconv_gelu_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ y, int n, int C) {
  const int p = blockIdx.x, b = blockIdx.y, py = p / n, px = p % n;
  const float* xb = x + (size_t)b * n * n * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = bias[c];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = py + dy - 1;
      if (yy < 0 || yy >= n) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = px + dx - 1;
        if (xx < 0 || xx >= n) continue;
        acc = fmaf(xb[(size_t)(yy * n + xx) * C + c], w[c * 9 + dy * 3 + dx], acc);
      }
    }
    y[((size_t)b * n * n + p) * C + c] = 0.5f * acc * (1.f + erff(acc * 0.70710678118654752440f));
  }
}

// ---- token-map resampling: bilinear align_corners upsample (interpolate2d_token, aggregation.py:58-63), average
// pooling by `pool` (einops reduce 'mean', :316-319) and nearest repeat by `pool` (einops repeat, :327-332) ----------
__global__ void upsample_tokens_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int m, int C) {
  const int p = blockIdx.x, b = blockIdx.y;
  const Lerp ly = lerp_ac(p / m, n, m), lx = lerp_ac(p % m, n, m);
  const float* xb = x + (size_t)b * n * n * C;
  const float *r00 = xb + (size_t)(ly.i0 * n + lx.i0) * C, *r01 = xb + (size_t)(ly.i0 * n + lx.i1) * C;
  const float *r10 = xb + (size_t)(ly.i1 * n + lx.i0) * C, *r11 = xb + (size_t)(ly.i1 * n + lx.i1) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    y[((size_t)b * m * m + p) * C + c] = ly.l0 * (lx.l0 * r00[c] + lx.l1 * r01[c]) + ly.l1 * (lx.l0 * r10[c] + lx.l1 * r11[c]);
}
__global__ void avgpool_tokens_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int pool, int C) {
  const int m = n / pool, p = blockIdx.x, b = blockIdx.y, py = p / m, px = p % m;
  const float* xb = x + (size_t)b * n * n * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < pool; ++i)
      for (int j = 0; j < pool; ++j) s += xb[(size_t)((py * pool + i) * n + px * pool + j) * C + c];
    y[((size_t)b * m * m + p) * C + c] = s / (float)(pool * pool);
  }
}
__global__ void repeat_tokens_kernel(const float* __restrict__ x, float* __restrict__ y, int m, int pool, int C) {
  const int n = m * pool, p = blockIdx.x, b = blockIdx.y, py = p / n, px = p % n;
  const float* src = x + ((size_t)b * m * m + (py / pool) * m + px / pool) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) y[((size_t)b * n * n + p) * C + c] = src[c];
}

// ---- cross attention with softmax(corr) as the attention matrix (aggregation.py:314,324-325) -----------------------
// corr (B, H, S, T); v (B, T or S, H, D). rows: out[b, s, h, :] = sum_t softmax_t(corr[b,h,s,:])[t] * vt[b, t, h, :]
//                                         cols: out[b, t, h, :] = sum_s softmax_s(corr[b,h,:,t])[s] * vs[b, s, h, :]
// One CTA of 256 threads per (position, h, b); the softmax axis has at most 1024 entries.
template <bool ROWS>
__global__ void __launch_bounds__(256) cross_attn_kernel(const float* __restrict__ corr, const float* __restrict__ v,
                                                         float* __restrict__ out, int H, int S, int T, int D) {
  extern __shared__ float p[];   // probabilities along the softmax axis
  __shared__ float red[8];
  const int pos = blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int n = ROWS ? T : S;
  const float* base = corr + ((size_t)b * H + h) * S * T + (ROWS ? (size_t)pos * T : (size_t)pos);
  const size_t stride = ROWS ? 1 : (size_t)T;
  float m = -INFINITY;
  for (int i = tid; i < n; i += 256) m = fmaxf(m, base[i * stride]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = tid; i < n; i += 256) {
    float e = expf(base[i * stride] - m);
    p[i] = e;
    s += e;
  }
  s = warp_sum_f(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  // out[d] = sum_i p[i] / tot * v[b, i, h, d]; 8 groups of threads split the axis, fixed-order combine
  __shared__ float part[8][32];
  const int d = tid & 31, grp = tid >> 5;
  float acc = 0.f;
  if (d < D)
    for (int i = grp; i < n; i += 8) acc = fmaf(p[i] / tot, v[(((size_t)b * n + i) * H + h) * D + d], acc);
  part[grp][d] = acc;
  __syncthreads();
  if (tid < D) {
    float r = 0.f;
    for (int g2 = 0; g2 < 8; ++g2) r += part[g2][tid];
    const int npos = ROWS ? S : T;
    out[(((size_t)b * npos + pos) * H + h) * D + tid] = r;
  }
}

}  // namespace

#define CPN_REQUIRE(cond, name)                      \
  do {                                               \
    if (!(cond)) {                                   \
      cpn_set_error(name ": bad argument");          \
      return CPN_ERR_ARG;                            \
    }                                                \
  } while (0)

extern "C" int cpn_layernorm(const float* x, const float* gamma, const float* beta, float* y, int tokens, int C, void* stream) {
  CPN_REQUIRE(x && gamma && beta && y && tokens > 0 && C > 0, "cpn_layernorm");
  layernorm_kernel<<<(tokens * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, tokens, C);
  CPN_CHECK_LAUNCH("layernorm_kernel");
  return CPN_OK;
}

extern "C" int cpn_corr_to_tokens(const float* corr, float* tok, int B, int H, int hs, int q, int n, int ld, int col0,
                                  void* stream) {
  CPN_REQUIRE(corr && tok && B > 0 && H > 0 && hs > 0 && q > 0 && n > 0 && ld >= col0 + H * q * q, "cpn_corr_to_tokens");
  corr_to_tokens_kernel<<<dim3(n * n, B), 256, 0, (cudaStream_t)stream>>>(corr, tok, H, hs, q, n, ld, col0);
  CPN_CHECK_LAUNCH("corr_to_tokens_kernel");
  return CPN_OK;
}

extern "C" int cpn_tokens_to_corr(const float* tok, float* corr, int B, int H, int hs, int q, int n, void* stream) {
  CPN_REQUIRE(corr && tok && B > 0 && H > 0 && hs > 0 && q > 0 && n > 0, "cpn_tokens_to_corr");
  tokens_to_corr_kernel<<<dim3(hs * hs, B), 256, 0, (cudaStream_t)stream>>>(tok, corr, H, hs, q, n);
  CPN_CHECK_LAUNCH("tokens_to_corr_kernel");
  return CPN_OK;
}

extern "C" int cpn_transpose_pq(const float* in, float* out, int batch, int P, int Q, void* stream) {
  CPN_REQUIRE(in && out && batch > 0 && batch <= 65535 && P > 0 && Q > 0, "cpn_transpose_pq");
  transpose_pq_kernel<<<dim3((Q + 31) / 32, (P + 31) / 32, batch), dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, P, Q);
  CPN_CHECK_LAUNCH("transpose_pq_kernel");
  return CPN_OK;
}

extern "C" int cpn_dwconv_gelu(const float* x, const float* w, const float* bias, float* y, int B, int n, int C, void* stream) {
  CPN_REQUIRE(x && w && bias && y && B > 0 && n > 0 && C > 0, "cpn_dwconv_gelu");
  dwconv_gelu_kernel<<<dim3(n * n, B), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, n, C);
  CPN_CHECK_LAUNCH("dwconv_gelu_kernel");
  return CPN_OK;
}

// mode 0: bilinear (align_corners) n -> m; 1: average pool n -> n / m_or_pool; 2: nearest repeat n -> n * m_or_pool
extern "C" int cpn_resample_tokens(const float* x, float* y, int B, int n, int m_or_pool, int C, int mode, void* stream) {
  CPN_REQUIRE(x && y && B > 0 && n > 0 && m_or_pool > 0 && C > 0 && mode >= 0 && mode <= 2, "cpn_resample_tokens");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    upsample_tokens_kernel<<<dim3(m_or_pool * m_or_pool, B), 256, 0, st>>>(x, y, n, m_or_pool, C);
  } else if (mode == 1) {
    CPN_REQUIRE(n % m_or_pool == 0, "cpn_resample_tokens");
    int m = n / m_or_pool;
    avgpool_tokens_kernel<<<dim3(m * m, B), 256, 0, st>>>(x, y, n, m_or_pool, C);
  } else {
    int m = n * m_or_pool;
    repeat_tokens_kernel<<<dim3(m * m, B), 256, 0, st>>>(x, y, n, m_or_pool, C);
  }
  CPN_CHECK_LAUNCH("resample_tokens");
  return CPN_OK;
}

extern "C" int cpn_cross_attention(const float* corr, const float* src_v, const float* trg_v, float* src_attn, float* trg_attn,
                                   int B, int H, int S, int T, int D, void* stream) {
  CPN_REQUIRE(corr && src_v && trg_v && src_attn && trg_attn && B > 0 && H > 0 && S > 0 && T > 0 && D > 0 && D <= 32 &&
                  S <= 4096 && T <= 4096, "cpn_cross_attention");
  cudaStream_t st = (cudaStream_t)stream;
  cross_attn_kernel<true><<<dim3(S, H, B), 256, T * sizeof(float), st>>>(corr, trg_v, src_attn, H, S, T, D);
  CPN_CHECK_LAUNCH("cross_attn_kernel<rows>");
  cross_attn_kernel<false><<<dim3(T, H, B), 256, S * sizeof(float), st>>>(corr, src_v, trg_attn, H, S, T, D);
  CPN_CHECK_LAUNCH("cross_attn_kernel<cols>");
  return CPN_OK;
}

// Cosine correlation of token features (aggregation.py:70-80): out[b, s, t] = <src_n[b, s], trg_n[b, t]>,
// x_n = x / (||x|| + 1e-5). From 2048 tokens on (the 64 x 64 level: a 4096 x 4096 x C product per pair and call, the largest
// GEMMs of the cost aggregation) the product runs on the tcgen05 Linear kernel with the normalised target as the "weight"
// (three fp16 MMAs per product, fp32-level: tests/test_ufc_native_gpu.py); below that on the CUDA-core GEMM.
// workspace: normalised src (B, L, C) + normalised (tensor-core path) or normalised, transposed trg + the packed target tiles.
static bool corr_tc(int L, int C) {
  static int simt = -1;   // CPN_CORR_SIMT=1: CUDA-core GEMM at every size (A/B runs)
  if (simt < 0) {
    const char* e = getenv("CPN_CORR_SIMT");
    simt = (e && atoi(e) != 0) ? 1 : 0;
  }
  return !simt && L >= 2048 && (L % 128) == 0 && (C % 8) == 0;
}
extern "C" size_t cpn_correlation_workspace_bytes(int B, int L, int C) {
  if (B <= 0 || L <= 0 || C <= 0) return 0;
  return (size_t)2 * B * L * C * sizeof(float) + 512 + (corr_tc(L, C) ? (cpn_linear_tc_packed_bytes(L, C) + 255) / 256 * 256 : 0);
}
int launch_ufc_normalize(const float* in, float* out, int tokens, int C, cudaStream_t st);   // ufc_tail.cu
extern "C" int cpn_correlation(const float* src, const float* trg, float* out, int B, int L, int C, void* workspace,
                               size_t workspace_bytes, void* stream) {
  CPN_REQUIRE(src && trg && out && workspace && B > 0 && L > 0 && C > 0 && !(C & 3) && !(L & 3) &&
                  workspace_bytes >= cpn_correlation_workspace_bytes(B, L, C), "cpn_correlation");
  cudaStream_t st = (cudaStream_t)stream;
  float* sn = reinterpret_cast<float*>(workspace);
  float* tnT = sn + (size_t)B * L * C;
  if ((size_t)L < (size_t)C) {
    cpn_set_error("cpn_correlation: L < C unsupported");
    return CPN_ERR_ARG;
  }
  int rc = launch_ufc_normalize(src, sn, B * L, C, st);
  if (rc != CPN_OK) return rc;
  if (corr_tc(L, C)) {
    float* tn = tnT;   // (B, L, C), row-major: the (N, K) layout cpn_linear_tc_pack takes
    unsigned char* packed = reinterpret_cast<unsigned char*>(workspace) + ((size_t)2 * B * L * C * sizeof(float) + 511) / 256 * 256;
    rc = launch_ufc_normalize(trg, tn, B * L, C, st);
    if (rc != CPN_OK) return rc;
    for (int b = 0; b < B; ++b) {
      rc = cpn_linear_tc_pack(tn + (size_t)b * L * C, L, C, packed, stream);
      if (rc != CPN_OK) return rc;
      rc = launch_linear_tc(packed, L, C, sn + (size_t)b * L * C, C, nullptr, out + (size_t)b * L * L, L, L, 0, CPN_TC_F16X3, 1.f, st);
      if (rc != CPN_OK) return rc;
    }
    return CPN_OK;
  }
  float* tn = out;   // the (B, L, L) output is large enough to stage the normalised target before the transpose
  rc = launch_ufc_normalize(trg, tn, B * L, C, st);
  if (rc != CPN_OK) return rc;
  transpose_pq_kernel<<<dim3((C + 31) / 32, (L + 31) / 32, B), dim3(32, 8), 0, st>>>(tn, tnT, L, C);
  CPN_CHECK_LAUNCH("transpose_pq_kernel");
  for (int b = 0; b < B; ++b) {
    rc = launch_gemm_simt(sn + (size_t)b * L * C, C, tnT + (size_t)b * C * L, nullptr, nullptr, 1, out + (size_t)b * L * L, L, L,
                          L, C, 0, st);
    if (rc != CPN_OK) return rc;
  }
  return CPN_OK;
}
