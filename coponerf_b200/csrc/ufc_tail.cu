// Closing stage of UFC.forward() (models/aggregation.py:527,539,549-561): from the refined source / target token
// features of the three pyramid levels to the averaged 4-D correlation volume `c` and the four flow fields.
//
// The reference builds three cosine-correlation volumes (16^4, 32^4, 64^4), upsamples each to 64^4 with two
// separable bilinear passes over 67 MB tensors (interpolate4d), averages them and runs two softmaxes over 4096
// positions. Bilinear interpolation is linear and a correlation entry is a dot product of one source and one
// target vector, so   interp4d(corr_l)[s, t] = < up(src_l)[s], up(trg_l)[t] >   with `up` the 2-D bilinear
// upsampling of the L2-normalised feature maps. Hence
//     c = (1/3) * [up(S_0) | up(S_1) | up(S_2)] * [up(T_0) | up(T_1) | up(T_2)]^T ,
// one (4096 x 768) x (768 x 4096) GEMM per pair: `c` is written once and nothing else of size 64^4 exists. It runs on the
// tcgen05 kernel with three fp16 MMAs per product (4e-6 of fp64 at K = 768; the soft-argmax at temperature 0.02 amplifies
// errors of c by 50, flows stay within 1e-4 px of the fp32 path): the target operand is packed per pair like a weight (cpn_linear_tc_pack), the source operand streams as fp32 rows.
// The soft-argmax passes then read `c` row-wise (flow_t_to_s / flow) and column-wise (flow_s_to_t / flow_flip).
#include <math.h>
#include "cpn_common.cuh"

namespace {

// ---- x / (||x||_2 + 1e-5) per token (aggregation.py:71-72). One warp per token.
__global__ void ufc_normalize_kernel(const float* __restrict__ in, float* __restrict__ out, int tokens, int C) {
  int tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  const float* x = in + (size_t)tok * C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) ss += x[c] * x[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  float inv = sqrtf(ss) + 1e-5f;
  for (int c = lane; c < C; c += 32) out[(size_t)tok * C + c] = x[c] / inv;
}

// ---- bilinear upsampling (align_corners=True, F.interpolate semantics) of one normalised level to out x out and
// packing into the GEMM operands: S[b][p][koff + c] (row-major, ld = ldk) or Tt[b][koff + c][p] (k-major).
// One CTA per destination pixel, threads along channels.
__global__ void ufc_upsample_pack_kernel(const float* __restrict__ tok, int n, int out, int C, int koff, int ldk,
                                         float* __restrict__ S, float* __restrict__ Tt) {
  const int p = blockIdx.x, b = blockIdx.y;
  const int y = p / out, x = p % out;
  const float scale = (out > 1) ? (float)(n - 1) / (float)(out - 1) : 0.f;
  float sy = scale * (float)y, sx = scale * (float)x;
  int y0 = (int)sy, x0 = (int)sx;
  int y1 = y0 + (y0 < n - 1 ? 1 : 0), x1 = x0 + (x0 < n - 1 ? 1 : 0);
  float ly1 = sy - (float)y0, lx1 = sx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* base = tok + (size_t)b * n * n * C;
  const float* p00 = base + (size_t)(y0 * n + x0) * C;
  const float* p01 = base + (size_t)(y0 * n + x1) * C;
  const float* p10 = base + (size_t)(y1 * n + x0) * C;
  const float* p11 = base + (size_t)(y1 * n + x1) * C;
  const size_t P = (size_t)out * out;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = ly0 * (lx0 * p00[c] + lx1 * p01[c]) + ly1 * (lx0 * p10[c] + lx1 * p11[c]);
    if (S) S[((size_t)b * P + p) * ldk + koff + c] = v;
    if (Tt) Tt[((size_t)b * ldk + koff + c) * P + p] = v;
  }
}

// ---- soft-argmax over the target positions of every source position (rows of c): softmax((x - max) / 0.02),
// expectation of the normalised x / y grid (aggregation.py:119-144), and the pixel flow (:30-48).
// One CTA per row.
__global__ void __launch_bounds__(256) ufc_softargmax_rows_kernel(const float* __restrict__ c, const float* __restrict__ lin,
                                                                  int out, float* __restrict__ fnorm,
                                                                  float* __restrict__ fpix) {
  __shared__ float red[3][8];
  const int P = out * out, s = blockIdx.x, b = blockIdx.y;
  const float* row = c + ((size_t)b * P + s) * P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY;
  for (int t = tid; t < P; t += 256) m = fmaxf(m, row[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[0][warp] = m;
  __syncthreads();
  m = red[0][0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[0][i]);
  __syncthreads();
  float se = 0.f, sx = 0.f, sy = 0.f;
  for (int t = tid; t < P; t += 256) {
    float e = expf((row[t] - m) / 0.02f);
    se += e;
    sx += e * lin[t % out];
    sy += e * lin[t / out];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if (lane == 0) {
    red[0][warp] = se;
    red[1][warp] = sx;
    red[2][warp] = sy;
  }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, bx = 0.f, by = 0.f;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      bx += red[1][i];
      by += red[2][i];
    }
    float gx = bx / a, gy = by / a;
    const int hs = s / out, ws = s % out;
    const size_t o0 = ((size_t)b * 2 + 0) * P + s, o1 = ((size_t)b * 2 + 1) * P + s;
    fnorm[o0] = gx;
    fnorm[o1] = gy;
    fpix[o0] = (gx + 1.f) * (float)(out - 1) / 2.0f - (float)ws;
    fpix[o1] = (gy + 1.f) * (float)(out - 1) / 2.0f - (float)hs;
  }
}

// ---- the same over the source positions of every target position (columns of c). CTA = 32 columns x 32 row
// slices: coalesced 128-byte reads of c, two passes (max, then sums) over the L2-resident volume.
__global__ void __launch_bounds__(1024) ufc_softargmax_cols_kernel(const float* __restrict__ c, const float* __restrict__ lin,
                                                                   int out, float* __restrict__ fnorm,
                                                                   float* __restrict__ fpix) {
  __shared__ float red[3][32][33];
  const int P = out * out, b = blockIdx.y;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + cx;
  const float* col = c + (size_t)b * P * P + t;
  float m = -INFINITY;
  if (t < P)
    for (int s = ry; s < P; s += 32) m = fmaxf(m, col[(size_t)s * P]);
  red[0][ry][cx] = m;
  __syncthreads();
  m = red[0][0][cx];
  for (int i = 1; i < 32; ++i) m = fmaxf(m, red[0][i][cx]);
  __syncthreads();
  float se = 0.f, sx = 0.f, sy = 0.f;
  if (t < P)
    for (int s = ry; s < P; s += 32) {
      float e = expf((col[(size_t)s * P] - m) / 0.02f);
      se += e;
      sx += e * lin[s % out];
      sy += e * lin[s / out];
    }
  red[0][ry][cx] = se;
  red[1][ry][cx] = sx;
  red[2][ry][cx] = sy;
  __syncthreads();
  if (ry == 0 && t < P) {
    float a = 0.f, bx = 0.f, by = 0.f;
    for (int i = 0; i < 32; ++i) {
      a += red[0][i][cx];
      bx += red[1][i][cx];
      by += red[2][i][cx];
    }
    float gx = bx / a, gy = by / a;
    const int ht = t / out, wt = t % out;
    const size_t o0 = ((size_t)b * 2 + 0) * P + t, o1 = ((size_t)b * 2 + 1) * P + t;
    fnorm[o0] = gx;
    fnorm[o1] = gy;
    fpix[o0] = (gx + 1.f) * (float)(out - 1) / 2.0f - (float)wt;
    fpix[o1] = (gy + 1.f) * (float)(out - 1) / 2.0f - (float)ht;
  }
}

}  // namespace

int launch_ufc_normalize(const float* in, float* out, int tokens, int C, cudaStream_t st) {
  ufc_normalize_kernel<<<(tokens * 32 + 255) / 256, 256, 0, st>>>(in, out, tokens, C);
  CPN_CHECK_LAUNCH("ufc_normalize_kernel");
  return CPN_OK;
}

namespace {

struct TailWs {
  float *ntok[2][3], *S, *Tt, *packed;
  size_t bytes;
};

TailWs carve_tail(void* base, int B, int C, int out, const int* sizes) {
  TailWs w;
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += (floats * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  for (int side = 0; side < 2; ++side)
    for (int l = 0; l < 3; ++l) w.ntok[side][l] = take((size_t)B * sizes[l] * sizes[l] * C);
  w.S = take((size_t)B * out * out * 3 * C);
  w.Tt = take((size_t)B * out * out * 3 * C);     // target operand, row-major (P, K) like S
  w.packed = take((cpn_linear_tc_packed_bytes(out * out, 3 * C) + 3) / 4);   // its tensor-core tiles (one pair at a time; 0 if the shape does not qualify)
  w.bytes = off;
  return w;
}

}  // namespace

extern "C" size_t cpn_ufc_tail_workspace_bytes(int B, int C, int out, const int* sizes) {
  if (B <= 0 || C <= 0 || out <= 0 || !sizes) return 0;
  return carve_tail(nullptr, B, C, out, sizes).bytes;
}

extern "C" int cpn_ufc_tail(const cpn_ufc_tail_args* args, void* stream) {
  if (!args) {
    cpn_set_error("cpn_ufc_tail: null args");
    return CPN_ERR_ARG;
  }
  const cpn_ufc_tail_args& a = *args;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.B <= 0 || a.C <= 0 || (a.C & 3) || a.out <= 1 || ((a.out * a.out) & 3) || !a.lin || !a.c || !a.flow || !a.flow_flip ||
      !a.flow_t_to_s || !a.flow_s_to_t || !a.workspace) {
    cpn_set_error("cpn_ufc_tail: bad argument");
    return CPN_ERR_ARG;
  }
  for (int l = 0; l < 3; ++l)
    if (!a.src[l] || !a.trg[l] || a.sizes[l] <= 0 || a.sizes[l] > a.out) {
      cpn_set_error("cpn_ufc_tail: bad level %d", l);
      return CPN_ERR_ARG;
    }
  TailWs w = carve_tail(a.workspace, a.B, a.C, a.out, a.sizes);
  if (w.bytes > a.workspace_bytes) {
    cpn_set_error("cpn_ufc_tail: workspace of %zu bytes needed, %zu given", w.bytes, a.workspace_bytes);
    return CPN_ERR_WORKSPACE;
  }
  const int P = a.out * a.out, K = 3 * a.C;
  const bool use_tc = (P % 128) == 0 && (K % 8) == 0;   // the tensor-core kernel's shape rules; else the fp32 CUDA-core GEMM
  for (int l = 0; l < 3; ++l) {
    int tokens = a.B * a.sizes[l] * a.sizes[l];
    for (int side = 0; side < 2; ++side) {
      ufc_normalize_kernel<<<(tokens * 32 + 255) / 256, 256, 0, st>>>(side ? a.trg[l] : a.src[l], w.ntok[side][l], tokens, a.C);
      CPN_CHECK_LAUNCH("ufc_normalize_kernel");
    }
    dim3 grid(P, a.B);
    ufc_upsample_pack_kernel<<<grid, 256, 0, st>>>(w.ntok[0][l], a.sizes[l], a.out, a.C, l * a.C, K, w.S, nullptr);
    CPN_CHECK_LAUNCH("ufc_upsample_pack_kernel");
    if (use_tc)   // target operand row-major (P, K), packed like a weight below
      ufc_upsample_pack_kernel<<<grid, 256, 0, st>>>(w.ntok[1][l], a.sizes[l], a.out, a.C, l * a.C, K, w.Tt, nullptr);
    else          // k-major for the fp32 CUDA-core GEMM
      ufc_upsample_pack_kernel<<<grid, 256, 0, st>>>(w.ntok[1][l], a.sizes[l], a.out, a.C, l * a.C, K, nullptr, w.Tt);
    CPN_CHECK_LAUNCH("ufc_upsample_pack_kernel");
  }
  for (int b = 0; b < a.B; ++b) {   // c[b] = S[b] * T[b]^T / 3
    int rc;
    if (use_tc) {
      rc = cpn_linear_tc_pack(w.Tt + (size_t)b * P * K, P, K, w.packed, stream);
      if (rc != CPN_OK) return rc;
      rc = launch_linear_tc(w.packed, P, K, w.S + (size_t)b * P * K, K, nullptr, a.c + (size_t)b * P * P, P, P, 0, CPN_TC_F16X3,
                            1.f / 3.f, st);
    } else {
      rc = launch_gemm_simt(w.S + (size_t)b * P * K, K, w.Tt + (size_t)b * K * P, nullptr, nullptr, 1, a.c + (size_t)b * P * P,
                            P, P, P, K, 0, st, 0, 3.0f);
    }
    if (rc != CPN_OK) return rc;
  }
  dim3 rows(P, a.B), cols((P + 31) / 32, a.B);
  ufc_softargmax_rows_kernel<<<rows, 256, 0, st>>>(a.c, a.lin, a.out, a.flow_t_to_s, a.flow);
  CPN_CHECK_LAUNCH("ufc_softargmax_rows_kernel");
  ufc_softargmax_cols_kernel<<<cols, 1024, 0, st>>>(a.c, a.lin, a.out, a.flow_s_to_t, a.flow_flip);
  CPN_CHECK_LAUNCH("ufc_softargmax_cols_kernel");
  return CPN_OK;
}
