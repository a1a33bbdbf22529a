// Per-pair pose features and pose head: the CUDA operators behind CrossBlock / CrossAttention
// (models/backbone.py:262-431) and the pose / rotation / translation regressors + r6d2mat of
// models/CoPoNeRF.py:33-52,106-128,194-204.  SURVEY.md section 8(f) rank 2.
//
//   cpn_dual_softmax    P = softmax(c, -1) * softmax(c, -2) of the (L x L) averaged correlation `c`
//                       (backbone.py:290-291; attn_fundamental_2 is its transpose and is never materialised)
//   cpn_gemm_tn         C = A^T B (+ bias, activation) with the long dimension split over CTAs: the
//                       (C+6) x L x (C+6) products v^T (P v) of backbone.py:311-312 and the Linear applied to a
//                       transposed matrix (proj_fundamental, :323-324) without a transposed copy
//   cpn_linear_skinny   y = act(x W^T + b) for a handful of rows and a very long K (pose_regressor[0]:
//                       134 144 -> 512, 275 MB of weights read once at HBM speed)
//   cpn_pose_head       the remaining ten small Linear layers, Gram-Schmidt (r6d2mat) and the 4 x 4 assembly
//
// All reductions run in a fixed order (two-stage split reductions, no atomics).
#include <math.h>
#include "cpn_common.cuh"

namespace {

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

// ---- softmax statistics ----------------------------------------------------------------------------------------
// rows: one warp per row of the (rows x L) matrix, float4 loads
__global__ void row_stats_kernel(const float* __restrict__ c, int rows, int L, float* __restrict__ rmax,
                                 float* __restrict__ rsum) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(c + (size_t)row * L);
  float m = -INFINITY;
  for (int i = lane; i < L / 4; i += 32) {
    float4 v = p[i];
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  m = warp_max_f(m);
  float s = 0.f;
  for (int i = lane; i < L / 4; i += 32) {
    float4 v = p[i];
    s += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
  }
  s = warp_sum_f(s);
  if (lane == 0) {
    rmax[row] = m;
    rsum[row] = s;
  }
}

// columns: thread per column, a CTA covers COL_ROWS rows; (max, sum) partials merged by col_merge_kernel
constexpr int COL_ROWS = 64;
__global__ void col_partial_kernel(const float* __restrict__ c, int L, int nblk, float* __restrict__ pmax,
                                   float* __restrict__ psum) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, blk = blockIdx.y, b = blockIdx.z;
  if (t >= L) return;
  const float* cb = c + (size_t)b * L * L + t;
  const int r0 = blk * COL_ROWS, r1 = min(L, r0 + COL_ROWS);
  float m = -INFINITY;
  for (int r = r0; r < r1; ++r) m = fmaxf(m, cb[(size_t)r * L]);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += expf(cb[(size_t)r * L] - m);
  pmax[((size_t)b * nblk + blk) * L + t] = m;
  psum[((size_t)b * nblk + blk) * L + t] = s;
}
__global__ void col_merge_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, int L, int nblk,
                                 float* __restrict__ cmax, float* __restrict__ csum) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (t >= L) return;
  float m = -INFINITY;
  for (int k = 0; k < nblk; ++k) m = fmaxf(m, pmax[((size_t)b * nblk + k) * L + t]);
  float s = 0.f;
  for (int k = 0; k < nblk; ++k) s += psum[((size_t)b * nblk + k) * L + t] * expf(pmax[((size_t)b * nblk + k) * L + t] - m);
  cmax[(size_t)b * L + t] = m;
  csum[(size_t)b * L + t] = s;
}

// P[s][t] = exp(c - rmax[s]) / rsum[s] * exp(c - cmax[t]) / csum[t]
__global__ void dual_softmax_kernel(const float* __restrict__ c, const float* __restrict__ rmax,
                                    const float* __restrict__ rsum, const float* __restrict__ cmax,
                                    const float* __restrict__ csum, float* __restrict__ P, int L) {
  const int s = blockIdx.y, b = blockIdx.z;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t >= L) return;
  const size_t off = ((size_t)b * L + s) * L + t;
  const float4 v = *reinterpret_cast<const float4*>(c + off);
  const float rm = rmax[(size_t)b * L + s], rs = rsum[(size_t)b * L + s];
  const float4 cm = *reinterpret_cast<const float4*>(cmax + (size_t)b * L + t);
  const float4 cs = *reinterpret_cast<const float4*>(csum + (size_t)b * L + t);
  float4 o;
  o.x = (expf(v.x - rm) / rs) * (expf(v.x - cm.x) / cs.x);
  o.y = (expf(v.y - rm) / rs) * (expf(v.y - cm.y) / cs.y);
  o.z = (expf(v.z - rm) / rs) * (expf(v.z - cm.z) / cs.z);
  o.w = (expf(v.w - rm) / rs) * (expf(v.w - cm.w) / cs.w);
  *reinterpret_cast<float4*>(P + off) = o;
}

// ---- C[i][j] = sum_l A[l][i] * B[l][j]: 64 x 64 tiles, the l range split over blockIdx.z ---------------------------
constexpr int TM = 64, TN = 64, TK = 16;
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ part,
               int Mi, int Nj, int L, int lsplit) {
  __shared__ __align__(16) float As[TK][TM];
  __shared__ __align__(16) float Bs[TK][TN];
  const int tid = threadIdx.x, i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
  const int l0 = blockIdx.z * lsplit, l1 = min(L, l0 + lsplit);
  const int tx = tid & 15, ty = tid >> 4;
  const int lc = tid & 63, lr = tid >> 6;   // loader: column 0..63, rows lr, lr+4, lr+8, lr+12
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int l = l0; l < l1; l += TK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ll = l + lr + r * 4;
      As[lr + r * 4][lc] = (ll < l1 && i0 + lc < Mi) ? A[(size_t)ll * lda + i0 + lc] : 0.f;
      Bs[lr + r * 4][lc] = (ll < l1 && j0 + lc < Nj) ? B[(size_t)ll * ldb + j0 + lc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.z * Mi * Nj;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = i0 + ty * 4 + i;
    if (m >= Mi) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = j0 + tx * 4 + j;
      if (n < Nj) out[(size_t)m * Nj + n] = acc[i][j];
    }
  }
}

// C[m][n] = act(sum_z part[z][m][n] + bias[n]), z ascending
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int nsplit, int M, int N, const float* __restrict__ bias,
                                     int act, float* __restrict__ C, int ldc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int m = idx / N, n = idx % N;
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * M * N + idx];
  if (bias) s += bias[n];
  C[(size_t)m * ldc + n] = apply_act(s, act);
}

// ---- y[m][n] = act(<x[m], W[n]> + b[n]) for M <= 8 rows: one warp per (output n, k-split) -----------------------------
constexpr int SK_MAXM = 8;
__global__ void __launch_bounds__(256)
linear_skinny_kernel(const float* __restrict__ x, const float* __restrict__ W, float* __restrict__ part, int M, int N,
                     int K, int ksplit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const int k0 = blockIdx.y * ksplit, k1 = min(K, k0 + ksplit);
  const float* wr = W + (size_t)n * K;
  float acc[SK_MAXM];
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) acc[m] = 0.f;
  for (int k = k0 + lane * 4; k < k1; k += 128) {
    const float4 w = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
    for (int m = 0; m < SK_MAXM; ++m) {
      if (m < M) {
        const float4 v = *reinterpret_cast<const float4*>(x + (size_t)m * K + k);
        acc[m] = fmaf(v.x, w.x, acc[m]);
        acc[m] = fmaf(v.y, w.y, acc[m]);
        acc[m] = fmaf(v.z, w.z, acc[m]);
        acc[m] = fmaf(v.w, w.w, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) {
    if (m < M) {
      const float s = warp_sum_f(acc[m]);
      if (lane == 0) part[((size_t)blockIdx.y * M + m) * N + n] = s;
    }
  }
}

// ---- pose head: one CTA per pair ------------------------------------------------------------------------------------
// out[n] = act(<in, W[n]> + b[n]), W row-major (N, K) as in the state_dict; warps stride over outputs
__device__ void dense(const float* in, int K, const float* __restrict__ W, const float* __restrict__ b, float* out, int N,
                      int relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int n = warp; n < N; n += nw) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(in[k], W[(size_t)n * K + k], s);
    s = warp_sum_f(s);
    if (lane == 0) {
      s += b[n];
      out[n] = relu ? fmaxf(s, 0.f) : s;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) pose_head_kernel(cpn_pose_head_args a) {
  __shared__ float h0[512], h1[256], lat[128], r1[64], r2[32], t1[64], t2[32], r6[6], tr[3];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < 512; i += blockDim.x) h0[i] = a.h0[(size_t)b * 512 + i];   // already ReLU'd
  __syncthreads();
  dense(h0, 512, a.w2, a.b2, h1, 256, 1);
  dense(h1, 256, a.w4, a.b4, lat, 128, 1);          // pose_regressor(...)[:, :128]; the leading nn.ReLU of both heads is a no-op
  dense(lat, 128, a.rw1, a.rb1, r1, 64, 1);
  dense(r1, 64, a.rw3, a.rb3, r2, 32, 1);
  dense(r2, 32, a.rw5, a.rb5, r6, 6, 0);
  dense(lat, 128, a.tw1, a.tb1, t1, 64, 1);
  dense(t1, 64, a.tw3, a.tb3, t2, 32, 1);
  dense(t2, 32, a.tw5, a.tb5, tr, 3, 0);
  if (tid == 0) {
    // r6d2mat (CoPoNeRF.py:106-128): F.normalize(x) = x / max(|x|, 1e-12)
    float a1[3] = {r6[0], r6[1], r6[2]}, a2[3] = {r6[3], r6[4], r6[5]}, b1[3], b2[3], b3[3];
    float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    for (int i = 0; i < 3; ++i) b1[i] = a1[i] / n1;
    const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    for (int i = 0; i < 3; ++i) b2[i] = a2[i] - d * b1[i];
    float n2 = fmaxf(sqrtf(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]), 1e-12f);
    for (int i = 0; i < 3; ++i) b2[i] = b2[i] / n2;
    b3[0] = b1[1] * b2[2] - b1[2] * b2[1];
    b3[1] = b1[2] * b2[0] - b1[0] * b2[2];
    b3[2] = b1[0] * b2[1] - b1[1] * b2[0];
    float* o = a.rel_pose + (size_t)b * 16;
    for (int i = 0; i < 3; ++i) {
      o[0 + i] = b1[i];
      o[4 + i] = b2[i];
      o[8 + i] = b3[i];
      o[4 * i + 3] = tr[i];
    }
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  }
}

int tn_splits(int Mi, int Nj, int L, int* lsplit) {
  const int tiles = ((Mi + TM - 1) / TM) * ((Nj + TN - 1) / TN);
  int want = (592 + tiles - 1) / tiles;                 // about 4 CTAs per SM
  const int maxs = (L + 63) / 64;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  int ls = (L + want - 1) / want;
  ls = (ls + TK - 1) / TK * TK;
  *lsplit = ls;
  return (L + ls - 1) / ls;
}

int skinny_splits(int N, int K, int* ksplit) {
  const int blocks = (N + 7) / 8;
  int want = (1184 + blocks - 1) / blocks;              // about 8 CTAs per SM
  const int maxs = (K + 2047) / 2048;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  int ks = (K + want - 1) / want;
  ks = (ks + 127) / 128 * 128;
  *ksplit = ks;
  return (K + ks - 1) / ks;
}

}  // namespace

#define CPN_REQUIRE(cond, name)                      \
  do {                                               \
    if (!(cond)) {                                   \
      cpn_set_error("%s: bad argument", name);       \
      return CPN_ERR_ARG;                            \
    }                                                \
  } while (0)

extern "C" size_t cpn_dual_softmax_workspace_bytes(int B, int L) {
  if (B <= 0 || L <= 0) return 0;
  const size_t nblk = (L + COL_ROWS - 1) / COL_ROWS;
  return ((size_t)4 * B * L + (size_t)2 * B * nblk * L) * sizeof(float);
}

extern "C" int cpn_dual_softmax(const float* c, float* P, int B, int L, void* workspace, size_t workspace_bytes,
                                void* stream) {
  CPN_REQUIRE(c && P && workspace && B > 0 && B <= 65535 && L > 0 && (L & 3) == 0 && L <= 65535, "cpn_dual_softmax");
  if (workspace_bytes < cpn_dual_softmax_workspace_bytes(B, L)) {
    cpn_set_error("cpn_dual_softmax: workspace of %zu bytes needed, %zu given", cpn_dual_softmax_workspace_bytes(B, L),
                  workspace_bytes);
    return CPN_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = (L + COL_ROWS - 1) / COL_ROWS;
  float* rmax = (float*)workspace;
  float* rsum = rmax + (size_t)B * L;
  float* cmax = rsum + (size_t)B * L;
  float* csum = cmax + (size_t)B * L;
  float* pmax = csum + (size_t)B * L;
  float* psum = pmax + (size_t)B * nblk * L;
  const int rows = B * L;
  row_stats_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(c, rows, L, rmax, rsum);
  CPN_CHECK_LAUNCH("row_stats_kernel");
  col_partial_kernel<<<dim3((L + 127) / 128, nblk, B), 128, 0, st>>>(c, L, nblk, pmax, psum);
  CPN_CHECK_LAUNCH("col_partial_kernel");
  col_merge_kernel<<<dim3((L + 127) / 128, B), 128, 0, st>>>(pmax, psum, L, nblk, cmax, csum);
  CPN_CHECK_LAUNCH("col_merge_kernel");
  dual_softmax_kernel<<<dim3((L / 4 + 255) / 256, L, B), 256, 0, st>>>(c, rmax, rsum, cmax, csum, P, L);
  CPN_CHECK_LAUNCH("dual_softmax_kernel");
  return CPN_OK;
}

extern "C" size_t cpn_gemm_tn_workspace_bytes(int Mi, int Nj, int L) {
  if (Mi <= 0 || Nj <= 0 || L <= 0) return 0;
  int ls;
  const int ns = tn_splits(Mi, Nj, L, &ls);
  return (size_t)ns * Mi * Nj * sizeof(float);
}

extern "C" int cpn_gemm_tn(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int Mi,
                           int Nj, int L, int act, void* workspace, size_t workspace_bytes, void* stream) {
  CPN_REQUIRE(A && B && C && workspace && Mi > 0 && Nj > 0 && L > 0 && lda >= Mi && ldb >= Nj && ldc >= Nj && act >= 0 &&
                  act <= 2, "cpn_gemm_tn");
  if (workspace_bytes < cpn_gemm_tn_workspace_bytes(Mi, Nj, L)) {
    cpn_set_error("cpn_gemm_tn: workspace of %zu bytes needed, %zu given", cpn_gemm_tn_workspace_bytes(Mi, Nj, L),
                  workspace_bytes);
    return CPN_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int ls;
  const int ns = tn_splits(Mi, Nj, L, &ls);
  gemm_tn_kernel<<<dim3((Nj + TN - 1) / TN, (Mi + TM - 1) / TM, ns), 256, 0, st>>>(A, lda, B, ldb, (float*)workspace, Mi,
                                                                                  Nj, L, ls);
  CPN_CHECK_LAUNCH("gemm_tn_kernel");
  splitk_reduce_kernel<<<(Mi * Nj + 255) / 256, 256, 0, st>>>((const float*)workspace, ns, Mi, Nj, bias, act, C, ldc);
  CPN_CHECK_LAUNCH("splitk_reduce_kernel");
  return CPN_OK;
}

extern "C" size_t cpn_linear_skinny_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int ks;
  const int ns = skinny_splits(N, K, &ks);
  return (size_t)ns * M * N * sizeof(float);
}

extern "C" int cpn_linear_skinny(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, int act,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  CPN_REQUIRE(x && W && y && workspace && M > 0 && M <= SK_MAXM && N > 0 && K > 0 && (K & 3) == 0 && act >= 0 && act <= 2,
              "cpn_linear_skinny");
  if (workspace_bytes < cpn_linear_skinny_workspace_bytes(M, N, K)) {
    cpn_set_error("cpn_linear_skinny: workspace of %zu bytes needed, %zu given", cpn_linear_skinny_workspace_bytes(M, N, K),
                  workspace_bytes);
    return CPN_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int ks;
  const int ns = skinny_splits(N, K, &ks);
  linear_skinny_kernel<<<dim3((N + 7) / 8, ns), 256, 0, st>>>(x, W, (float*)workspace, M, N, K, ks);
  CPN_CHECK_LAUNCH("linear_skinny_kernel");
  splitk_reduce_kernel<<<(M * N + 255) / 256, 256, 0, st>>>((const float*)workspace, ns, M, N, bias, act, y, N);
  CPN_CHECK_LAUNCH("splitk_reduce_kernel");
  return CPN_OK;
}

extern "C" int cpn_pose_head(const cpn_pose_head_args* args, void* stream) {
  CPN_REQUIRE(args, "cpn_pose_head");
  const cpn_pose_head_args& a = *args;
  CPN_REQUIRE(a.B > 0 && a.h0 && a.w2 && a.b2 && a.w4 && a.b4 && a.rw1 && a.rb1 && a.rw3 && a.rb3 && a.rw5 && a.rb5 &&
                  a.tw1 && a.tb1 && a.tw3 && a.tb3 && a.tw5 && a.tb5 && a.rel_pose, "cpn_pose_head");
  pose_head_kernel<<<a.B, 256, 0, (cudaStream_t)stream>>>(a);
  CPN_CHECK_LAUNCH("pose_head_kernel");
  return CPN_OK;
}
