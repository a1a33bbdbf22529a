// fp32 CUDA-core GEMM for the 1x1-conv layers whose K or N is too small for the tensor-core path
// (query_embed 16->128, the 128->128 second layers, encode_latent) and as the bring-up / cross-check
// path for the large layers:  C[M,N] = act(A[M,K] * Wt[K,N] + bias[N] + rowbias[m / rows_per_bias, N]).
//
// Replaces the nn.Conv2d(kernel 1) calls of models/CoPoNeRF.py:387-408,446,467-473.
// The k-loop order is fixed and independent of the tile a row lands in, so a ray's result does not
// depend on the chunk or rank that renders it.
#include "cpn_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

// activation code `relu`: 0 none, 1 ReLU, 2 exact GELU (nn.GELU default, erf form)
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(NT, 2)
gemm_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wt, const float* __restrict__ bias,
                 const float* __restrict__ rowbias, int rows_per_bias, float* __restrict__ C, int ldc, int M, int N,
                 int K, int relu, int remap256, float out_div) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // A loader: thread -> (row, 8 consecutive k)
  const int a_row = tid & (BM - 1), a_k = (tid >> 7) * 8;
  const bool a_ok = (m0 + a_row) < M;
  const float* a_ptr = A + (size_t)(m0 + a_row) * lda + a_k;
  // B loader: thread -> (k, float4 column), two k rows
  const int b_k = tid >> 5, b_n = (tid & 31) * 4;
  const bool b_ok = (n0 + b_n) < N;
  const float* b_ptr = Wt + (size_t)b_k * N + n0 + b_n;

  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int k = k0 + a_k + j * 4;
      ra[j] = (a_ok && k < K) ? *reinterpret_cast<const float4*>(a_ptr + k0 + j * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      int kb = k0 + b_k + j * 8;
      rb[j] = (b_ok && kb < K) ? *reinterpret_cast<const float4*>(b_ptr + (size_t)(k0 + j * 8) * N)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      As[buf][a_k + j * 4 + 0][a_row] = ra[j].x;
      As[buf][a_k + j * 4 + 1][a_row] = ra[j].y;
      As[buf][a_k + j * 4 + 2][a_row] = ra[j].z;
      As[buf][a_k + j * 4 + 3][a_row] = ra[j].w;
      *reinterpret_cast<float4*>(&Bs[buf][b_k + j * 8][b_n]) = rb[j];
    }
  };

  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each an 8 x 8 micro-tile split in 4 + 4
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = (k0 + BK) < K;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const float* rbp = rowbias ? rowbias + (size_t)(m / rows_per_bias) * N : nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int n = n0 + h * 64 + tx * 4;
      if (n >= N) continue;
      float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      if (bias) {
        float4 bb = *reinterpret_cast<const float4*>(bias + n);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      if (rbp) {
        float4 bb = *reinterpret_cast<const float4*>(rbp + n);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      if (relu == 1) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      } else if (relu == 2) {
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
      }
      if (out_div != 0.f) {
        v.x /= out_div; v.y /= out_div; v.z /= out_div; v.w /= out_div;
      }
      size_t orow = remap256 ? (size_t)(m >> 8) * 128 + (m & 127) : (size_t)m;
      int ocol = remap256 ? ((m >> 7) & 1) * N + n : n;
      *reinterpret_cast<float4*>(C + orow * ldc + ocol) = v;
    }
  }
}

// 64 x 64 tiles, 4 x 4 per thread: for problems whose 128 x 128 tiling would leave most of the 148 SMs idle (the
// per-pair layers of the cost aggregation: M = 256 .. 4096 tokens). Same k order per output as the large kernel, so
// both give bit-identical results.
constexpr int SBM = 64, SBN = 64;

__global__ void __launch_bounds__(NT, 2)
gemm_simt_small_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wt, const float* __restrict__ bias,
                       const float* __restrict__ rowbias, int rows_per_bias, float* __restrict__ C, int ldc, int M, int N,
                       int K, int relu, int remap256, float out_div, float* __restrict__ part, int ksplit) {
  // part != nullptr: split-K. CTA z covers k in [z * ksplit, min(K, (z + 1) * ksplit)) and writes its raw partial
  // sums to part[z][M][N]; splitk_finish_kernel adds them in ascending z and applies bias / activation.
  __shared__ __align__(16) float As[2][BK][SBM];
  __shared__ __align__(16) float Bs[2][BK][SBN];
  const int kbeg = part ? blockIdx.z * ksplit : 0;
  if (part) K = min(K, kbeg + ksplit);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int a_row = tid & (SBM - 1), a_k = (tid >> 6) * 4;          // 64 rows x 4 groups of 4 k
  const bool a_ok = (m0 + a_row) < M;
  const float* a_ptr = A + (size_t)(m0 + a_row) * lda + a_k;
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;                   // 16 k rows x 16 float4 columns
  const bool b_ok = (n0 + b_n) < N;
  const float* b_ptr = Wt + (size_t)b_k * N + n0 + b_n;
  float4 ra, rb;
  auto load_tiles = [&](int k0) {
    ra = (a_ok && k0 + a_k < K) ? *reinterpret_cast<const float4*>(a_ptr + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    rb = (b_ok && k0 + b_k < K) ? *reinterpret_cast<const float4*>(b_ptr + (size_t)k0 * N) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_tiles = [&](int buf) {
    As[buf][a_k + 0][a_row] = ra.x;
    As[buf][a_k + 1][a_row] = ra.y;
    As[buf][a_k + 2][a_row] = ra.z;
    As[buf][a_k + 3][a_row] = ra.w;
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
  };
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  load_tiles(kbeg);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < K; k0 += BK) {
    const bool more = (k0 + BK) < K;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i, n = n0 + tx * 4;
    if (m >= M || n >= N) continue;
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (part) {
      *reinterpret_cast<float4*>(part + ((size_t)blockIdx.z * M + m) * N + n) = v;
      continue;
    }
    if (bias) {
      float4 bb = *reinterpret_cast<const float4*>(bias + n);
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    }
    if (rowbias) {
      float4 bb = *reinterpret_cast<const float4*>(rowbias + (size_t)(m / rows_per_bias) * N + n);
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    }
    if (relu == 1) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    } else if (relu == 2) {
      v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
    }
    if (out_div != 0.f) {
      v.x /= out_div; v.y /= out_div; v.z /= out_div; v.w /= out_div;
    }
    size_t orow = remap256 ? (size_t)(m >> 8) * 128 + (m & 127) : (size_t)m;
    int ocol = remap256 ? ((m >> 7) & 1) * N + n : n;
    *reinterpret_cast<float4*>(C + orow * ldc + ocol) = v;
  }
}

// ---- 16 -> 128 ReLU layer written straight as the operand image of the 128 -> 128 layer that follows -------------
// query_embed (CoPoNeRF.py:446) and the local_coords half of query_repeat_embed (:467-473, the z_embed half enters as
// a per-ray bias): out[row][n] = relu(sum_j x[row][j] * Wt[j][n] + bias[n] + rowbias[row / rows_per_bias][n]).
// With fp32 output the consumer GEMM has to convert its A operand itself (producer warps, DRAM-latency bound); the
// image form lets it bulk-copy 16 KB blocks. CTA = one 128-row tile, thread -> (row, half of the 128 outputs).
struct Mlp16Extra {
  const float* sdot1;   // [128] vector + [128] constant: s1[row] = <out, vec> + const   (nullable)
  float* s1;
  const float* sdot2;
  float* s2;
  const float* dotv;    // CB16 fp32 (M, 128): lg[row] = (<out, dotv[row]> + rowadd[row]) / div   (nullable)
  const float* rowadd;
  float div;
  float* lg;
  int dot_blocks, dot_block0;   // dotv rows: [row tile][dot_blocks][128][16], this layer's 8 blocks start at dot_block0
};

template <bool F8>
__global__ void __launch_bounds__(128) mlp16_image_kernel(const float* __restrict__ x, const float* __restrict__ Wt,
                                                          const float* __restrict__ bias, const float* __restrict__ rowbias,
                                                          int rows_per_bias, int M, unsigned char* __restrict__ img,
                                                          Mlp16Extra e) {
  // CTA = two 128-row tiles, thread -> the same row of each tile with all 128 outputs: every shared-memory weight load
  // feeds two rows, and the optional per-row dot products stay inside one thread
  __shared__ __align__(16) float ws[16][CPN_HIDDEN];
  __shared__ __align__(16) float bs[CPN_HIDDEN];
  __shared__ __align__(16) float sd[2][CPN_HIDDEN];
  const int t = threadIdx.x, rloc = t;
  for (int i = t; i < 16 * CPN_HIDDEN; i += 128) ws[i / CPN_HIDDEN][i % CPN_HIDDEN] = Wt[i];
  bs[t] = bias ? bias[t] : 0.f;
  sd[0][t] = e.sdot1 ? e.sdot1[t] : 0.f;
  sd[1][t] = e.sdot2 ? e.sdot2[t] : 0.f;
  __syncthreads();
  const size_t tile0 = (size_t)blockIdx.x * 2;
  float in[2][16];
  const float* rb[2];
  bool ok[2];
  float d1[2] = {0.f, 0.f}, d2[2] = {0.f, 0.f}, dl[2] = {0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const size_t row = (tile0 + q) * 128 + rloc;
    ok[q] = row < (size_t)M;
    const size_t rr = ok[q] ? row : 0;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + rr * 16 + j));
      in[q][j] = v.x; in[q][j + 1] = v.y; in[q][j + 2] = v.z; in[q][j + 3] = v.w;
    }
    rb[q] = rowbias ? rowbias + (rr / rows_per_bias) * CPN_HIDDEN : nullptr;
  }
#pragma unroll 1
  for (int n0 = 0; n0 < CPN_HIDDEN; n0 += 8) {
    float acc[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = bs[n0 + c];
      if (rb[q]) {
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(rb[q] + n0)), r1 = __ldg(reinterpret_cast<const float4*>(rb[q] + n0 + 4));
        acc[q][0] += r0.x; acc[q][1] += r0.y; acc[q][2] += r0.z; acc[q][3] += r0.w;
        acc[q][4] += r1.x; acc[q][5] += r1.y; acc[q][6] += r1.z; acc[q][7] += r1.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[j][n0]), w1 = *reinterpret_cast<const float4*>(&ws[j][n0 + 4]);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[q][c] = fmaf(in[q][j], wv[c], acc[q][c]);
    }
    const int g = (n0 % ACT_BK) / 8;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!ok[q]) continue;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = fmaxf(acc[q][c], 0.f);
      if (e.s1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          d1[q] = fmaf(acc[q][c], sd[0][n0 + c], d1[q]);
          d2[q] = fmaf(acc[q][c], sd[1][n0 + c], d2[q]);
        }
      }
      if (e.dotv) {   // CB16: [row tile][16-column block][128 rows][16]
        const float* dv = e.dotv + (((tile0 + q) * e.dot_blocks + e.dot_block0 + n0 / 16) * 128 + rloc) * 16 + (n0 & 8);
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(dv)), a1 = __ldg(reinterpret_cast<const float4*>(dv + 4));
        dl[q] = fmaf(acc[q][3], a0.w, fmaf(acc[q][2], a0.z, fmaf(acc[q][1], a0.y, fmaf(acc[q][0], a0.x, dl[q]))));
        dl[q] = fmaf(acc[q][7], a1.w, fmaf(acc[q][6], a1.z, fmaf(acc[q][5], a1.y, fmaf(acc[q][4], a1.x, dl[q]))));
      }
      if (!img) continue;
      unsigned char* chunk = img + ((tile0 + q) * (size_t)(CPN_HIDDEN / ACT_BK) + n0 / ACT_BK) * ACT_CHUNK_BYTES + rloc * 16;
      const float4 v0 = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
      const float4 v1 = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
      if (F8) {
        uint2 h0, h1, l8, x8;
        tc::split4_f8(v0, h0, l8.x, x8.x);
        tc::split4_f8(v1, h1, l8.y, x8.y);
        *reinterpret_cast<uint4*>(chunk + g * 2048) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint2*>(chunk + ACT_LO8 + (g >> 1) * 2048 + (g & 1) * 8) = l8;
        *reinterpret_cast<uint2*>(chunk + ACT_X8 + (g >> 1) * 2048 + (g & 1) * 8) = x8;
      } else {
        uint4 hi, lo;
        tc::split8(v0, v1, hi, lo);
        *reinterpret_cast<uint4*>(chunk + g * 2048) = hi;
        *reinterpret_cast<uint4*>(chunk + ACT_LO + g * 2048) = lo;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (!ok[q]) continue;
    const size_t row = (tile0 + q) * 128 + rloc;
    if (e.s1) {
      e.s1[row] = d1[q] + e.sdot1[CPN_HIDDEN];
      if (e.s2) e.s2[row] = d2[q] + e.sdot2[CPN_HIDDEN];
    }
    if (e.dotv) e.lg[row] = (dl[q] + (e.rowadd ? e.rowadd[row] : 0.f)) / e.div;
  }
}

// second stage of the split-K GEMM: C[m][n] = act(sum_z part[z][m][n] + bias[n]), z ascending
__global__ void splitk_finish_kernel(const float* __restrict__ part, int nsplit, int M, int N, const float* __restrict__ bias,
                                     int act, float* __restrict__ C, int ldc) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (idx >= M * N) return;
  const int m = idx / N, n = idx % N;
  float4 s = *reinterpret_cast<const float4*>(part + idx);
  for (int z = 1; z < nsplit; ++z) {
    const float4 p = *reinterpret_cast<const float4*>(part + (size_t)z * M * N + idx);
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  if (bias) {
    const float4 b = *reinterpret_cast<const float4*>(bias + n);
    s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
  }
  if (act == 1) {
    s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f);
  } else if (act == 2) {
    s.x = gelu_erf(s.x); s.y = gelu_erf(s.y); s.z = gelu_erf(s.z); s.w = gelu_erf(s.w);
  }
  *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = s;
}

// split count for the 64 x 64 tiling: enough CTAs for about two per SM, at least 128 k per split
int simt_splits(int M, int N, int K, int* ksplit) {
  const long long tiles = (long long)((N + SBN - 1) / SBN) * ((M + SBM - 1) / SBM);
  int want = (int)((296 + tiles - 1) / tiles);
  const int maxs = K / 128;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  int ks = (K + want - 1) / want;
  ks = (ks + BK - 1) / BK * BK;
  *ksplit = ks;
  return (K + ks - 1) / ks;
}

}  // namespace

int launch_mlp16_image(const float* x, const float* wt, const float* bias, const float* rowbias, int rows_per_bias, int M,
                       void* img, int f8, cudaStream_t st, const float* sdot1, float* s1, const float* sdot2, float* s2,
                       const float* dotv, const float* rowadd, float div, float* lg, int dot_blocks, int dot_block0) {
  if (M <= 0) return CPN_OK;
  if ((!img && !dotv) || (dotv && !lg) || (s1 && !sdot1) || (s2 && (!sdot2 || !s1))) {
    cpn_set_error("mlp16_image: inconsistent outputs");
    return CPN_ERR_ARG;
  }
  const unsigned ctas = (unsigned)((M + 255) / 256);   // two 128-row tiles per CTA
  const Mlp16Extra e{sdot1, s1, sdot2, s2, dotv, rowadd, div, lg, dot_blocks, dot_block0};
  if (f8)
    mlp16_image_kernel<true><<<ctas, 128, 0, st>>>(x, wt, bias, rowbias, rows_per_bias > 0 ? rows_per_bias : 1, M,
                                                   reinterpret_cast<unsigned char*>(img), e);
  else
    mlp16_image_kernel<false><<<ctas, 128, 0, st>>>(x, wt, bias, rowbias, rows_per_bias > 0 ? rows_per_bias : 1, M,
                                                    reinterpret_cast<unsigned char*>(img), e);
  CPN_CHECK_LAUNCH("mlp16_image_kernel");
  return CPN_OK;
}

int launch_gemm_simt(const float* A, int lda, const float* wt, const float* bias, const float* rowbias,
                     int rows_per_bias, float* C, int ldc, int M, int N, int K, int relu, cudaStream_t st, int remap256,
                     float out_div) {
  if (M <= 0) return CPN_OK;
  if ((K & 3) || (N & 3) || (lda & 3) || (ldc & 3)) {
    cpn_set_error("gemm_simt: K, N, lda and ldc must be multiples of 4 (K=%d N=%d lda=%d ldc=%d)", K, N, lda, ldc);
    return CPN_ERR_ARG;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  if ((long long)grid.x * grid.y < 96) {   // too few large tiles to fill the 148 SMs
    dim3 g2((N + SBN - 1) / SBN, (M + SBM - 1) / SBM);
    gemm_simt_small_kernel<<<g2, NT, 0, st>>>(A, lda, wt, bias, rowbias, rows_per_bias > 0 ? rows_per_bias : 1, C, ldc, M,
                                              N, K, relu, remap256, out_div, nullptr, 0);
    CPN_CHECK_LAUNCH("gemm_simt_small_kernel");
    return CPN_OK;
  }
  if (grid.y > 65535) {
    cpn_set_error("gemm_simt: M=%d too large for one launch", M);
    return CPN_ERR_ARG;
  }
  gemm_simt_kernel<<<grid, NT, 0, st>>>(A, lda, wt, bias, rowbias, rows_per_bias > 0 ? rows_per_bias : 1, C, ldc, M, N,
                                        K, relu, remap256, out_div);
  CPN_CHECK_LAUNCH("gemm_simt_kernel");
  return CPN_OK;
}

extern "C" int cpn_gemm_simt(const float* A, int lda, const float* wt, const float* bias, float* C, int ldc, int M,
                             int N, int K, int relu, void* stream) {
  if (!A || !wt || !C) {
    cpn_set_error("cpn_gemm_simt: null pointer");
    return CPN_ERR_ARG;
  }
  return launch_gemm_simt(A, lda, wt, bias, nullptr, 1, C, ldc, M, N, K, relu, (cudaStream_t)stream, 0, 0.f);
}

// Split-K variant for GEMMs whose 64 x 64 tiling leaves most SMs idle (the token layers of the cost aggregation at
// 256 / 1024 tokens with K up to 2304): the k range is split over blockIdx.z and reduced in a fixed order.
extern "C" size_t cpn_gemm_simt_splitk_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int ks;
  const int ns = simt_splits(M, N, K, &ks);
  return ns > 1 ? (size_t)ns * M * N * sizeof(float) : 0;
}

extern "C" int cpn_gemm_simt_splitk(const float* A, int lda, const float* wt, const float* bias, float* C, int ldc, int M,
                                    int N, int K, int act, void* workspace, size_t workspace_bytes, void* stream) {
  if (!A || !wt || !C) {
    cpn_set_error("cpn_gemm_simt_splitk: null pointer");
    return CPN_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int ks;
  const int ns = M > 0 && N > 0 && K > 0 ? simt_splits(M, N, K, &ks) : 1;
  if (ns <= 1) return launch_gemm_simt(A, lda, wt, bias, nullptr, 1, C, ldc, M, N, K, act, st, 0, 0.f);
  if ((K & 3) || (N & 3) || (lda & 3) || (ldc & 3)) {
    cpn_set_error("gemm_simt_splitk: K, N, lda and ldc must be multiples of 4 (K=%d N=%d lda=%d ldc=%d)", K, N, lda, ldc);
    return CPN_ERR_ARG;
  }
  const size_t need = (size_t)ns * M * N * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    cpn_set_error("cpn_gemm_simt_splitk: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    return CPN_ERR_WORKSPACE;
  }
  dim3 grid((N + SBN - 1) / SBN, (M + SBM - 1) / SBM, ns);
  gemm_simt_small_kernel<<<grid, NT, 0, st>>>(A, lda, wt, nullptr, nullptr, 1, nullptr, 0, M, N, K, 0, 0, 0.f,
                                              (float*)workspace, ks);
  CPN_CHECK_LAUNCH("gemm_simt_small_kernel");
  splitk_finish_kernel<<<(M * N / 4 + 255) / 256, 256, 0, st>>>((const float*)workspace, ns, M, N, bias, act, C, ldc);
  CPN_CHECK_LAUNCH("splitk_finish_kernel");
  return CPN_OK;
}
