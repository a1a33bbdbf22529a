// Linear attention of the cost aggregation (models/aggregation.py:84-117, "Transformers are RNNs"):
//   Q = elu(q) + 1, K = elu(k) + 1, V' = V / S
//   KV[h] = K[:, h]^T V'[:, h]  (D x Dv),   Z[l, h] = 1 / (Q[l, h] . sum_s K[s, h] + 1e-6)
//   out[l, h] = (Q[l, h] KV[h]) * Z[l, h] * S
// q, k (N, L, H, D), v (N, S, H, Dv); D = 32, Dv in {32, 256}, 20 calls per stereo pair.
// Two kernels: a reduction over the S keys into KV / Ksum (one CTA per (n, h, 32-wide slice of Dv), fixed summation
// order), then one thread per output element group.
#include <math.h>
#include "cpn_common.cuh"

namespace {

constexpr int LA_D = 32;

__device__ __forceinline__ float elu1(float x) { return (x > 0.f ? x : expm1f(x)) + 1.f; }

// KV[n][h][d][v] for a 32-wide v slice; the v-slice-0 CTA also writes Ksum[n][h][d].
// block (32, 8): threadIdx.x = v within the slice, threadIdx.y = 4-row group of d; loop over keys in tiles of 32.
__global__ void __launch_bounds__(256) la_kv_kernel(const float* __restrict__ k, const float* __restrict__ v, int S, int H,
                                                    int Dv, float* __restrict__ KV, float* __restrict__ Ksum) {
  __shared__ float ks[32][LA_D + 1];   // [key in tile][d]
  __shared__ float vs[32][33];         // [key in tile][v in slice]
  const int n = blockIdx.z, h = blockIdx.y, v0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ksum = 0.f;   // threads with ty == 0 accumulate Ksum[d = tx]
  const float invS = (float)S;
  for (int s0 = 0; s0 < S; s0 += 32) {
    for (int i = tid; i < 32 * LA_D; i += 256) {
      int sk = i / LA_D, d = i % LA_D;
      ks[sk][d] = (s0 + sk < S) ? elu1(k[(((size_t)n * S + s0 + sk) * H + h) * LA_D + d]) : 0.f;
    }
    for (int i = tid; i < 32 * 32; i += 256) {
      int sk = i / 32, vv = i % 32;
      vs[sk][vv] = (s0 + sk < S && v0 + vv < Dv) ? v[(((size_t)n * S + s0 + sk) * H + h) * Dv + v0 + vv] / invS : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int sk = 0; sk < 32; ++sk) {
      float vv = vs[sk][tx];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(ks[sk][ty * 4 + j], vv, acc[j]);
      if (ty == 0) ksum += ks[sk][tx];
    }
    __syncthreads();
  }
  if (v0 + tx < Dv) {
#pragma unroll
    for (int j = 0; j < 4; ++j) KV[(((size_t)n * H + h) * LA_D + ty * 4 + j) * Dv + v0 + tx] = acc[j];
  }
  if (blockIdx.x == 0 && ty == 0) Ksum[((size_t)n * H + h) * LA_D + tx] = ksum;
}

// out[n][l][h][v]: one warp per (n, l, h); lanes stride over v
__global__ void __launch_bounds__(256) la_out_kernel(const float* __restrict__ q, const float* __restrict__ KV,
                                                     const float* __restrict__ Ksum, int L, int S, int H, int Dv, int N,
                                                     float* __restrict__ out) {
  const long long w = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)N * L * H) return;
  const int h = (int)(w % H);
  const long long nl = w / H;
  const int n = (int)(nl / L);
  const float qd = elu1(q[(size_t)w * LA_D + lane]);      // lane = d
  float dot = qd * Ksum[((size_t)n * H + h) * LA_D + lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float z = 1.f / (dot + 1e-6f);
  const float* kv = KV + ((size_t)n * H + h) * LA_D * Dv;
  for (int v0 = 0; v0 < Dv; v0 += 32) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < LA_D; ++d) {
      float qv = __shfl_sync(0xffffffffu, qd, d);
      if (v0 + lane < Dv) acc = fmaf(qv, kv[(size_t)d * Dv + v0 + lane], acc);
    }
    if (v0 + lane < Dv) out[(size_t)w * Dv + v0 + lane] = acc * z * (float)S;
  }
}

}  // namespace

extern "C" size_t cpn_linear_attention_workspace_bytes(int N, int H, int Dv) {
  if (N <= 0 || H <= 0 || Dv <= 0) return 0;
  return ((size_t)N * H * LA_D * Dv + (size_t)N * H * LA_D) * sizeof(float);
}

extern "C" int cpn_linear_attention(const float* q, const float* k, const float* v, int N, int L, int S, int H, int D,
                                    int Dv, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !out || !workspace || N <= 0 || L <= 0 || S <= 0 || H <= 0 || Dv <= 0 || D != LA_D) {
    cpn_set_error("cpn_linear_attention: bad argument (head dim must be %d)", LA_D);
    return CPN_ERR_ARG;
  }
  if (cpn_linear_attention_workspace_bytes(N, H, Dv) > workspace_bytes) {
    cpn_set_error("cpn_linear_attention: workspace too small");
    return CPN_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* KV = reinterpret_cast<float*>(workspace);
  float* Ksum = KV + (size_t)N * H * LA_D * Dv;
  dim3 g1((Dv + 31) / 32, H, N), b1(32, 8);
  la_kv_kernel<<<g1, b1, 0, st>>>(k, v, S, H, Dv, KV, Ksum);
  CPN_CHECK_LAUNCH("la_kv_kernel");
  long long warps = (long long)N * L * H;
  la_out_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(q, KV, Ksum, L, S, H, Dv, N, out);
  CPN_CHECK_LAUNCH("la_out_kernel");
  return CPN_OK;
}
