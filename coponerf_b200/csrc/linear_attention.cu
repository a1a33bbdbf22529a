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
// The keys are split into `nsplit` contiguous ranges (blockIdx.z = n * nsplit + split); each CTA writes a partial
// KV / Ksum, la_reduce_kernel adds the partials in split order.
__global__ void __launch_bounds__(256) la_kv_kernel(const float* __restrict__ k, const float* __restrict__ v, int S, int H,
                                                    int Dv, int nsplit, float* __restrict__ KV, float* __restrict__ Ksum) {
  __shared__ float ks[32][LA_D + 1];   // [key in tile][d]
  __shared__ float vs[32][33];         // [key in tile][v in slice]
  const int n = blockIdx.z / nsplit, split = blockIdx.z % nsplit, h = blockIdx.y, v0 = blockIdx.x * 32;
  const int per = ((S + nsplit - 1) / nsplit + 31) / 32 * 32, s_begin = split * per, s_end = min(S, s_begin + per);
  const int N = gridDim.z / nsplit;
  KV += (size_t)split * N * H * LA_D * Dv;
  Ksum += (size_t)split * N * H * LA_D;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ksum = 0.f;   // threads with ty == 0 accumulate Ksum[d = tx]
  const float invS = (float)S;
  for (int s0 = s_begin; s0 < s_end; s0 += 32) {
    for (int i = tid; i < 32 * LA_D; i += 256) {
      int sk = i / LA_D, d = i % LA_D;
      ks[sk][d] = (s0 + sk < s_end) ? elu1(k[(((size_t)n * S + s0 + sk) * H + h) * LA_D + d]) : 0.f;
    }
    for (int i = tid; i < 32 * 32; i += 256) {
      int sk = i / 32, vv = i % 32;
      vs[sk][vv] = (s0 + sk < s_end && v0 + vv < Dv) ? v[(((size_t)n * S + s0 + sk) * H + h) * Dv + v0 + vv] / invS : 0.f;
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

__global__ void la_reduce_kernel(float* __restrict__ buf, size_t count, int nsplit) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = buf[i];
  for (int j = 1; j < nsplit; ++j) s += buf[i + (size_t)j * count];
  buf[i] = s;
}

// out[n][l][h][:] = (Q[n,l,h,:] KV[n,h]) * Z * S. CTA = (n, h, 64 consecutive l, one 256-wide slice of Dv): that slice of
// KV[n,h] (32 x 256) and Ksum sit in shared memory; each of the 8 warps handles 8 positions, lane = d for the feature
// map / normaliser and lane = v (8 chunks of 32, accumulated together) for the output.
constexpr int LA_LT = 64, LA_MAXC = 8, LA_VS = 32 * LA_MAXC;
__global__ void __launch_bounds__(256) la_out_kernel(const float* __restrict__ q, const float* __restrict__ KV,
                                                     const float* __restrict__ Ksum, int L, int S, int H, int Dv,
                                                     int vpasses, float* __restrict__ out) {
  __shared__ float kvs[LA_D * LA_VS];
  __shared__ float ksum_s[LA_D];
  const int n = blockIdx.z / vpasses, vbase = (blockIdx.z % vpasses) * LA_VS, h = blockIdx.y, l0 = blockIdx.x * LA_LT;
  const int vw = min(LA_VS, Dv - vbase);                       // width of this slice
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* kv = KV + ((size_t)n * H + h) * LA_D * Dv;
  for (int i = threadIdx.x; i < LA_D * vw; i += 256) kvs[(i / vw) * LA_VS + i % vw] = kv[(size_t)(i / vw) * Dv + vbase + i % vw];
  if (threadIdx.x < LA_D) ksum_s[threadIdx.x] = Ksum[((size_t)n * H + h) * LA_D + threadIdx.x];
  __syncthreads();
  const int nchunk = (vw + 31) / 32;
  for (int li = warp; li < LA_LT; li += 8) {
    const int l = l0 + li;
    if (l >= L) break;
    const size_t row = ((size_t)n * L + l) * H + h;
    const float qd = elu1(q[row * LA_D + lane]);
    float dot = qd * ksum_s[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float z = 1.f / (dot + 1e-6f);
    float acc[LA_MAXC];
#pragma unroll
    for (int c = 0; c < LA_MAXC; ++c) acc[c] = 0.f;
#pragma unroll 4
    for (int d = 0; d < LA_D; ++d) {
      const float qv = __shfl_sync(0xffffffffu, qd, d);
#pragma unroll
      for (int c = 0; c < LA_MAXC; ++c)
        if (c < nchunk && c * 32 + lane < vw) acc[c] = fmaf(qv, kvs[d * LA_VS + c * 32 + lane], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < LA_MAXC; ++c)
      if (c < nchunk && c * 32 + lane < vw) out[row * Dv + vbase + c * 32 + lane] = acc[c] * z * (float)S;
  }
}

}  // namespace

constexpr int LA_MAX_SPLIT = 32;
static int la_splits(int S) { int n = (S + 127) / 128; return n < 1 ? 1 : (n > LA_MAX_SPLIT ? LA_MAX_SPLIT : n); }

extern "C" size_t cpn_linear_attention_workspace_bytes(int N, int H, int Dv) {
  if (N <= 0 || H <= 0 || Dv <= 0) return 0;
  return ((size_t)N * H * LA_D * Dv + (size_t)N * H * LA_D) * sizeof(float) * LA_MAX_SPLIT;
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
  const int nsplit = la_splits(S);
  const size_t kv_count = (size_t)N * H * LA_D * Dv, ks_count = (size_t)N * H * LA_D;
  float* KV = reinterpret_cast<float*>(workspace);
  float* Ksum = KV + kv_count * LA_MAX_SPLIT;
  dim3 g1((Dv + 31) / 32, H, N * nsplit), b1(32, 8);
  la_kv_kernel<<<g1, b1, 0, st>>>(k, v, S, H, Dv, nsplit, KV, Ksum);
  CPN_CHECK_LAUNCH("la_kv_kernel");
  if (nsplit > 1) {
    la_reduce_kernel<<<(unsigned)((kv_count + 255) / 256), 256, 0, st>>>(KV, kv_count, nsplit);
    la_reduce_kernel<<<(unsigned)((ks_count + 255) / 256), 256, 0, st>>>(Ksum, ks_count, nsplit);
    CPN_CHECK_LAUNCH("la_reduce_kernel");
  }
  const int vpasses = (Dv + LA_VS - 1) / LA_VS;
  dim3 g2((L + LA_LT - 1) / LA_LT, H, N * vpasses);
  la_out_kernel<<<g2, 256, 0, st>>>(q, KV, Ksum, L, S, H, Dv, vpasses, out);
  CPN_CHECK_LAUNCH("la_out_kernel");
  return CPN_OK;
}
