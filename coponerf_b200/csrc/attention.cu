// Two rounds of the joint (view x sample) softmax attention and the light-field decoder phi.
//   round 1: models/CoPoNeRF.py:450-463   logits <K, Q> / 11.31, softmax over 2S, readout of V
//   round 2: models/CoPoNeRF.py:467-485   logits <Q2, Q> / 11.31, softmax, readout + residual
//   phi:     models/lightfield.py:131-167 (ResnetFC, 3 blocks), called at models/CoPoNeRF.py:545-559
// One CTA of 2S threads per ray; thread t owns sample-row t of the ray (rows of a ray are contiguous:
// view 0 samples, then view 1 samples). Reductions run in a fixed order per ray, so the result of a
// ray does not depend on which chunk or rank renders it.
#include <math.h>
#include "cpn_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// logits of the rows this warp owns: lane l ends up with the logit of row (warp * 32 + l)
__device__ __forceinline__ float warp_row_logits(const float* __restrict__ x, const float* __restrict__ y, size_t row0,
                                                 int warp, int lane) {
  float mine = 0.f;
  for (int r = 0; r < 32; ++r) {
    size_t row = row0 + warp * 32 + r;
    float4 a = __ldg(reinterpret_cast<const float4*>(x + row * CPN_HIDDEN) + lane);
    float4 b = __ldg(reinterpret_cast<const float4*>(y + row * CPN_HIDDEN) + lane);
    float d = warp_sum(a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w);
    if (lane == r) mine = d;
  }
  return mine / 11.31f;
}

// softmax over the 2S logits of the ray (one per thread); returns this thread's weight
__device__ __forceinline__ float block_softmax(float logit, float* red, int warp, int lane, int nwarps) {
  float m = warp_max(logit);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  float mx = red[0];
  for (int i = 1; i < nwarps; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float e = expf(logit - mx);
  float s = warp_sum(e);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < nwarps; ++i) tot += red[i];
  __syncthreads();
  return e / tot;
}

// Round 1. key, qemb (R,128); value (R,416); rowaux (R,8). Outputs: at_wt / at_wt_max in the reference's
// (2B, N, S) layout, r1 (rays,416), wp (rays,4) = sum_v sum_s w * clamp(pt, +-100).
__global__ void attn1_kernel(cpn_render_args a, int ray0, int nr, const float* __restrict__ key,
                             const float* __restrict__ qemb, const float* __restrict__ value,
                             const float* __restrict__ rowaux, float* __restrict__ r1, float* __restrict__ wp) {
  extern __shared__ float sm[];
  const int S = a.S, S2 = 2 * S;
  float* w = sm;            // [2S]
  float* red = sm + S2;     // [8]
  const int ray = blockIdx.x;  // b * nr + nl
  const int b = ray / nr, nl = ray % nr, n = ray0 + nl;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = S2 >> 5;
  const size_t row0 = (size_t)ray * S2;
  float logit = warp_row_logits(key, qemb, row0, warp, lane);
  float wt = block_softmax(logit, red, warp, lane, nwarps);
  w[t] = wt;
  {
    int v = t / S, s = t % S;
    a.at_wt[(((size_t)(b * 2 + v)) * a.N + n) * S + s] = wt;
  }
  __syncthreads();
  if (t < 2) {  // torch.argmax: first index of the maximum
    int best = 0;
    float bw = w[t * S];
    for (int s = 1; s < S; ++s) {
      float x = w[t * S + s];
      if (x > bw) { bw = x; best = s; }
    }
    a.at_wt_max[((size_t)(b * 2 + t)) * a.N + n] = best;
  }
  for (int c = t; c < CPN_LATENT; c += S2) {
    const float* vp = value + row0 * CPN_LATENT + c;
    float acc0 = 0.f, acc1 = 0.f;
    for (int s = 0; s < S; ++s) acc0 += vp[(size_t)s * CPN_LATENT] * w[s];
    for (int s = 0; s < S; ++s) acc1 += vp[(size_t)(S + s) * CPN_LATENT] * w[S + s];
    r1[(size_t)ray * CPN_LATENT + c] = acc0 + acc1;
  }
  if (t >= 32 && t < 35) {
    int c = t - 32;
    const float* pp = rowaux + row0 * CPN_ROWAUX + 4 + c;
    float acc0 = 0.f, acc1 = 0.f;
    for (int s = 0; s < S; ++s) acc0 += w[s] * pp[(size_t)s * CPN_ROWAUX];
    for (int s = 0; s < S; ++s) acc1 += w[S + s] * pp[(size_t)(S + s) * CPN_ROWAUX];
    wp[(size_t)ray * 4 + c] = acc0 + acc1;
  }
}

// Round 2. z = sum_v (sum_s w2 * V + R1) = R2 + 2 R1.
__global__ void attn2_kernel(cpn_render_args a, int nr, const float* __restrict__ q2, const float* __restrict__ qemb,
                             const float* __restrict__ value, const float* __restrict__ r1, float* __restrict__ z) {
  extern __shared__ float sm[];
  const int S = a.S, S2 = 2 * S;
  float* w = sm;
  float* red = sm + S2;
  const int ray = blockIdx.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = S2 >> 5;
  const size_t row0 = (size_t)ray * S2;
  float logit = warp_row_logits(q2, qemb, row0, warp, lane);
  w[t] = block_softmax(logit, red, warp, lane, nwarps);
  __syncthreads();
  for (int c = t; c < CPN_LATENT; c += S2) {
    const float* vp = value + row0 * CPN_LATENT + c;
    float acc0 = 0.f, acc1 = 0.f;
    for (int s = 0; s < S; ++s) acc0 += vp[(size_t)s * CPN_LATENT] * w[s];
    for (int s = 0; s < S; ++s) acc1 += vp[(size_t)(S + s) * CPN_LATENT] * w[S + s];
    float r = r1[(size_t)ray * CPN_LATENT + c];
    z[(size_t)ray * CPN_LATENT + c] = (acc0 + r) + (acc1 + r);
  }
}

// phi. One CTA of 128 threads renders PHI_RAYS rays; thread j owns hidden channel j.
// Input of lin_z is cat(z, z) (both views carry the same latent, models/CoPoNeRF.py:545-552), coords18 is
// [plucker6 + origin3] of view 0 then view 1.
constexpr int PHI_RAYS = 4;   // small tiles: the grid (rays / 4 CTAs) hides the serial k-loops' latency

__device__ __forceinline__ void phi_dense128(const float* __restrict__ wT, const float* __restrict__ bias,
                                             float (*in)[CPN_HIDDEN], float* out, int j, bool relu_in) {
  // out[r] = bias[j] + sum_k act(in[r][k]) * wT[k][j]
  float bj = bias[j];
#pragma unroll
  for (int r = 0; r < PHI_RAYS; ++r) out[r] = bj;
#pragma unroll 8
  for (int k = 0; k < CPN_HIDDEN; ++k) {
    float wv = __ldg(wT + k * CPN_HIDDEN + j);
#pragma unroll
    for (int r = 0; r < PHI_RAYS; ++r) {
      float x = in[r][k];
      if (relu_in) x = fmaxf(x, 0.f);
      out[r] = fmaf(x, wv, out[r]);
    }
  }
}

__global__ void __launch_bounds__(128) phi_kernel(cpn_render_args a, int ray0, int nr, const float* __restrict__ z,
                                                  const float* __restrict__ seg, float* __restrict__ rgb_raw) {
  __shared__ float zs[PHI_RAYS][CPN_LATENT];
  __shared__ float xs[PHI_RAYS][CPN_HIDDEN];
  __shared__ float hs[PHI_RAYS][CPN_HIDDEN];
  __shared__ float cs[PHI_RAYS][20];
  const float* W = reinterpret_cast<const float*>(a.weights);
  const int j = threadIdx.x;
  const int total = a.B * nr;
  const int base = blockIdx.x * PHI_RAYS;
  for (int i = j; i < PHI_RAYS * CPN_LATENT; i += 128) {
    int r = i / CPN_LATENT, c = i % CPN_LATENT;
    zs[r][c] = (base + r < total) ? z[(size_t)(base + r) * CPN_LATENT + c] : 0.f;
  }
  for (int i = j; i < PHI_RAYS * 18; i += 128) {
    int r = i / 18, c = i % 18, v = c / 9, cc = c % 9;
    float val = 0.f;
    if (base + r < total) {
      int b = (base + r) / nr, n = ray0 + (base + r) % nr;
      val = a.coords[(((size_t)(b * 2 + v)) * a.N + n) * 9 + cc];
    }
    cs[r][c] = val;
  }
  __syncthreads();
  float x[PHI_RAYS], tmp[PHI_RAYS];
  {  // lin_in
    float bj = W[pw::PHI_BIN + j];
#pragma unroll
    for (int r = 0; r < PHI_RAYS; ++r) x[r] = bj;
    for (int k = 0; k < 18; ++k) {
      float wv = W[pw::PHI_INT + k * CPN_HIDDEN + j];
#pragma unroll
      for (int r = 0; r < PHI_RAYS; ++r) x[r] = fmaf(cs[r][k], wv, x[r]);
    }
  }
  for (int blk = 0; blk < 3; ++blk) {
    // x += lin_z[blk](cat(z, z))
    const float* wz = W + pw::PHI_ZT + (size_t)blk * 2 * CPN_LATENT * CPN_HIDDEN;
    float bj = W[pw::PHI_BZ + blk * CPN_HIDDEN + j];
#pragma unroll
    for (int r = 0; r < PHI_RAYS; ++r) tmp[r] = bj;
#pragma unroll 8
    for (int k = 0; k < 2 * CPN_LATENT; ++k) {
      float wv = __ldg(wz + (size_t)k * CPN_HIDDEN + j);
      int kk = k < CPN_LATENT ? k : k - CPN_LATENT;
#pragma unroll
      for (int r = 0; r < PHI_RAYS; ++r) tmp[r] = fmaf(zs[r][kk], wv, tmp[r]);
    }
#pragma unroll
    for (int r = 0; r < PHI_RAYS; ++r) {
      x[r] += tmp[r];
      xs[r][j] = x[r];
    }
    __syncthreads();
    // net = fc_0(relu(x)); dx = fc_1(relu(net)); x = x + dx   (lightfield.py:52-62)
    phi_dense128(W + pw::PHI_F0T + (size_t)blk * CPN_HIDDEN * CPN_HIDDEN, W + pw::PHI_B0 + blk * CPN_HIDDEN, xs, tmp, j,
                 true);
#pragma unroll
    for (int r = 0; r < PHI_RAYS; ++r) hs[r][j] = tmp[r];
    __syncthreads();
    phi_dense128(W + pw::PHI_F1T + (size_t)blk * CPN_HIDDEN * CPN_HIDDEN, W + pw::PHI_B1 + blk * CPN_HIDDEN, hs, tmp, j,
                 true);
#pragma unroll
    for (int r = 0; r < PHI_RAYS; ++r) x[r] += tmp[r];
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < PHI_RAYS; ++r) xs[r][j] = fmaxf(x[r], 0.f);
  __syncthreads();
  // lin_out (3 x 128) + white background for rays with no valid epipolar segment (CoPoNeRF.py:562-566)
  if (j < PHI_RAYS * 3) {
    int r = j / 3, c = j % 3;
    if (base + r < total) {
      float acc = 0.f;
      for (int k = 0; k < CPN_HIDDEN; ++k) acc = fmaf(xs[r][k], W[pw::PHI_OUT + c * CPN_HIDDEN + k], acc);
      acc += W[pw::PHI_BOUT + c];
      int b = (base + r) / nr, nl = (base + r) % nr, n = ray0 + nl;
      const float* sg = seg + ((size_t)(base + r) * 2) * 6;
      float valid = (sg[4] != 0.f || sg[6 + 4] != 0.f) ? 1.f : 0.f;
      a.rgb[((size_t)b * a.N + n) * 3 + c] = acc * valid + 1.f * (1.f - valid);
      if (c == 0) a.valid_mask[(size_t)b * a.N + n] = valid;
      (void)rgb_raw;
    }
  }
}

}  // namespace

int launch_attn1(const cpn_render_args& a, int ray0, int nr, const float* key, const float* qemb, const float* value,
                 const float* rowaux, float* r1, float* wp, cudaStream_t st) {
  int S2 = 2 * a.S;
  attn1_kernel<<<a.B * nr, S2, (S2 + 8) * sizeof(float), st>>>(a, ray0, nr, key, qemb, value, rowaux, r1, wp);
  CPN_CHECK_LAUNCH("attn1_kernel");
  return CPN_OK;
}

int launch_attn2(const cpn_render_args& a, int nr, const float* q2, const float* qemb, const float* value,
                 const float* r1, float* z, cudaStream_t st) {
  int S2 = 2 * a.S;
  attn2_kernel<<<a.B * nr, S2, (S2 + 8) * sizeof(float), st>>>(a, nr, q2, qemb, value, r1, z);
  CPN_CHECK_LAUNCH("attn2_kernel");
  return CPN_OK;
}

int launch_phi(const cpn_render_args& a, int ray0, int nr, const float* z, const float* seg, cudaStream_t st) {
  int total = a.B * nr;
  phi_kernel<<<(total + PHI_RAYS - 1) / PHI_RAYS, 128, 0, st>>>(a, ray0, nr, z, seg, nullptr);
  CPN_CHECK_LAUNCH("phi_kernel");
  return CPN_OK;
}
