// Two rounds of the joint (view x sample) softmax attention and the light-field decoder phi.
//   round 1: models/CoPoNeRF.py:450-463   logits <K, Q> / 11.31, softmax over 2S, readout of V
//   round 2: models/CoPoNeRF.py:467-485   logits <Q2, Q> / 11.31, softmax, readout + residual
//   phi:     models/lightfield.py:131-167 (ResnetFC, 3 blocks), called at models/CoPoNeRF.py:545-559
// One CTA of 2S threads per ray; thread t owns sample-row t of the ray (rows of a ray are contiguous:
// view 0 samples, then view 1 samples). Reductions run in a fixed order per ray, so the result of a
// ray does not depend on which chunk or rank renders it.
#include <math.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "cpn_common.cuh"
#include "tc_common.cuh"

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
  for (int r0 = 0; r0 < 32; r0 += 4) {   // four rows in flight: 8 independent 512-byte loads, interleaved reductions
    float d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      size_t row = row0 + warp * 32 + r0 + i;
      float4 a = __ldg(reinterpret_cast<const float4*>(x + row * CPN_HIDDEN) + lane);
      float4 b = __ldg(reinterpret_cast<const float4*>(y + row * CPN_HIDDEN) + lane);
      d[i] = ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] += __shfl_xor_sync(0xffffffffu, d[i], o);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (lane == r0 + i) mine = d[i];
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
                             const float* __restrict__ rowaux, float* __restrict__ r1, float* __restrict__ wp,
                             const float* __restrict__ logits, float* __restrict__ wts_out, const float* __restrict__ gh,
                             float* __restrict__ rbias) {
  extern __shared__ float sm[];
  const int S = a.S, S2 = 2 * S;
  float* w = sm;            // [2S]
  float* red = sm + S2;     // [8] softmax scratch, [8] argmax indices, [8 x 3] weighted-point partials
  const int ray = blockIdx.x;  // b * nr + nl
  const int b = ray / nr, nl = ray % nr, n = ray0 + nl;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = S2 >> 5;
  const size_t row0 = (size_t)ray * S2;
  float logit = logits ? logits[row0 + t] : warp_row_logits(key, qemb, row0, warp, lane);
  float wt = block_softmax(logit, red, warp, lane, nwarps);
  w[t] = wt;
  if (wts_out) wts_out[row0 + t] = wt;   // late readout (readout_image_kernel): the weights leave, V is never formed
  {
    int v = t / S, s = t % S;
    a.at_wt[(((size_t)(b * 2 + v)) * a.N + n) * S + s] = wt;
  }
  __syncthreads();
  {  // torch.argmax over the S samples of each view: first index of the maximum (ties -> lowest index)
    float bv = wt;
    int bi = t % S;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      red[warp] = bv;
      reinterpret_cast<int*>(red)[8 + warp] = bi;   // red has 8 floats; the int slots live in the 8 after it
    }
  }
  // weighted 3-D point: every thread contributes its own row, reduced per view in the fixed order s = 0..S-1
  // inside a warp tree and then across the warps of the view
  float p3[3];
  {
    const float* pp = rowaux + (row0 + t) * CPN_ROWAUX + 4;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = wt * pp[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      p3[c] = x;
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) red[16 + warp * 3 + c] = p3[c];
    }
  }
  __syncthreads();
  const int wpv = nwarps / 2;   // warps per view
  if (t < 2) {
    float bv = red[t * wpv];
    int bi = reinterpret_cast<int*>(red)[8 + t * wpv];
    for (int i = 1; i < wpv; ++i) {
      float ov = red[t * wpv + i];
      int oi = reinterpret_cast<int*>(red)[8 + t * wpv + i];
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    a.at_wt_max[((size_t)(b * 2 + t)) * a.N + n] = bi;
  }
  if (t >= 32 && t < 35) {
    int c = t - 32;
    float acc0 = 0.f, acc1 = 0.f;
    for (int i = 0; i < wpv; ++i) acc0 += red[16 + i * 3 + c];
    for (int i = 0; i < wpv; ++i) acc1 += red[16 + (wpv + i) * 3 + c];
    wp[(size_t)ray * 4 + c] = acc0 + acc1;
  }
  if (gh) {
    // per-ray bias of query_repeat_embed (CoPoNeRF.py:463-472) = sum_rows w1 (G h + g0), rows in ascending order:
    // thread c owns one of the 128 channels, a warp reads 128 contiguous bytes per row
    // gh is column-blocked: [row tile of 128][8 blocks of 16 channels][128 rows][16] (the layer-10 epilogue, gemm_tc.cu)
    for (int c = t; c < CPN_HIDDEN; c += S2) {
      float acc = 0.f;
#pragma unroll 8
      for (int r = 0; r < S2; ++r) {
        const size_t row = row0 + r;
        acc = fmaf(w[r], __ldg(gh + (((row >> 7) * (CPN_HIDDEN / 16) + (c >> 4)) * 128 + (row & 127)) * 16 + (c & 15)), acc);
      }
      rbias[(size_t)ray * CPN_HIDDEN + c] = acc;
    }
  }
  if (!value) return;
  for (int c = t; c < CPN_LATENT; c += S2) {
    const float* vp = value + row0 * CPN_LATENT + c;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
    for (int s = 0; s < S; ++s) acc0 += vp[(size_t)s * CPN_LATENT] * w[s];
#pragma unroll 8
    for (int s = 0; s < S; ++s) acc1 += vp[(size_t)(S + s) * CPN_LATENT] * w[S + s];
    r1[(size_t)ray * CPN_LATENT + c] = acc0 + acc1;
  }
}

// Round 2. z = sum_v (sum_s w2 * V + R1) = R2 + 2 R1.
__global__ void attn2_kernel(cpn_render_args a, int ray0, int nr, const float* __restrict__ q2,
                             const float* __restrict__ qemb, const float* __restrict__ value,
                             const float* __restrict__ r1, float* __restrict__ z_all, const float* __restrict__ logits,
                             float* __restrict__ wts_out, const float* __restrict__ w1) {
  extern __shared__ float sm[];
  const int S = a.S, S2 = 2 * S;
  float* w = sm;
  float* red = sm + S2;
  const int ray = blockIdx.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = S2 >> 5;
  const size_t row0 = (size_t)ray * S2;
  float logit = logits ? logits[row0 + t] : warp_row_logits(q2, qemb, row0, warp, lane);
  w[t] = block_softmax(logit, red, warp, lane, nwarps);
  // w1 given: z = R2 + 2 R1 = WVF (sum_rows (w2 + 2 w1) h) + 3 b, so one readout with the combined weight serves both rounds
  if (wts_out) wts_out[row0 + t] = w1 ? fmaf(2.f, w1[row0 + t], w[t]) : w[t];
  if (!value) return;
  __syncthreads();
  for (int c = t; c < CPN_LATENT; c += S2) {
    const float* vp = value + row0 * CPN_LATENT + c;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
    for (int s = 0; s < S; ++s) acc0 += vp[(size_t)s * CPN_LATENT] * w[s];
#pragma unroll 8
    for (int s = 0; s < S; ++s) acc1 += vp[(size_t)(S + s) * CPN_LATENT] * w[S + s];
    float r = r1[(size_t)ray * CPN_LATENT + c];
    const int b = ray / nr, n = ray0 + ray % nr;
    z_all[((size_t)b * a.N + n) * CPN_LATENT + c] = (acc0 + r) + (acc1 + r);
  }
}

// ---- late readout -----------------------------------------------------------------------------------------------
// latent_value (folded with query_encode_latent_2: V = WVF [h_p ; h_s] + b) is linear and the softmax weights of a ray
// sum to one, so  sum_rows w V = WVF (sum_rows w [h_p ; h_s]) + b : the attention reads out the 1664-wide hidden
// layer straight from its operand image (fp16 head + correction plane, the same bytes the GEMM would consume) and the
// 416 x 1664 layer runs once per RAY instead of once per SAMPLE (128x fewer rows). V is never formed.
//   hbar[ray][br * 832 + k] = sum_{r < 2S} w[ray * 2S + r] * h_br[row r of the ray][k],  rows in ascending order,
// written as the operand image of the per-ray GEMM (K = 1664, rows = rays).
// CTA -> (ray, branch), 4 warps; warp -> a (k-chunk, group of 8 k) unit at a time, lane -> row: a warp load is 32
// consecutive 16-byte rows = 512 contiguous bytes (lanes along k-groups instead made every load 32 wavefronts and
// the L1 data pipe, not HBM, the limit: 83 % LSU-wavefront utilisation at 3.9 TB/s). Each lane accumulates its rows
// r = lane, lane + 32, ..., then the 8 sums are reduced over the lanes with a halving butterfly (9 shuffles).
template <bool F8>
__global__ void __launch_bounds__(128) readout_image_kernel(const unsigned char* __restrict__ img,
                                                            const float* __restrict__ wts, int S2,
                                                            float* __restrict__ hbar, int nr, int out_N, int out_ray0,
                                                            int chunk_bytes) {
  extern __shared__ float sm[];   // [2S] weights of this ray, [832] weighted sums of this branch
  float* hb = sm + S2;
  const int ray = blockIdx.x, br = blockIdx.y, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const size_t row0 = (size_t)ray * S2;
  for (int i = t; i < S2; i += blockDim.x) sm[i] = wts[row0 + i];
  __syncthreads();
  constexpr int KC = CPN_FEAT_DIM / ACT_BK;   // 26 k-chunks per hidden tile
  for (int u = warp; u < KC * 4; u += 4) {
    const int kc = u >> 2, g = u & 3;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 4
    for (int r = lane; r < S2; r += 32) {
      const size_t row = row0 + r;
      const unsigned char* p = img + (((row >> 7) * 2 + br) * KC + kc) * (size_t)chunk_bytes + (row & 127) * 16;
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p + g * 2048));
      const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w};
      float x[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
        x[2 * j] = f.x;
        x[2 * j + 1] = f.y;
      }
      if (F8) {
        const uint2 lo = __ldg(reinterpret_cast<const uint2*>(p + ACT_LO8 + (g >> 1) * 2048 + (g & 1) * 8));
        const uint32_t lw[2] = {lo.x, lo.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = tc::lo8_to_float2(lw[j >> 1] >> ((j & 1) * 16));   // e5m2((x - hi) * 2^10), tc_common.cuh
          x[2 * j] += f.x;
          x[2 * j + 1] += f.y;
        }
      } else {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p + ACT_LO + g * 2048));
        const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
          x[2 * j] += f.x;
          x[2 * j + 1] += f.y;
        }
      }
      const float w = sm[r];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, x[j], acc[j]);
    }
    // halving butterfly: after the xor-16 / 8 / 4 steps every lane holds one of the 8 values (index from lane bits
    // 4, 3, 2), summed over 8 lanes; xor-2 and xor-1 finish it. Fixed order, independent of chunking.
    float a4[4], a2[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float send = (lane & 16) ? acc[j] : acc[j + 4], keep = (lane & 16) ? acc[j + 4] : acc[j];
      a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float send = (lane & 8) ? a4[j] : a4[j + 2], keep = (lane & 8) ? a4[j + 2] : a4[j];
      a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = (lane & 4) ? a2[0] : a2[1], keep = (lane & 4) ? a2[1] : a2[0];
    float a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    if ((lane & 3) == 0) hb[u * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = a1;
  }
  __syncthreads();
  if (t >= KC * 4) return;
  // hbar leaves as the operand image of the per-ray GEMM (rows = rays, 52 k-chunks per 128-ray tile): thread t's
  // 8 values are exactly one 16-byte fp16 group plus half a 16-byte group in each correction plane
  const int kc = t >> 2, g = t & 3;
  const float4 v0 = *reinterpret_cast<const float4*>(hb + t * 8), v1 = *reinterpret_cast<const float4*>(hb + t * 8 + 4);
  // output row: the chunk-local ray (out_N == nr, out_ray0 == 0) or the ray's index in the whole image b * N + n
  const size_t oray = (size_t)(ray / nr) * out_N + out_ray0 + ray % nr;
  unsigned char* oblk = reinterpret_cast<unsigned char*>(hbar) +
                        (((oray >> 7) * (2 * KC) + br * KC + kc) * (size_t)ACT_CHUNK_BYTES) + (oray & 127) * 16;
  if (F8) {
    uint2 h0, h1, l8, x8;
    tc::split4_f8(v0, h0, l8.x, x8.x);
    tc::split4_f8(v1, h1, l8.y, x8.y);
    *reinterpret_cast<uint4*>(oblk + g * 2048) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    *reinterpret_cast<uint2*>(oblk + ACT_LO8 + (g >> 1) * 2048 + (g & 1) * 8) = l8;
    *reinterpret_cast<uint2*>(oblk + ACT_X8 + (g >> 1) * 2048 + (g & 1) * 8) = x8;
  } else {
    uint4 hi4, lo4;
    tc::split8(v0, v1, hi4, lo4);
    *reinterpret_cast<uint4*>(oblk + g * 2048) = hi4;
    *reinterpret_cast<uint4*>(oblk + ACT_LO + g * 2048) = lo4;
  }
}

// z = sum_v (R2 + R1) = R2 + 2 R1 (CoPoNeRF.py:481-485), R2 = WVF hbar2 + b from the late readout of round 2.
// Per chunk R1 is parked in z_all at the ray's image index; once every chunk is done, one GEMM over all rays of the
// image gives R2 and finish_z_kernel adds it (one 65 536-row GEMM instead of 32 of 2048 rows).
__global__ void park_r1_kernel(cpn_render_args a, int ray0, int nr, const float* __restrict__ r1, float* __restrict__ z_all) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * nr * CPN_LATENT) return;
  const int ray = i / CPN_LATENT, c = i % CPN_LATENT, b = ray / nr, n = ray0 + ray % nr;
  z_all[((size_t)b * a.N + n) * CPN_LATENT + c] = r1[i];
}
__global__ void finish_z_kernel(const float* __restrict__ r2, float* __restrict__ z_all, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float r = z_all[i];
  z_all[i] = (r2[i] + r) + r;
}

// combined readout: r2 = WVF sum_rows (w2 + 2 w1) h + b (one GEMM over the image) -> z = r2 + 2 b
__global__ void finish_z_bias_kernel(const float* __restrict__ r2, const float* __restrict__ bias, float* __restrict__ z_all,
                                     size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  z_all[i] = fmaf(2.f, bias[i % CPN_LATENT], r2[i]);
}

// phi. One CTA of 128 threads renders PHI_RAYS = 16 rays. Thread (cq = t % 32, rg = t / 32) owns a 4-channel x
// 4-ray register tile: per k one coalesced float4 of weights and one broadcast float4 of inputs feed 16 FMAs.
// Input of lin_z is cat(z, z) (both views carry the same latent, models/CoPoNeRF.py:545-552), coords18 is
// [plucker6 + origin3] of view 0 then view 1.
constexpr int PHI_RAYS = 16;

// acc[c][r] += sum_k act(in[k][rg*4 + r]) * (wT[k][cq*4 + c] (+ wT2[k][cq*4 + c]))
template <bool RELU_IN, bool TWO>
__device__ __forceinline__ void phi_dense(const float* __restrict__ wT, const float* __restrict__ wT2, int K,
                                          const float (*in)[PHI_RAYS], int cq, int rg, float (&acc)[4][4]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float4 w = __ldg(reinterpret_cast<const float4*>(wT + (size_t)k * CPN_HIDDEN) + cq);
    if (TWO) {
      float4 w2 = __ldg(reinterpret_cast<const float4*>(wT2 + (size_t)k * CPN_HIDDEN) + cq);
      w.x += w2.x; w.y += w2.y; w.z += w2.z; w.w += w2.w;
    }
    float4 x = *reinterpret_cast<const float4*>(&in[k][rg * 4]);
    if (RELU_IN) {
      x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
    }
    const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[c][r] = fmaf(wv[c], xv[r], acc[c][r]);
  }
}

__device__ __forceinline__ void phi_set_bias(const float* __restrict__ bias, int cq, float (&acc)[4][4]) {
  float4 b = __ldg(reinterpret_cast<const float4*>(bias) + cq);
  const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[c][r] = bv[c];
}

__global__ void __launch_bounds__(128) phi_kernel(cpn_render_args a, const float* __restrict__ z) {
  __shared__ __align__(16) float zt[CPN_LATENT][PHI_RAYS];   // transposed inputs: [k][ray]
  __shared__ __align__(16) float xt[CPN_HIDDEN][PHI_RAYS];
  __shared__ __align__(16) float ht[CPN_HIDDEN][PHI_RAYS];
  __shared__ __align__(16) float ct[20][PHI_RAYS];
  const float* W = reinterpret_cast<const float*>(a.weights);
  const int t = threadIdx.x, cq = t & 31, rg = t >> 5;
  const int total = a.B * a.N;   // ray index = b * N + n
  const int base = blockIdx.x * PHI_RAYS;
  for (int i = t; i < PHI_RAYS * CPN_LATENT; i += 128) {
    int r = i / CPN_LATENT, c = i % CPN_LATENT;
    zt[c][r] = (base + r < total) ? z[(size_t)(base + r) * CPN_LATENT + c] : 0.f;
  }
  for (int i = t; i < PHI_RAYS * 18; i += 128) {
    int r = i / 18, c = i % 18, v = c / 9, cc = c % 9;
    float val = 0.f;
    if (base + r < total) {
      int b = (base + r) / a.N, n = (base + r) % a.N;
      val = a.coords[(((size_t)(b * 2 + v)) * a.N + n) * 9 + cc];
    }
    ct[c][r] = val;
  }
  __syncthreads();
  float x[4][4], tmp[4][4];
  phi_set_bias(W + pw::PHI_BIN, cq, x);
  phi_dense<false, false>(W + pw::PHI_INT, nullptr, 18, ct, cq, rg, x);
  for (int blk = 0; blk < 3; ++blk) {
    // x += lin_z[blk](cat(z, z)): both halves of the weight see the same z
    const float* wz = W + pw::PHI_ZT + (size_t)blk * 2 * CPN_LATENT * CPN_HIDDEN;
    phi_set_bias(W + pw::PHI_BZ + blk * CPN_HIDDEN, cq, tmp);
    phi_dense<false, true>(wz, wz + (size_t)CPN_LATENT * CPN_HIDDEN, CPN_LATENT, zt, cq, rg, tmp);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        x[c][r] += tmp[c][r];
        xt[cq * 4 + c][rg * 4 + r] = x[c][r];
      }
    __syncthreads();
    // net = fc_0(relu(x)); dx = fc_1(relu(net)); x = x + dx   (lightfield.py:52-62)
    phi_set_bias(W + pw::PHI_B0 + blk * CPN_HIDDEN, cq, tmp);
    phi_dense<true, false>(W + pw::PHI_F0T + (size_t)blk * CPN_HIDDEN * CPN_HIDDEN, nullptr, CPN_HIDDEN, xt, cq, rg, tmp);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) ht[cq * 4 + c][rg * 4 + r] = tmp[c][r];
    __syncthreads();
    phi_set_bias(W + pw::PHI_B1 + blk * CPN_HIDDEN, cq, tmp);
    phi_dense<true, false>(W + pw::PHI_F1T + (size_t)blk * CPN_HIDDEN * CPN_HIDDEN, nullptr, CPN_HIDDEN, ht, cq, rg, tmp);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) x[c][r] += tmp[c][r];
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) xt[cq * 4 + c][rg * 4 + r] = fmaxf(x[c][r], 0.f);
  __syncthreads();
  // lin_out (3 x 128) + white background for rays with no valid epipolar segment (CoPoNeRF.py:562-566)
  if (t < PHI_RAYS * 3) {
    int r = t / 3, c = t % 3;
    if (base + r < total) {
      float acc = 0.f;
      for (int k = 0; k < CPN_HIDDEN; ++k) acc = fmaf(xt[k][r], W[pw::PHI_OUT + c * CPN_HIDDEN + k], acc);
      acc += W[pw::PHI_BOUT + c];
      float valid = a.valid_mask[base + r];   // written by ray_epilogue_kernel
      a.rgb[(size_t)(base + r) * 3 + c] = acc * valid + 1.f * (1.f - valid);
    }
  }
}

}  // namespace

int launch_attn1(const cpn_render_args& a, int ray0, int nr, const float* key, const float* qemb, const float* value,
                 const float* rowaux, float* r1, float* wp, cudaStream_t st, const float* logits, float* wts_out,
                 const float* gh, float* rbias) {
  int S2 = 2 * a.S;
  attn1_kernel<<<a.B * nr, S2, (S2 + 48) * sizeof(float), st>>>(a, ray0, nr, key, qemb, value, rowaux, r1, wp, logits,
                                                                wts_out, gh, rbias);
  CPN_CHECK_LAUNCH("attn1_kernel");
  return CPN_OK;
}

int launch_attn2(const cpn_render_args& a, int ray0, int nr, const float* q2, const float* qemb, const float* value,
                 const float* r1, float* z_all, cudaStream_t st, const float* logits, float* wts_out, const float* w1) {
  int S2 = 2 * a.S;
  attn2_kernel<<<a.B * nr, S2, (S2 + 8) * sizeof(float), st>>>(a, ray0, nr, q2, qemb, value, r1, z_all, logits, wts_out, w1);
  CPN_CHECK_LAUNCH("attn2_kernel");
  return CPN_OK;
}

int launch_readout_image(const cpn_render_args& a, int nr, const void* h1_image, const float* wts, float* hbar, int f8,
                         cudaStream_t st, int out_N, int out_ray0, int chunk_bytes) {
  const int S2 = 2 * a.S;
  if (S2 % 64) {
    cpn_set_error("readout_image: S=%d unsupported (S must be a multiple of 32)", a.S);
    return CPN_ERR_ARG;
  }
  if (out_N <= 0) {   // chunk-local output rows
    out_N = nr;
    out_ray0 = 0;
  }
  const size_t smem = (S2 + CPN_FEAT_DIM) * sizeof(float);
  if (f8)
    readout_image_kernel<true><<<dim3(a.B * nr, 2), 128, smem, st>>>(reinterpret_cast<const unsigned char*>(h1_image), wts, S2,
                                                                     hbar, nr, out_N, out_ray0, chunk_bytes);
  else
    readout_image_kernel<false><<<dim3(a.B * nr, 2), 128, smem, st>>>(reinterpret_cast<const unsigned char*>(h1_image), wts, S2,
                                                                      hbar, nr, out_N, out_ray0, chunk_bytes);
  CPN_CHECK_LAUNCH("readout_image_kernel");
  return CPN_OK;
}

int launch_park_r1(const cpn_render_args& a, int ray0, int nr, const float* r1, float* z_all, cudaStream_t st) {
  const int total = a.B * nr * CPN_LATENT;
  park_r1_kernel<<<(total + 255) / 256, 256, 0, st>>>(a, ray0, nr, r1, z_all);
  CPN_CHECK_LAUNCH("park_r1_kernel");
  return CPN_OK;
}

int launch_finish_z(const cpn_render_args& a, const float* r2_all, float* z_all, cudaStream_t st) {
  const size_t total = (size_t)a.B * a.N * CPN_LATENT;
  finish_z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(r2_all, z_all, total);
  CPN_CHECK_LAUNCH("finish_z_kernel");
  return CPN_OK;
}

int launch_finish_z_bias(const cpn_render_args& a, const float* r2_all, const float* bias, float* z_all, cudaStream_t st) {
  const size_t total = (size_t)a.B * a.N * CPN_LATENT;
  finish_z_bias_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(r2_all, bias, z_all, total);
  CPN_CHECK_LAUNCH("finish_z_bias_kernel");
  return CPN_OK;
}

int launch_phi(const cpn_render_args& a, const float* z_all, cudaStream_t st) {
  int total = a.B * a.N;
  phi_kernel<<<(total + PHI_RAYS - 1) / PHI_RAYS, 128, 0, st>>>(a, z_all);
  CPN_CHECK_LAUNCH("phi_kernel");
  return CPN_OK;
}
