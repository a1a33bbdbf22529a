// Centre-pivot 4-D convolution block of the cost aggregation: Conv4d -> GroupNorm(1 group) -> ReLU
// (models/conv4d.py:57-135 Conv4d, :7-32 MaxPool4d, :138-163 Encoder4D; 63 calls per stereo pair).
//
// Conv4d is the sum of a 2-D convolution over the query axes and one over the support axes; with stride s the
// branch that convolves one pair of axes first max-pools the other pair by s (ceil_mode). Here one thread owns one
// output position (b, hq, wq, hs, ws) and all Co output channels: every input tap is read (and pooled on the fly)
// once, the weights sit in shared memory as [ci][tap][co] and are read as warp-wide broadcasts. No rearranged or
// pooled copy of the 4-D volume is ever materialised (the reference's einops rearranges are 44 % of its UFC time).
// GroupNorm over all Co x positions of a sample is a two-phase reduction: per-CTA (sum, sum of squares) in
// double, then a normalise + ReLU pass over the 2-8 MB output.
#include <math.h>
#include <stdlib.h>
#include "cpn_common.cuh"

namespace {

constexpr int C4_POS = 32;        // output positions per CTA (one per lane: consecutive ws -> coalesced taps and stores)
constexpr int C4_WARPS = 8;       // the (ci, branch, dy) tap rows of a position are dealt round-robin to 8 warps
constexpr int C4_THREADS = C4_POS * C4_WARPS;
constexpr int C4_GROUPS = 4;      // position groups per CTA
constexpr int C4_CTA_POS = C4_POS * C4_GROUPS;

// K, S: kernel size and stride as compile-time constants (0 = take them from the arguments), so that the tap and
// pooling loops unroll for the three geometries UFC uses: (3, 1), (3, 2), (5, 4).
//
// The volumes are small (16^4 = 65 536 output positions), so one thread per position left the SMs at 20 % occupancy
// with every thread walking Ci x 2 x k x k dependent tap loads (80 us for 75 MFMA). Here a position is shared by 8
// warps, each taking whole tap rows (ci, branch, dy); the partial sums meet in shared memory and are added in warp
// order (fixed, so the result does not depend on the launch).
template <int CO, int K, int S>
__global__ void __launch_bounds__(C4_THREADS) conv4d_kernel(cpn_conv4d_args a, int oq, int os, double* __restrict__ partials) {
  extern __shared__ __align__(16) float wsm[];   // [2 branches][Ci][k*k][CO], then the partial sums [8 warps][CO][32 positions]
  const int k = K ? K : a.k, kk = k * k, s = S ? S : a.stride, p = a.pad, Ci = a.Ci, Hq = a.Hq, Hs = a.Hs;
  {  // weights: coalesced reads of the (Co, Ci, k, k) tensors, scattered into the [ci][tap][co] shared layout
    const int per = Ci * kk * CO;
    for (int j = threadIdx.x; j < 2 * per; j += C4_THREADS) {
      const int br = j >= per, jj = j - br * per, tap = jj % kk, ci = (jj / kk) % Ci, co = jj / (kk * Ci);
      wsm[((size_t)(br * Ci + ci) * kk + tap) * CO + co] = (br ? a.ws : a.wq)[jj];
    }
  }
  float* red = wsm + 2 * Ci * kk * CO;
  const int P = oq * oq * os * os;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  double sum = 0.0, sq = 0.0;
  // C4_GROUPS groups of 32 positions per CTA: the weight staging and the CTA launch are paid once per 128 positions
  for (int grp = 0; grp < C4_GROUPS; ++grp) {
  const int pos0 = (blockIdx.x * C4_GROUPS + grp) * C4_POS, pos = pos0 + lane;
  __syncthreads();   // weights staged (first group) / partial sums of the previous group consumed
  if (pos0 >= P) break;
  float acc[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) acc[co] = 0.f;
  if (pos < P) {
    const int ws_ = pos % os, hs = (pos / os) % os, wq = (pos / (os * os)) % oq, hq = pos / (os * os * oq);
    const size_t sHs = (size_t)Hs, plane = (size_t)Hq * Hq * Hs * Hs;
    const float* xb = a.x + (size_t)b * Ci * plane;
    if (K > 0 && Ci >= C4_WARPS) {
      // Tap table of this position, shared by all input channels: element offset of every tap (of its pooling window)
      // inside one channel plane, or -1 outside the padded volume. A warp then owns whole (ci, branch) pairs, so a
      // tap costs one load (per pooled element), two 16-byte weight loads per 8 outputs and the FMAs.
      constexpr int KK = K > 0 ? K * K : 1, SS = S > 0 ? S : 1;
      int qoff[KK], soff[KK];
#pragma unroll
      for (int dy = 0; dy < (K > 0 ? K : 1); ++dy)
#pragma unroll
        for (int dx = 0; dx < (K > 0 ? K : 1); ++dx) {
          const int qy = hq * SS + dy - p, qx = wq * SS + dx - p, sy = hs * SS + dy - p, sx = ws_ * SS + dx - p;
          qoff[dy * K + dx] = (qy < 0 || qy >= Hq || qx < 0 || qx >= Hq) ? -1 : ((qy * Hq + qx) * Hs + hs * SS) * Hs + ws_ * SS;
          soff[dy * K + dx] = (sy < 0 || sy >= Hs || sx < 0 || sx >= Hs) ? -1 : ((hq * SS * Hq + wq * SS) * Hs + sy) * Hs + sx;
        }
      const int nqi = min(SS, Hs - hs * SS), nqj = min(SS, Hs - ws_ * SS), nsi = min(SS, Hq - hq * SS), nsj = min(SS, Hq - wq * SS);
      const int sstep = Hs * Hs;
      for (int it = warp; it < 2 * Ci; it += C4_WARPS) {
        const int br = it & 1, ci = it >> 1;
        const float* xc = xb + (size_t)ci * plane;
        const float* wrow = wsm + (size_t)(br * Ci + ci) * KK * CO;
        if (br == 0) {
#pragma unroll
          for (int tp = 0; tp < KK; ++tp) {
            float v;
            if (S == 1) {   // no pooling: a padded tap is a zero, no branch
              v = qoff[tp] >= 0 ? __ldg(xc + qoff[tp]) : 0.f;
            } else {
              if (qoff[tp] < 0) continue;
              const float* base = xc + qoff[tp];
              v = -INFINITY;
#pragma unroll
              for (int i = 0; i < SS; ++i)
#pragma unroll
                for (int j = 0; j < SS; ++j)
                  if (i < nqi && j < nqj) v = fmaxf(v, __ldg(base + i * Hs + j));
            }
#pragma unroll
            for (int co = 0; co < CO; co += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(wrow + tp * CO + co);
              acc[co] = fmaf(v, w4.x, acc[co]);
              acc[co + 1] = fmaf(v, w4.y, acc[co + 1]);
              acc[co + 2] = fmaf(v, w4.z, acc[co + 2]);
              acc[co + 3] = fmaf(v, w4.w, acc[co + 3]);
            }
          }
        } else {
#pragma unroll
          for (int tp = 0; tp < KK; ++tp) {
            float v;
            if (S == 1) {
              v = soff[tp] >= 0 ? __ldg(xc + soff[tp]) : 0.f;
            } else {
              if (soff[tp] < 0) continue;
              const float* base = xc + soff[tp];
              v = -INFINITY;
#pragma unroll
              for (int i = 0; i < SS; ++i)
#pragma unroll
                for (int j = 0; j < SS; ++j)
                  if (i < nsi && j < nsj) v = fmaxf(v, __ldg(base + (size_t)(i * Hq + j) * sstep));
            }
#pragma unroll
            for (int co = 0; co < CO; co += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(wrow + tp * CO + co);
              acc[co] = fmaf(v, w4.x, acc[co]);
              acc[co + 1] = fmaf(v, w4.y, acc[co + 1]);
              acc[co + 2] = fmaf(v, w4.z, acc[co + 2]);
              acc[co + 3] = fmaf(v, w4.w, acc[co + 3]);
            }
          }
        }
      }
    } else {
    const int items = Ci * 2 * k;
    for (int it = warp; it < items; it += C4_WARPS) {
      const int dy = it % k, br = (it / k) & 1, ci = it / (2 * k);
      const float* xc = xb + (size_t)ci * plane;
      const float* wrow = wsm + ((size_t)(br * Ci + ci) * kk + dy * k) * CO;
      if (br == 0) {
        // query branch: conv over (Hq, Wq) of the input max-pooled over the support window of this output position
        const int qy = hq * s + dy - p;
        if (qy < 0 || qy >= Hq) continue;
#pragma unroll
        for (int dx = 0; dx < k; ++dx) {
          const int qx = wq * s + dx - p;
          if (qx < 0 || qx >= Hq) continue;
          const float* base = xc + ((size_t)qy * Hq + qx) * sHs * sHs;
          float v = -INFINITY;
#pragma unroll
          for (int i = 0; i < s; ++i) {
            const int y = hs * s + i;
            if (y >= Hs) break;
#pragma unroll
            for (int j = 0; j < s; ++j) {
              const int x = ws_ * s + j;
              if (x >= Hs) break;
              v = fmaxf(v, __ldg(base + (size_t)y * Hs + x));
            }
          }
          const float* w = wrow + dx * CO;
#pragma unroll
          for (int co = 0; co < CO; ++co) acc[co] = fmaf(v, w[co], acc[co]);
        }
      } else {
        // support branch: conv over (Hs, Ws) of the input max-pooled over the query window
        const int sy = hs * s + dy - p;
        if (sy < 0 || sy >= Hs) continue;
#pragma unroll
        for (int dx = 0; dx < k; ++dx) {
          const int sx = ws_ * s + dx - p;
          if (sx < 0 || sx >= Hs) continue;
          float v = -INFINITY;
#pragma unroll
          for (int i = 0; i < s; ++i) {
            const int y = hq * s + i;
            if (y >= Hq) break;
#pragma unroll
            for (int j = 0; j < s; ++j) {
              const int x = wq * s + j;
              if (x >= Hq) break;
              v = fmaxf(v, __ldg(xc + (((size_t)y * Hq + x) * sHs + sy) * sHs + sx));
            }
          }
          const float* w = wrow + dx * CO;
#pragma unroll
          for (int co = 0; co < CO; ++co) acc[co] = fmaf(v, w[co], acc[co]);
        }
      }
    }
    }   // generic path
  }
#pragma unroll
  for (int co = 0; co < CO; ++co) red[(warp * CO + co) * C4_POS + lane] = acc[co];
  __syncthreads();
  // (co, position) pairs: add the 8 partial sums in warp order, add the biases, store, GroupNorm partial sums
  for (int i = threadIdx.x; i < CO * C4_POS; i += C4_THREADS) {
    const int co = i / C4_POS, l = i % C4_POS, ppos = pos0 + l;
    if (ppos >= P) continue;
    float v = red[co * C4_POS + l];
#pragma unroll
    for (int w = 1; w < C4_WARPS; ++w) v += red[(w * CO + co) * C4_POS + l];
    v += a.bq[co] + a.bs[co];
    a.y[((size_t)b * CO + co) * P + ppos] = v;
    sum += (double)v;
    sq += (double)v * (double)v;
  }
  }  // position groups
  // per-CTA partial sums for GroupNorm (fixed order: warp tree, then warps in order)
  __shared__ double red2[2][C4_WARPS];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) {
    red2[0][warp] = sum;
    red2[1][warp] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = 0; i < C4_WARPS; ++i) {
      t0 += red2[0][i];
      t1 += red2[1][i];
    }
    partials[((size_t)b * gridDim.x + blockIdx.x) * 2 + 0] = t0;
    partials[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = t1;
  }
}

// ---- stride-1, 3 x 3 blocks (55 of the 63 Conv4d calls of a pair) ---------------------------------------------------
// The kernel above fetches every tap from global memory per output position (2.9 TFLOP/s: tap-load latency). For k = 3,
// stride 1 the inputs of a whole (query position, band of 256 support positions) tile are nine contiguous planes: a CTA
// stages them in shared memory once per group of 8 input channels (8 neighbour bands + the zero-padded centre band, 76 KB)
// and every thread then runs pure shared-memory -> FMA loops for two output positions and all Co channels
// (one staged value feeds Co FMAs, one weight vector two positions). Sum order: ci ascending, query taps then support
// taps, fixed -- so the result does not depend on the launch.
constexpr int S1_THREADS = 128, S1_POS = 256, S1_CI = 8;

template <int CO>
__global__ void __launch_bounds__(S1_THREADS) conv4d_s1_kernel(cpn_conv4d_args a, double* __restrict__ partials) {
  extern __shared__ __align__(16) float sm1[];
  const int Ci = a.Ci, Hq = a.Hq, Hs = a.Hs;
  const int SB = S1_POS / Hs;                  // support rows per band
  const int nbands = Hs / SB;
  const int PW = Hs + 2, CTR = (SB + 2) * PW;  // padded centre band
  float* wq_s = sm1;                           // [Ci][9][CO]
  float* ws_s = wq_s + Ci * 9 * CO;
  float* nb = ws_s + Ci * 9 * CO;              // [8 neighbours][S1_CI][S1_POS]
  float* ctr = nb + 8 * S1_CI * S1_POS;        // [S1_CI][CTR]
  const int t = threadIdx.x, b = blockIdx.y;
  const int band = blockIdx.x % nbands, q = blockIdx.x / nbands, qy = q / Hq, qx = q % Hq;
  const int row0 = band * SB;
  for (int j = t; j < Ci * 9 * CO; j += S1_THREADS) {   // (Co, Ci, 3, 3) -> [ci][tap][co]
    const int tap = j % 9, ci = (j / 9) % Ci, co = j / (9 * Ci);
    wq_s[(ci * 9 + tap) * CO + co] = a.wq[j];
    ws_s[(ci * 9 + tap) * CO + co] = a.ws[j];
  }
  const size_t plane = (size_t)Hs * Hs, cplane = (size_t)Hq * Hq * plane;
  const float* xb = a.x + (size_t)b * Ci * cplane;
  float acc[2][CO];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[i][co] = 0.f;
  // this thread's two positions inside the band and their offsets in the padded centre band
  const int p0 = t, p1 = t + S1_THREADS;
  const int c0 = (p0 / Hs + 1) * PW + p0 % Hs + 1, c1 = (p1 / Hs + 1) * PW + p1 % Hs + 1;
  for (int ci0 = 0; ci0 < Ci; ci0 += S1_CI) {
    __syncthreads();   // weights staged / the previous channel group consumed
    // 8 neighbour bands (zeros outside the query grid): float4 copies of 1 KB runs, eight loads in flight per thread (the
    // first version issued one load per loop iteration and spent 80 % of the kernel waiting for L2)
    constexpr int NB_V4 = 8 * S1_CI * (S1_POS / 4);
    static_assert(NB_V4 % (S1_THREADS * 8) == 0, "neighbour bands: whole batches of 8 loads per thread");
#pragma unroll 1
    for (int j0 = t; j0 < NB_V4; j0 += S1_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * S1_THREADS;
        const int v4 = j % (S1_POS / 4), ci = (j / (S1_POS / 4)) % S1_CI, n = j / ((S1_POS / 4) * S1_CI);
        const int tap = n < 4 ? n : n + 1, ny = qy + tap / 3 - 1, nx = qx + tap % 3 - 1;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ny >= 0 && ny < Hq && nx >= 0 && nx < Hq)
          v[u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(ci0 + ci) * cplane + ((size_t)ny * Hq + nx) * plane + (size_t)row0 * Hs) + v4);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) reinterpret_cast<float4*>(nb)[j0 + u * S1_THREADS] = v[u];
    }
    // centre band with a one-element zero halo (rows above / below the band come from the plane itself)
    const int nctr = S1_CI * CTR;
#pragma unroll 1
    for (int j0 = t; j0 < nctr; j0 += S1_THREADS * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * S1_THREADS;
        v[u] = 0.f;
        if (j < nctr) {
          const int ci = j / CTR, r = (j % CTR) / PW, c = (j % CTR) % PW;
          const int sy = row0 + r - 1, sx = c - 1;
          if (sy >= 0 && sy < Hs && sx >= 0 && sx < Hs)
            v[u] = __ldg(xb + (size_t)(ci0 + ci) * cplane + ((size_t)qy * Hq + qx) * plane + (size_t)sy * Hs + sx);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (j0 + u * S1_THREADS < nctr) ctr[j0 + u * S1_THREADS] = v[u];
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < S1_CI; ++ci) {
      const float* wq_c = wq_s + (size_t)(ci0 + ci) * 9 * CO;
      const float* ws_c = ws_s + (size_t)(ci0 + ci) * 9 * CO;
      const float* cc = ctr + ci * CTR;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {      // query branch: the same support position in the nine neighbouring planes
        float v0, v1;
        if (tap == 4) {
          v0 = cc[c0];
          v1 = cc[c1];
        } else {
          const float* np = nb + ((size_t)(tap < 4 ? tap : tap - 1) * S1_CI + ci) * S1_POS;
          v0 = np[p0];
          v1 = np[p1];
        }
#pragma unroll
        for (int co = 0; co < CO; co += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wq_c + tap * CO + co);
          acc[0][co] = fmaf(v0, w4.x, acc[0][co]); acc[0][co + 1] = fmaf(v0, w4.y, acc[0][co + 1]);
          acc[0][co + 2] = fmaf(v0, w4.z, acc[0][co + 2]); acc[0][co + 3] = fmaf(v0, w4.w, acc[0][co + 3]);
          acc[1][co] = fmaf(v1, w4.x, acc[1][co]); acc[1][co + 1] = fmaf(v1, w4.y, acc[1][co + 1]);
          acc[1][co + 2] = fmaf(v1, w4.z, acc[1][co + 2]); acc[1][co + 3] = fmaf(v1, w4.w, acc[1][co + 3]);
        }
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {      // support branch: the 3 x 3 neighbourhood inside the centre plane
        const int off = (tap / 3 - 1) * PW + tap % 3 - 1;
        const float v0 = cc[c0 + off], v1 = cc[c1 + off];
#pragma unroll
        for (int co = 0; co < CO; co += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(ws_c + tap * CO + co);
          acc[0][co] = fmaf(v0, w4.x, acc[0][co]); acc[0][co + 1] = fmaf(v0, w4.y, acc[0][co + 1]);
          acc[0][co + 2] = fmaf(v0, w4.z, acc[0][co + 2]); acc[0][co + 3] = fmaf(v0, w4.w, acc[0][co + 3]);
          acc[1][co] = fmaf(v1, w4.x, acc[1][co]); acc[1][co + 1] = fmaf(v1, w4.y, acc[1][co + 1]);
          acc[1][co + 2] = fmaf(v1, w4.z, acc[1][co + 2]); acc[1][co + 3] = fmaf(v1, w4.w, acc[1][co + 3]);
        }
      }
    }
  }
  // biases, store (consecutive threads -> consecutive support positions), GroupNorm partial sums in double
  const size_t P = cplane;
  const size_t pos_base = (size_t)q * plane + (size_t)row0 * Hs;
  double sum = 0.0, sq = 0.0;
#pragma unroll
  for (int co = 0; co < CO; ++co) {
    const float bb = a.bq[co] + a.bs[co];
    const float v0 = acc[0][co] + bb, v1 = acc[1][co] + bb;
    float* yo = a.y + ((size_t)b * CO + co) * P + pos_base;
    yo[p0] = v0;
    yo[p1] = v1;
    sum += (double)v0 + (double)v1;
    sq += (double)v0 * (double)v0 + (double)v1 * (double)v1;
  }
  __shared__ double red2[2][S1_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if ((t & 31) == 0) {
    red2[0][t >> 5] = sum;
    red2[1][t >> 5] = sq;
  }
  __syncthreads();
  if (t == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = 0; i < S1_THREADS / 32; ++i) {
      t0 += red2[0][i];
      t1 += red2[1][i];
    }
    partials[((size_t)b * gridDim.x + blockIdx.x) * 2 + 0] = t0;
    partials[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = t1;
  }
}

// GroupNorm(1, Co) (biased variance, eps 1e-5, per-channel affine) + ReLU, in place.
__global__ void __launch_bounds__(256) gn_relu_kernel(float* __restrict__ y, const double* __restrict__ partials, int nparts,
                                                      int Co, int P, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta) {
  __shared__ double red[2][8];
  __shared__ float stat[2];
  const int b = blockIdx.y;
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) {
    s0 += partials[((size_t)b * nparts + i) * 2 + 0];
    s1 += partials[((size_t)b * nparts + i) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0;
    red[1][threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = 0; i < 8; ++i) {
      t0 += red[0][i];
      t1 += red[1][i];
    }
    const double n = (double)Co * (double)P, mean = t0 / n, var = t1 / n - mean * mean;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + 1e-5));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  const size_t total = (size_t)Co * P;
  float* yb = y + (size_t)b * total;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int co = (int)(i / P);
    float v = (yb[i] - mean) * rstd * gamma[co] + beta[co];
    yb[i] = fmaxf(v, 0.f);
  }
}

int out_size(int H, int k, int s, int p) { return (H + 2 * p - k) / s + 1; }

}  // namespace

extern "C" size_t cpn_conv4d_workspace_bytes(int B, int Hq, int Hs, int k, int stride, int pad) {
  if (B <= 0 || Hq <= 0 || Hs <= 0 || k <= 0 || stride <= 0) return 0;
  int oq = out_size(Hq, k, stride, pad), os = out_size(Hs, k, stride, pad);
  size_t P = (size_t)oq * oq * os * os;
  return (size_t)B * ((P + C4_CTA_POS - 1) / C4_CTA_POS) * 2 * sizeof(double);
}

extern "C" int cpn_conv4d(const cpn_conv4d_args* args, void* stream) {
  if (!args) {
    cpn_set_error("cpn_conv4d: null args");
    return CPN_ERR_ARG;
  }
  const cpn_conv4d_args& a = *args;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.B <= 0 || a.Ci <= 0 || (a.Co != 8 && a.Co != 32) || a.Hq <= 0 || a.Hs <= 0 || a.k <= 0 || a.stride <= 0 || a.pad < 0 ||
      !a.x || !a.wq || !a.bq || !a.ws || !a.bs || !a.y || !a.workspace || (a.norm_relu && (!a.gamma || !a.beta))) {
    cpn_set_error("cpn_conv4d: bad argument (Co must be 8 or 32)");
    return CPN_ERR_ARG;
  }
  const int oq = out_size(a.Hq, a.k, a.stride, a.pad), os = out_size(a.Hs, a.k, a.stride, a.pad);
  // the pooled axes must give the same output size as the convolved ones (MaxPool4d, ceil_mode)
  if (oq <= 0 || os <= 0 || (a.Hq + a.stride - 1) / a.stride != oq || (a.Hs + a.stride - 1) / a.stride != os) {
    cpn_set_error("cpn_conv4d: inconsistent output sizes for Hq=%d Hs=%d k=%d stride=%d pad=%d", a.Hq, a.Hs, a.k, a.stride, a.pad);
    return CPN_ERR_ARG;
  }
  const size_t need = cpn_conv4d_workspace_bytes(a.B, a.Hq, a.Hs, a.k, a.stride, a.pad);
  if (need > a.workspace_bytes) {
    cpn_set_error("cpn_conv4d: workspace of %zu bytes needed, %zu given", need, a.workspace_bytes);
    return CPN_ERR_WORKSPACE;
  }
  double* partials = reinterpret_cast<double*>(a.workspace);
  // stride-1 3 x 3 blocks with whole 256-position bands and channel groups of 8: the shared-memory tile kernel
  static int direct_only = -1;   // CPN_CONV4D_DIRECT=1: the per-position kernel for every geometry (A/B runs)
  if (direct_only < 0) {
    const char* e = getenv("CPN_CONV4D_DIRECT");
    direct_only = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (!direct_only && a.k == 3 && a.stride == 1 && a.pad == 1 && a.Hq == oq && a.Hs == os && (a.Ci % S1_CI) == 0 &&
      a.Hs <= S1_POS && (S1_POS % a.Hs) == 0 && (a.Hs % (S1_POS / a.Hs)) == 0 && (a.Hs % 4) == 0) {
    const int nbands = a.Hs / (S1_POS / a.Hs), nblk1 = a.Hq * a.Hq * nbands;
    const size_t smem1 = ((size_t)2 * a.Ci * 9 * a.Co + (size_t)8 * S1_CI * S1_POS +
                          (size_t)S1_CI * (S1_POS / a.Hs + 2) * (a.Hs + 2)) * sizeof(float);
    void (*k1)(cpn_conv4d_args, double*) = a.Co == 8 ? conv4d_s1_kernel<8> : conv4d_s1_kernel<32>;
    CPN_CHECK_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    k1<<<dim3(nblk1, a.B), S1_THREADS, smem1, st>>>(a, partials);
    CPN_CHECK_LAUNCH("conv4d_s1_kernel");
    if (a.norm_relu) {
      dim3 g2(256, a.B);
      gn_relu_kernel<<<g2, 256, 0, st>>>(a.y, partials, nblk1, a.Co, oq * oq * os * os, a.gamma, a.beta);
      CPN_CHECK_LAUNCH("gn_relu_kernel");
    }
    return CPN_OK;
  }
  const int P = oq * oq * os * os, nblk = (P + C4_CTA_POS - 1) / C4_CTA_POS;
  const size_t smem = ((size_t)2 * a.Ci * a.k * a.k * a.Co + (size_t)C4_WARPS * a.Co * C4_POS) * sizeof(float);
  if (smem > 96 * 1024) {
    cpn_set_error("cpn_conv4d: weights + partial sums of %zu bytes do not fit in shared memory", smem);
    return CPN_ERR_ARG;
  }
  dim3 grid(nblk, a.B);
  void (*kern)(cpn_conv4d_args, int, int, double*);
  const int ks = a.k * 10 + a.stride;
  if (a.Co == 8)
    kern = ks == 31 ? conv4d_kernel<8, 3, 1> : ks == 32 ? conv4d_kernel<8, 3, 2> : ks == 54 ? conv4d_kernel<8, 5, 4>
                                                                                           : conv4d_kernel<8, 0, 0>;
  else
    kern = ks == 31 ? conv4d_kernel<32, 3, 1> : conv4d_kernel<32, 0, 0>;
  CPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, C4_THREADS, smem, st>>>(a, oq, os, partials);
  CPN_CHECK_LAUNCH("conv4d_kernel");
  if (a.norm_relu) {
    dim3 g2(256, a.B);
    gn_relu_kernel<<<g2, 256, 0, st>>>(a.y, partials, nblk, a.Co, P, a.gamma, a.beta);
    CPN_CHECK_LAUNCH("gn_relu_kernel");
  }
  return CPN_OK;
}
